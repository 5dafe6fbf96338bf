"""opensc2_b200 -- batched B200 (sm_100a) implementation of OPENSC2's transient time step.

Public surface:
  * `opensc2_b200.batch.BatchedStep` -- B conductors per call through the C ABI
    (include/opensc2_b200.h);
  * `opensc2_b200.step.step` -- drop-in for the reference's
    `step(conductor, environment, qsource, num_step)`;
  * `opensc2_b200.problem` -- neutral topology/state description.
The CUDA library is loaded lazily and there is no CPU fallback.
"""
__version__ = "0.1.0"
