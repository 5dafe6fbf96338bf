"""Do the element-wise kernels (K_P, K_G) overlap with the latency-bound sweep K_F when they run on another stream?
Two handles (own streams): A steps (K_G + K_F + K_B + K_N on given coefficients), B evaluates operating conditions
(K_P).  Alone vs together, host-timed around synchronised regions."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from bench import deck_image
from opensc2_b200.model import synthetic_helium_table
from opensc2_b200.runner import Transient

batch, nodes = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 3001
deck, tab = deck_image("CASE_1_ITER_like_LTS"), synthetic_helium_table()
rng = np.random.default_rng(0)
A = Transient(deck, [tab], batch, n_elems=nodes - 1, q0=rng.uniform(50, 1000, batch))
B = Transient(deck, [tab], batch, n_elems=nodes - 1, q0=rng.uniform(50, 1000, batch))
for tr in (A, B):
    tr.bs.advance(0.1, 0.1, 1); tr.bs.advance(0.2, 0.1, 2); tr.bs.synchronize()
def sync():
    A.bs.synchronize(); B.bs.synchronize()
def timeit(fa, fb, reps=6):
    sync(); t0 = time.perf_counter()
    for k in range(reps):
        if fa: fa(k)
        if fb: fb(k)
    sync(); return (time.perf_counter() - t0) / reps * 1e3
step_a = lambda k: A.bs.step(0.1, 3 + k)
props_b = lambda k: (B.bs.eval_operating_conditions(0.3), B.bs.eval_operating_conditions(0.3))
ta, tb, tab_ = timeit(step_a, None), timeit(None, props_b), timeit(step_a, props_b)
print(f"nodes {nodes}: A step alone {ta:.2f} ms, B 2 x props alone {tb:.2f} ms, together {tab_:.2f} ms (sum {ta + tb:.2f})")
step_b = lambda k: B.bs.step(0.1, 3 + k)
tb2, tab2 = timeit(None, step_b), timeit(step_a, step_b)
print(f"A step + B step: alone {ta:.2f} + {tb2:.2f}, together {tab2:.2f}")
