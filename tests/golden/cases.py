"""Named step() parity cases shared by make_golden.py and the tests."""
from opensc2_b200.problem import case1_topology, case2_topology, case3_topology

# name -> (topology factory, factory kwargs, n_nodes, synthetic_step_arrays kwargs)
STEP_CASES = {
    "case1_pdrop_n21": (case1_topology, dict(intial=1), 21, dict()),
    "case1_pinl_vout_first_step_n17": (case1_topology, dict(intial=2), 17, dict(num_step=1)),
    "case1_vinl_pout_cn_refined_n12": (case1_topology, dict(intial=3, theta=0.5, method="CN"), 12, dict(refined_mesh=True)),
    "case1_min_mesh_n4": (case1_topology, dict(intial=2), 4, dict()),
    "case1_shipped_size_n201": (case1_topology, dict(intial=1), 201, dict()),
    "case1_inexact_squares_n64": (case1_topology, dict(intial=1), 64, dict(exact_squares=False, seed=7)),
    "case3_backward_n15": (case3_topology, dict(), 15, dict(refined_mesh=True)),
    "case3_rectangular_env_n9": (case3_topology, dict(is_rectangular=True, height=0.3, width=0.2, es_choice2=(True,)), 9, dict()),
    "case3_shipped_size_n201": (case3_topology, dict(), 201, dict(seed=11)),
    "case2_n6": (case2_topology, dict(), 6, dict()),
    "case2_first_step_n5": (case2_topology, dict(), 5, dict(num_step=1, seed=3)),
}
# cases whose pre-solve SYSMAT/Known are stored too (kept small)
WITH_SYSTEM = {"case1_pdrop_n21", "case1_pinl_vout_first_step_n17", "case1_vinl_pout_cn_refined_n12",
               "case1_min_mesh_n4", "case3_backward_n15", "case3_rectangular_env_n9"}


def make_case(name):
    from opensc2_b200.problem import synthetic_step_arrays
    fac, fkw, n, skw = STEP_CASES[name]
    topo = fac(**fkw)
    return topo, synthetic_step_arrays(topo, n, **skw)
