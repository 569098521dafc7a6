"""GPU parity against the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py:
refined-truth outputs of the global KKT solve on seeded inputs).  Every kernel family is covered: thread per
instance (cartpole, double integrator, Dubins), warp / half-warp per instance (n=12), CTA per instance (n=64)."""
import os

import numpy as np
import pytest

from lqr_b200 import ops

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name,kernel", [("cartpole_kkt", "kkt_tpi<4,1"), ("double_integrator_kkt", "kkt_tpi<6,3"),
                                         ("dubins_kkt", "kkt_tpi<3,2"), ("quad_kkt", "kkt_hw<12,4"),
                                         ("large_kkt", "kkt_cta_dmma<64,16"), ("mid32_kkt", "kkt_cta_dmma<32,8"),
                                         ("mid24_kkt", "kkt_cta_dmma<24,8"), ("dubins_stage_kkt", "kkt_tpi<3,2,p=3/1/3"),
                                         ("explicit_d2_kkt", "kkt_coop"), ("quad_stage_kkt", "kkt_wp_dmma<12,4,p=12/2/12,hess=0"),
                                         ("large_stage_kkt", "kkt_cta_dmma<64,16,p=64/1/64"),
                                         ("arm_padded_kkt", "kkt_cta_dmma<16,8,p=16/1/16"),
                                         ("free_final_kkt", "kkt_wp_dmma<12,4,p=12/1/12")])
def test_kkt_golden(handle, oracle_mod, name, kernel):
    from tests.golden.make_golden import KKT_CASES
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    prob = KKT_CASES[name]()
    dz, lam, info = ops.kkt_solve_problem(prob, handle=handle)
    assert handle.last_kernel.startswith(kernel) and (info == 0).all()
    # 1e-10 against the golden (extended-precision) vectors; where the reference's own operation order (the CPU
    # oracle) is itself further than that from them (short-horizon n >= 24 cases, cond ~1e6), 4x its error
    dzo, lamo, _ = oracle_mod.kkt_solve(prob)
    tol = max(TOL, 4.0 * max(_rel(dzo[0], z["dz"]), _rel(lamo[0], z["mult"])))
    assert _rel(dz[0], z["dz"]) <= tol and _rel(lam[0], z["mult"]) <= tol, (tol, handle.last_kernel)


@pytest.mark.parametrize("name,kernel", [("cartpole_riccati", "riccati_tpi<4,1"), ("quad_riccati", "riccati_dmma<12,4"),
                                         ("large_riccati", "riccati_cta_dmma<64,16"), ("mid24_riccati", "riccati_cta_dmma<24,8"),
                                         ("lti_riccati", "riccati_dmma<8,4"), ("padded_riccati", "riccati_dmma<12,3")])
def test_riccati_golden(handle, name, kernel):
    from tests.golden.make_golden import RICCATI_CASES
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    X, U, _, _, info = ops.riccati_solve_problem(RICCATI_CASES[name](), want_gains=False, handle=handle)
    assert handle.last_kernel.startswith(kernel) and (info == 0).all()
    assert _rel(X[: z["X"].shape[0]], z["X"]) <= TOL and _rel(U[: z["U"].shape[0]], z["U"]) <= TOL
