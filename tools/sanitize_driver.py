"""Small-batch run of every kernel family through the C ABI with host buffers (no torch), for compute-sanitizer
(memcheck / racecheck / synccheck): batches <= 64, short horizons.  Prints the kernels that ran."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from lqr_b200 import _lib, ops, problems  # noqa: E402
from lqr_b200.sqp import DubinsSQP  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
h = _lib.Handle(0)
ran = []


def ric(n, m, N, b, **kw):
    X, U, K, kff, info = ops.riccati_solve_problem(problems.random_lqr_riccati(n, m, N, b, seed=1, **kw), handle=h)
    assert (info == 0).all() and np.isfinite(X).all()
    ran.append(h.last_kernel)


def kkt(n, m, N, b, soc=False, **kw):
    dz, lam, info, res = ops.kkt_solve_problem(problems.random_lqr_kkt(n, m, N, b, seed=2, **kw), soc=soc, want_res=True, handle=h)
    assert (info == 0).all() and np.isfinite(dz).all()
    ran.append(h.last_kernel)


if which in ("all", "riccati"):
    ric(4, 1, 12, 40)            # riccati_tpi
    ric(12, 4, 12, 9)            # riccati_dmma (warp per instance, bulk copies + mbarrier)
    ric(64, 16, 6, 3)            # riccati_cta_dmma
    ric(24, 8, 6, 3)
    ric(5, 2, 8, 5)              # riccati_coop
    ric(4, 1, 12, 40, lti=True)
if which in ("all", "kkt"):
    kkt(3, 2, 12, 40)            # kkt_tpi
    kkt(6, 3, 8, 10, mid_p=1, hess_mode=0)
    kkt(12, 4, 12, 9)            # kkt_hinv + kkt_hw2 (half warp per instance)
    kkt(12, 4, 12, 9, soc=True)
    kkt(64, 16, 6, 3)            # kkt_cta_ri + kkt_cta_prep + kkt_cta
    kkt(24, 8, 6, 3, hess_mode=2)
    kkt(5, 2, 8, 5, mid_p=1, explicit_D2=True)   # kkt_coop (shared-memory workspace)
    h.set_option("kkt_cond_bits", 0)             # force the re-solve path of the tuned kernels
    kkt(12, 4, 10, 5)
    h.set_option("kkt_cond_bits", 12)
if which in ("all", "factor"):
    import lqr_b200 as LQR
    s = LQR.CholeskySolver(problems.random_lqr_kkt(6, 3, 8, 3, seed=3, mid_p=1), handle=h).factor_()
    s.solve_factored_()
    LQR.get_shur_factors(s)
    ran.append(h.last_kernel + " (factor / solve_factored / get_shur)")
if which in ("all", "sqp"):
    Z0, x0, xf, o = problems.dubins_turn90(40, N=21)
    Z0 = Z0 + 0.2 * np.random.default_rng(0).standard_normal(Z0.shape)
    s = DubinsSQP(x0, xf, N=21, tf=3.0, iters=4, handle=h)
    s.solve_(Z0)
    ran.append(h.last_kernel)
h.close()
print("SANITIZE_DRIVER_OK", which, ran)
