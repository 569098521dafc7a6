"""Generates tests/golden/*.npz: refined-truth outputs of the global KKT solve (oracle/dense_kkt.py, the
analogue of src/sparse_solver.jl:267-292 / test/cholesky_solve.jl:42) on seeded inputs.

The reference itself keeps no golden vectors and cannot be executed here (Julia absent), so these vectors
are NOT reference outputs; they freeze the refined truth so that later edits to the oracle or kernels are
caught.  Re-generate with:  python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lqr_b200 import problems  # noqa: E402

KKT_CASES = {
    "cartpole_kkt": lambda: problems.cartpole_fixture(),
    "double_integrator_kkt": lambda: problems.double_integrator_fixture(),
    "dubins_kkt": lambda: problems.dubins_kkt_batch(2, seed=1, N=201),
    # config 5 shapes at short horizons (the tuned large-size kernels)
    "quad_kkt": lambda: problems.random_lqr_kkt(12, 4, 41, 2, seed=31, mid_p=0, hess_mode=1),
    "large_kkt": lambda: problems.random_lqr_kkt(64, 16, 8, 2, seed=32, mid_p=0, hess_mode=1),
    # the other CTA sizes, a stage-constrained Dubins problem, a diagonal-Hessian problem, an explicit-D2 problem
    "mid32_kkt": lambda: problems.random_lqr_kkt(32, 8, 14, 2, seed=35, mid_p=0, hess_mode=1),
    "mid24_kkt": lambda: problems.random_lqr_kkt(24, 8, 12, 2, seed=36, mid_p=0, hess_mode=2),
    "dubins_stage_kkt": lambda: problems.dubins_kkt_batch(2, seed=5, N=61, mid_p=1),
    "explicit_d2_kkt": lambda: problems.random_lqr_kkt(5, 2, 12, 2, seed=37, mid_p=1, hess_mode=0, explicit_D2=True),
}

RICCATI_CASES = {
    "cartpole_riccati": lambda: problems.riccati_cartpole_batch(4, seed=0),
    "quad_riccati": lambda: problems.random_lqr_riccati(12, 4, 41, 2, seed=33),
    "large_riccati": lambda: problems.random_lqr_riccati(64, 16, 8, 2, seed=34),
    "mid24_riccati": lambda: problems.random_lqr_riccati(24, 8, 30, 2, seed=38),
    "lti_riccati": lambda: problems.random_lqr_riccati(8, 4, 60, 2, seed=39, lti=True),
}


def main():
    from oracle import dense_kkt
    here = os.path.dirname(os.path.abspath(__file__))
    for name, make in KKT_CASES.items():
        prob = make()
        dz, lam = dense_kkt.kkt_truth(prob, 0)
        np.savez_compressed(os.path.join(here, name + ".npz"), dz=dz, mult=lam)
    for name, make in RICCATI_CASES.items():
        prob = make()
        kp = dense_kkt.riccati_as_kkt(prob)
        n, m, N, b = prob["n"], prob["m"], prob["N"], prob["x0"].shape[0]
        X = np.zeros((b, N, n))
        U = np.zeros((b, N - 1, m))
        for i in range(b):
            zt, _ = dense_kkt.kkt_truth(kp, i)
            body = zt[:(N - 1) * (n + m)].reshape(N - 1, n + m)
            X[i, :-1], U[i], X[i, -1] = body[:, :n], body[:, n:], zt[(N - 1) * (n + m):]
        np.savez_compressed(os.path.join(here, name + ".npz"), X=X, U=U, batch=b, seed=0)


if __name__ == "__main__":
    main()
