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
    # round 2: stage rows on the tensor-core kernels, a size embedded in the next tuned class, a free final state
    "quad_stage_kkt": lambda: problems.random_lqr_kkt(12, 4, 30, 2, seed=41, mid_p=2, hess_mode=0),
    "large_stage_kkt": lambda: problems.random_lqr_kkt(64, 16, 10, 2, seed=42, mid_p=1, hess_mode=1),
    "arm_padded_kkt": lambda: problems.random_lqr_kkt(14, 7, 14, 2, seed=43, mid_p=1, hess_mode=1),
    "free_final_kkt": lambda: _free_final(problems.random_lqr_kkt(12, 4, 25, 2, seed=44, mid_p=1, hess_mode=1)),
}


def _free_final(prob):
    """No goal rows: p_N = 0."""
    prob["p"] = prob["p"].copy()
    prob["p"][-1] = 0
    b = prob["q"].shape[0]
    prob["C"][-1] = np.zeros((b, 0, prob["n"]))
    prob["c"][-1] = np.zeros((b, 0))
    return prob


RICCATI_CASES = {
    "cartpole_riccati": lambda: problems.riccati_cartpole_batch(4, seed=0),
    "quad_riccati": lambda: problems.random_lqr_riccati(12, 4, 41, 2, seed=33),
    "large_riccati": lambda: problems.random_lqr_riccati(64, 16, 8, 2, seed=34),
    "mid24_riccati": lambda: problems.random_lqr_riccati(24, 8, 30, 2, seed=38),
    "lti_riccati": lambda: problems.random_lqr_riccati(8, 4, 60, 2, seed=39, lti=True),
    "padded_riccati": lambda: problems.random_lqr_riccati(10, 3, 40, 2, seed=45),
}


def main():
    from oracle import dense_kkt
    here = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])  # optional: the names to (re)generate; default all
    for name, make in KKT_CASES.items():
        if only and name not in only:
            continue
        prob = make()
        dz, lam = dense_kkt.kkt_truth(prob, 0)
        np.savez_compressed(os.path.join(here, name + ".npz"), dz=dz, mult=lam)
    for name, make in RICCATI_CASES.items():
        if only and name not in only:
            continue
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
