import sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
import oracle
from lqr_b200 import _lib, ops, problems
h = _lib.Handle(0)
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
rng = np.random.default_rng(0)
worst = {}
for trial in range(12):
    qs = 10.0 ** rng.uniform(-3, 3); rs_ = 10.0 ** rng.uniform(-3, 3)
    for (n, m, N, b) in [(12, 4, 60, 4), (8, 4, 40, 4), (64, 16, 12, 2), (24, 8, 20, 2)]:
        p = problems.random_lqr_riccati(n, m, N, b, seed=100 + trial)
        p["Q"] = p["Q"] * qs; p["R"] = p["R"] * rs_; p["Qf"] = p["Qf"] * qs
        X, U, K, kff, info = ops.riccati_solve_problem(p, handle=h)
        k1 = h.last_kernel
        Xo, Uo, Ko, kffo, _ = oracle.riccati(p)
        e = max(rel(X, Xo), rel(U, Uo), rel(K, Ko), rel(kff, kffo))
        h.set_option("riccati_variant", 2)
        X2, U2, K2, kff2, _ = ops.riccati_solve_problem(p, handle=h)
        h.set_option("riccati_variant", 0)
        e2 = max(rel(X2, Xo), rel(U2, Uo), rel(K2, Ko), rel(kff2, kffo))
        key = ("ric", n, m)
        worst[key] = max(worst.get(key, (0, 0)), (e, e2), key=lambda t: t[0])
        if e > 1e-10: print("RICCATI", n, m, "scale", qs, rs_, "err tuned", e, "coop", e2, k1, info.max())
    for (n, m, N, b) in [(12, 4, 40, 4), (8, 4, 30, 4), (64, 16, 12, 2)]:
        p = problems.random_lqr_kkt(n, m, N, b, seed=200 + trial, mid_p=0, hess_mode=1)
        p["Q"] = p["Q"] * qs; p["R"] = p["R"] * rs_
        dz, lam, info = ops.kkt_solve_problem(p, handle=h)
        k1 = h.last_kernel
        dzo, lamo, _ = oracle.kkt_solve(p)
        e = max(rel(dz, dzo), rel(lam, lamo))
        h.set_option("kkt_variant", 2)
        dz2, lam2, _ = ops.kkt_solve_problem(p, handle=h)
        h.set_option("kkt_variant", 0)
        e2 = max(rel(dz2, dzo), rel(lam2, lamo))
        key = ("kkt", n, m)
        worst[key] = max(worst.get(key, (0, 0)), (e, e2), key=lambda t: t[0])
        if e > 1e-10: print("KKT", n, m, "scale %.2e %.2e" % (qs, rs_), "err tuned %.2e coop %.2e" % (e, e2), k1, info.max())
for k, v in worst.items(): print(k, "worst tuned %.2e (coop on the same case %.2e)" % v)
