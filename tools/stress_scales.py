"""Conditioning stress grid for the tuned large-size KKT kernels (explicit block inverses) — calibration of the
`kkt_cond_bits` threshold of the automatic re-solve (csrc/kkt.cu: kkt_resolve_ill_conditioned).

For every (size, cost rescaling) case prints: the worst log2 pivot ratio the kernel reports, the error against the
extended-precision global KKT solve (oracle/dense_kkt.py) of (a) the tuned kernel alone (kkt_refine = 0), (b) the
default dispatch (automatic re-solve of the flagged instances), (c) the Cholesky-based general kernel, (d) the CPU
oracle — (c) and (d) are the reference's operation order, i.e. the accuracy the reference itself delivers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import oracle  # noqa: E402
from lqr_b200 import _lib, ops, problems  # noqa: E402
from oracle import dense_kkt  # noqa: E402

h = _lib.Handle(0)
DEFAULT_BITS = 6   # LQRB_KKT_COND_BITS in csrc/kkt.cu


def err_vs_truth(prob, dz, lam, insts):
    e = 0.0
    for i in insts:
        zt, lt = dense_kkt.kkt_truth(prob, i)
        e = max(e, np.linalg.norm(dz[i] - zt) / np.linalg.norm(zt), np.linalg.norm(lam[i] - lt) / np.linalg.norm(lt))
    return e


rng = np.random.default_rng(0)
scales = [(1.0, 1.0), (1e3, 1e-3), (1e-3, 1e3), (1e5, 1.0), (1.0, 1e-5), (1e2, 1e-2), (1e-2, 1e2), (1e4, 1e-1)]
scales += [(10.0 ** rng.uniform(-3, 3), 10.0 ** rng.uniform(-3, 3)) for _ in range(6)]
print("n m N | Q-scale R-scale | bits(max) resolved | tuned-only default coop oracle  (rel. err vs refined truth)")
for (n, m, N, b) in [(12, 4, 40, 4), (8, 4, 30, 4), (64, 16, 12, 2), (24, 8, 20, 2), (12, 4, 301, 2)]:
    for t, (qs, rs_) in enumerate(scales):
        p = problems.random_lqr_kkt(n, m, N, b, seed=200 + t, mid_p=0, hess_mode=1)
        p["Q"] = p["Q"] * qs
        p["R"] = p["R"] * rs_
        insts = range(b)
        h.set_option("kkt_cond_bits", 99)  # estimates are read back, nothing is re-solved: the tuned kernel alone
        dz0, lam0, i0 = ops.kkt_solve_problem(p, handle=h)
        bits, _ = h.kkt_last_condition(b)
        h.set_option("kkt_cond_bits", DEFAULT_BITS)
        dz1, lam1, i1 = ops.kkt_solve_problem(p, handle=h)
        k1 = h.last_kernel
        _, nres = h.kkt_last_condition(b)
        h.set_option("kkt_variant", 2)
        dz2, lam2, i2 = ops.kkt_solve_problem(p, handle=h)
        h.set_option("kkt_variant", 0)
        dzo, lamo, _ = oracle.kkt_solve(p)
        e0, e1, e2, eo = (err_vs_truth(p, a, c, insts) for a, c in ((dz0, lam0), (dz1, lam1), (dz2, lam2), (dzo, lamo)))
        flag = "  <-- tuned-only above 1e-10" if e0 > 1e-10 else ""
        print(f"{n} {m} {N} | {qs:.1e} {rs_:.1e} | {bits.max():3d} {nres:3d} | {e0:.2e} {e1:.2e} {e2:.2e} {eo:.2e} "
              f"| info {int(i0.max())} {k1.split('+')[0][:24]}{flag}", flush=True)
