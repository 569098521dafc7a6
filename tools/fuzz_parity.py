"""Randomised parity sweep of the whole dispatch (run on a GPU box): random sizes, horizons, batch widths,
per-knot stage-row patterns, Hessian modes, explicit D2, SOC / LTI flags through the C ABI against the oracle
and the refined dense KKT truth.  Prints one line per failing case and a summary; exit code 1 on any failure.

    python tools/fuzz_parity.py --cases 400 --seed 1

The rule for a failure (the test suite's rule, tests/test_gpu_kkt.py): relative error against the oracle above
1e-10 AND more than 20 x the oracle's own error against the refined truth (random shapes can be ill-conditioned;
then both implementations lose the same digits), or info[] disagreeing with the oracle's."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def random_kkt(rng, n, m, N, batch, hess_mode, explicit_d2, pattern):
    """Like problems.random_lqr_kkt but with a per-knot stage-row pattern p[k]."""
    from lqr_b200 import problems
    base = problems.random_lqr_kkt(n, m, N, batch, seed=int(rng.integers(1 << 30)), hess_mode=hess_mode,
                                   explicit_D2=explicit_d2)
    p = np.asarray(pattern, dtype=np.int32)
    Cs, cs = [], []
    for k in range(N):
        w = n + (m if k < N - 1 else 0)
        if k == 0 and p[k] == n:
            C = np.zeros((batch, n, w))
            C[:, :, :n] = np.eye(n)
        elif k == N - 1 and p[k] == n:
            C = np.broadcast_to(np.eye(n), (batch, n, n)).copy()
        else:
            C = rng.standard_normal((batch, int(p[k]), w))
        Cs.append(C)
        cs.append(0.1 * rng.standard_normal((batch, int(p[k]))))
    base["p"], base["C"], base["c"] = p, Cs, cs
    return base


def sample_pattern(rng, n, m, N):
    """A well-posed pattern: interior rows <= m - 1 (one free control per knot when m > 1), terminal rows only as
    many as the remaining degrees of freedom can reach."""
    kind = int(rng.integers(0, 6))
    p = np.zeros(N, dtype=np.int32)
    p[0] = n if kind != 5 else int(rng.integers(0, n + 1))
    hi = max(0, m - 1)
    if kind == 1 and hi:
        p[1:N - 1] = int(rng.integers(1, hi + 1))
    elif kind in (2, 5) and hi:
        p[1:N - 1] = rng.integers(0, hi + 1, size=max(0, N - 2))
    elif kind == 3 and hi and N > 3:
        p[int(rng.integers(1, N - 1))] = int(rng.integers(1, hi + 1))
    free = (N - 1) * m - int(p[1:N - 1].sum()) + (n - int(p[0]))
    term = n if kind != 4 else int(rng.integers(0, n + 1))
    p[N - 1] = max(0, min(term, free - 1, n))
    return p


KKT_SIZES = [(1, 1), (2, 1), (3, 2), (4, 1), (4, 2), (5, 2), (6, 3), (7, 2), (8, 1), (8, 2), (8, 3), (8, 4), (9, 3),
             (10, 4), (12, 1), (12, 2), (12, 3), (12, 4), (12, 5), (13, 4), (16, 8), (16, 4), (24, 8), (24, 16),
             (32, 8), (32, 16), (40, 8), (48, 16), (64, 16), (64, 8), (20, 6)]
RIC_SIZES = [(1, 1), (2, 1), (3, 2), (4, 1), (4, 2), (6, 3), (7, 3), (2, 2), (3, 1), (3, 3), (4, 3), (5, 1), (5, 2), (5, 3), (6, 1), (6, 2), (8, 2), (8, 4), (9, 2), (12, 4), (12, 3),
             (12, 1), (13, 4), (16, 8), (16, 16), (24, 8), (32, 16), (40, 8), (48, 16), (64, 16), (64, 8), (20, 5), (70, 9)]


# shapes with a tuned kernel (csrc/kkt.cu: KKT_TPI_SIZES, KKT_HW_SIZES, KKT_WP_SIZES, KKT_CTA_SIZES): (n, m, interior rows)
TUNED = [(4, 1, 0), (3, 2, 0), (3, 2, 1), (2, 1, 0), (4, 2, 1), (6, 3, 1), (4, 2, 0), (6, 3, 0), (2, 2, 0), (2, 2, 1), (3, 1, 0),
         (3, 3, 0), (3, 3, 1), (5, 1, 0), (6, 1, 0), (5, 2, 0), (5, 2, 1), (6, 2, 0), (6, 2, 1), (4, 3, 0), (4, 3, 1), (5, 3, 0), (5, 3, 1),
         (12, 4, 0), (8, 4, 0), (12, 4, 1), (12, 4, 2), (12, 4, 3), (8, 4, 1), (8, 4, 3), (12, 3, 0), (12, 3, 1),
         (12, 3, 2), (8, 3, 0), (8, 3, 2), (12, 2, 0), (12, 2, 1), (8, 2, 0), (8, 2, 1), (12, 1, 0), (8, 1, 0),
         (64, 16, 0), (48, 16, 0), (32, 8, 0), (24, 8, 0), (16, 8, 0), (64, 16, 1), (48, 16, 3), (32, 8, 2), (24, 8, 1),
         (16, 8, 4), (16, 8, 1), (16, 16, 0), (24, 16, 2), (32, 16, 1)]


def draw_kkt_case(rng, case):
    """(description, problem dict) of one random KKT case; no GPU work."""
    if rng.integers(0, 3) > 0:
        n, m, pm = TUNED[int(rng.integers(len(TUNED)))]
        big = n >= 16
        lo = 2 + -(-n // max(1, m - pm))  # enough knots for the goal to be reachable
        N = int(lo + rng.choice([0, 1, 2, 5, 11, 30, 64] if not big else [0, 1, 3, 6]))
        batch = int(rng.choice([1, 2, 5, 16, 31, 32, 33, 47, 64, 70, 150] if not big else [1, 2, 3, 5, 9, 20]))
        hess = int(rng.integers(0, 3))
        d2x = False
        soc = bool(rng.integers(0, 6) == 0)
        p = np.full(N, pm, dtype=np.int32)
        p[0] = p[-1] = n
        if rng.integers(0, 4) == 0:
            p[-1] = 0  # free final state
    else:
        n, m = KKT_SIZES[int(rng.integers(len(KKT_SIZES)))]
        big = n >= 16
        N = int(rng.choice([2, 3, 4, 5, 7, 9, 12, 17, 30, 41] if not big else [2, 3, 5, 8, 11]))
        batch = int(rng.choice([1, 2, 5, 16, 31, 32, 33, 47, 64, 70] if not big else [1, 2, 3, 5, 9]))
        hess = int(rng.integers(0, 3))
        d2x = bool(rng.integers(0, 5) == 0)
        soc = bool(rng.integers(0, 6) == 0)
        p = sample_pattern(rng, n, m, N)
    desc = dict(case=case, kind="kkt", n=n, m=m, N=N, batch=batch, hess=hess, d2x=d2x, soc=soc, p=p.tolist())
    return desc, random_kkt(rng, n, m, N, batch, hess, d2x, p)


def run_kkt_case(rng, h, oracle, dense_kkt, ops, case):
    desc, prob = draw_kkt_case(rng, case)
    soc = desc["soc"]
    try:
        dz, lam, info, res = ops.kkt_solve_problem(prob, soc=soc, want_res=True, handle=h)
    except Exception as e:  # noqa: BLE001
        desc.update(fail=f"exception {e!r}")
        return desc
    desc["kernel"] = h.last_kernel
    dzo, lamo, infoo, reso = oracle.kkt_solve(prob, soc=soc, want_res=True)
    if ((info != 0) != (infoo != 0)).any():
        # ill-posed draw (rank-deficient constraints): both must say so
        desc.update(fail=f"info mismatch cuda {info[:4].tolist()} oracle {infoo[:4].tolist()}")
        return desc
    ok = (infoo == 0)
    if not ok.any():
        desc["skipped"] = "ill-posed for both"
        return desc
    worst = 0.0
    for i in np.flatnonzero(ok):
        e = max(rel(dz[i], dzo[i]), rel(lam[i], lamo[i]) if lamo.shape[1] else 0.0)
        er = float(np.linalg.norm(res[i] - reso[i]) / max(1.0, np.linalg.norm(reso[i])))
        worst = max(worst, e, er)
    desc["err"] = worst
    if worst > 1e-10:
        i = int(np.flatnonzero(ok)[0])
        zt, lt = dense_kkt.kkt_truth(prob, i, soc=soc)
        eo = max(rel(dzo[i], zt), rel(lamo[i], lt) if lt.size else 0.0)
        ec = max(rel(dz[i], zt), rel(lam[i], lt) if lt.size else 0.0)
        desc.update(err_oracle_truth=eo, err_cuda_truth=ec)
        worst_i = max(np.flatnonzero(ok), key=lambda j: rel(dz[j], dzo[j]))
        zt, lt = dense_kkt.kkt_truth(prob, int(worst_i), soc=soc)
        eo2 = max(rel(dzo[worst_i], zt), rel(lamo[worst_i], lt) if lt.size else 0.0)
        ec2 = max(rel(dz[worst_i], zt), rel(lam[worst_i], lt) if lt.size else 0.0)
        desc.update(worst_instance=int(worst_i), worst_oracle_truth=eo2, worst_cuda_truth=ec2)
        if ec2 > max(1e-10, 20 * eo2) or ec > max(1e-10, 20 * eo):
            # an ill-conditioned draw can still explain it: a backward-stable solve of S lam = h loses cond(S) eps
            H, g, D, d = dense_kkt.assemble(prob, int(worst_i), soc=soc)[:4]
            H, D = np.asarray(H.todense() if hasattr(H, "todense") else H), np.asarray(D.todense() if hasattr(D, "todense") else D)
            cS = float(np.linalg.cond(D @ np.linalg.solve(H, D.T)))
            desc["cond_S"] = cS
            if max(ec, ec2) > 10 * 2.2e-16 * cS:
                desc["fail"] = "error above 1e-10, above 20x the oracle's own and above 10 eps cond(S)"
    return desc


def draw_riccati_case(rng, problems, case):
    n, m = RIC_SIZES[int(rng.integers(len(RIC_SIZES)))]
    big = n >= 16
    N = int(rng.choice([2, 3, 4, 6, 11, 25, 60, 101] if not big else [2, 3, 6, 13, 30]))
    batch = int(rng.choice([1, 2, 7, 31, 32, 33, 64, 65, 130] if not big else [1, 2, 3, 6, 10]))
    lti = bool(rng.integers(0, 4) == 0)
    desc = dict(case=case, kind="riccati", n=n, m=m, N=N, batch=batch, lti=lti)
    return desc, problems.random_lqr_riccati(n, m, N, batch, seed=int(rng.integers(1 << 30)), lti=lti)


def run_riccati_case(rng, h, oracle, ops, problems, case):
    desc, prob = draw_riccati_case(rng, problems, case)
    batch = desc["batch"]
    try:
        X, U, K, kff, info = ops.riccati_solve_problem(prob, handle=h)
    except Exception as e:  # noqa: BLE001
        desc.update(fail=f"exception {e!r}")
        return desc
    desc["kernel"] = h.last_kernel
    Xo, Uo, Ko, kffo, infoo = oracle.riccati(prob)
    if (info != 0).any() or (infoo != 0).any():
        desc.update(fail=f"info cuda {info[:4].tolist()} oracle {infoo[:4].tolist()}")
        return desc
    e = max(max(rel(X[i], Xo[i]), rel(U[i], Uo[i]), rel(K[i], Ko[i]), rel(kff[i], kffo[i])) for i in range(batch))
    desc["err"] = e
    if e > 1e-10:
        desc["fail"] = "error above 1e-10"
    return desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--budget-s", type=float, default=900.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import oracle
    from oracle import dense_kkt
    from lqr_b200 import _lib, ops, problems
    h = _lib.Handle(0)
    t0 = time.time()
    fails, done, kernels, worst = [], 0, {}, {}
    out = open(args.out, "w") if args.out else None
    for case in range(args.cases):
        if time.time() - t0 > args.budget_s:
            break
        rng = np.random.default_rng([args.seed, case])  # every case reproducible on its own
        if rng.integers(0, 4) == 0:
            d = run_riccati_case(rng, h, oracle, ops, problems, case)
        else:
            d = run_kkt_case(rng, h, oracle, dense_kkt, ops, case)
        done += 1
        k = d.get("kernel", "?").split("<")[0]
        kernels[k] = kernels.get(k, 0) + 1
        if "err" in d:
            worst[k] = max(worst.get(k, 0.0), d["err"])
        if out:
            out.write(json.dumps(d) + "\n")
        if "fail" in d:
            fails.append(d)
            print("FAIL", json.dumps(d), flush=True)
            if "exception" in d["fail"]:
                # a sticky CUDA error poisons the context: start over with a new handle if possible
                try:
                    h.close()
                except Exception:  # noqa: BLE001
                    pass
                try:
                    h = _lib.Handle(0)
                except Exception as e:  # noqa: BLE001
                    print("cannot recreate handle:", e)
                    break
    print(json.dumps(dict(cases=done, failures=len(fails), kernels=kernels, worst_err_by_kernel=worst,
                          seconds=round(time.time() - t0, 1))))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
