#!/usr/bin/env python
"""Kernel-time probe for the secondary configs (3, 5a, 5b): packs a small seeded base batch on the host,
replicates its tiles on the device to the configured batch, times the packed solve with CUDA events and
prints solves/s plus the algorithmic-bytes / flops rooflines of SURVEY §8d.  Not the headline bench."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqr_b200 import _lib, ops, problems  # noqa: E402


def tri(k):
    return k * (k + 1) // 2


def time_it(fn, steps, warmup, stream):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def probe_riccati(h, stream, n, m, N, batch, base, steps, warmup, peak_gbs):
    prob = problems.random_lqr_riccati(n, m, N, base, seed=3) if (n, m) != (4, 1) else \
        problems.riccati_cartpole_batch(base, seed=0, N=N)
    f = ops.riccati_flatten(prob)
    L = _lib.riccati_layout(n, m, N)
    names = ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0")
    dev = {k: torch.from_numpy(f[k]).cuda() for k in names}
    knots = torch.empty(base * L.knot_count * L.rows_per_knot, dtype=torch.float64, device="cuda")
    term = torch.empty(base * L.term_rows, dtype=torch.float64, device="cuda")
    ops.riccati_pack(h, n, m, N, base, 0, *[dev[k] for k in names], knots, term)
    torch.cuda.synchronize()
    reps = batch // base
    knots, term = knots.repeat(reps), term.repeat(reps)
    Z = torch.empty(batch * L.z_rows, dtype=torch.float64, device="cuda")
    gains = torch.empty(batch * L.gain_rows, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ms = time_it(lambda: ops.riccati_solve_packed(h, n, m, N, batch, 0, knots, term, Z, gains, info), steps, warmup, stream)
    assert int(info.abs().max()) == 0
    by = 8 * ((N - 1) * (n * n + n * m + tri(n) + tri(m) + n + m) + tri(n) + 2 * n + N * n + (N - 1) * m)
    fl = (N - 1) * (4 * n**3 + 8 * n * n * m + 4 * n * m * m + m**3 / 3 + 2 * n * n + 4 * n * m)
    return dict(kind="riccati", n=n, m=m, N=N, batch=batch, kernel=h.last_kernel, ms=ms, solves_per_s=batch / ms * 1e3,
                alg_GBs=by * batch / ms / 1e6, hbm_frac=by * batch / ms / 1e6 / peak_gbs, alg_TFLOPs=fl * batch / ms / 1e9)


def probe_kkt(h, stream, n, m, N, batch, base, steps, warmup, peak_gbs, dubins=False, mid_p=0):
    prob = problems.dubins_kkt_batch(base, seed=1, N=N, mid_p=mid_p) if dubins else \
        problems.random_lqr_kkt(n, m, N, base, seed=3, mid_p=mid_p)
    f = ops.kkt_flatten(prob)
    p, hess = f["p"], f["hess_mode"]
    rows = _lib.kkt_data_rows(n, m, N, p, hess, False)
    NN, P = _lib.num_vars(n, m, N), _lib.num_cons(n, N, p)
    names = ("Q", "R", "Hux", "q", "r", "A", "B", "d", "D2", "C", "c")
    dev = [None if f[k] is None else torch.from_numpy(f[k]).cuda() for k in names]
    data = torch.empty(base * rows, dtype=torch.float64, device="cuda")
    ops.kkt_pack(h, n, m, N, base, p, hess, *dev, data)
    torch.cuda.synchronize()
    data = data.repeat(batch // base)
    dz = torch.empty(batch * NN, dtype=torch.float64, device="cuda")
    mult = torch.empty(batch * P, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ms = time_it(lambda: ops.kkt_solve_packed(h, n, m, N, batch, p, hess, False, 0, data, dz, mult, None, info), steps, warmup, stream)
    assert int(info.abs().max()) == 0
    by = 8 * (rows + NN + P)
    return dict(kind="kkt", n=n, m=m, N=N, batch=batch, kernel=h.last_kernel, ms=ms, solves_per_s=batch / ms * 1e3,
                alg_GBs=by * batch / ms / 1e6, hbm_frac=by * batch / ms / 1e6 / peak_gbs, bytes_per_solve=by)


def probe_sqp(h, stream, batch, N, steps, warmup, line_search):
    """config 4: batched Dubins SQP, 10 outer iterations, device-resident iterates; metric = KKT solves/s."""
    Z0, x0, xf, o = problems.dubins_turn90(256, N=N)
    reps = batch // 256
    x0d = torch.from_numpy(np.tile(x0, (reps, 1))).cuda()
    xfd = torch.from_numpy(np.tile(xf, (reps, 1))).cuda()
    Z0d = torch.from_numpy(np.tile(np.broadcast_to(Z0, (256, Z0.shape[-1])), (reps, 1))).cuda()
    Zd = Z0d.clone()
    fp = torch.zeros(batch, dtype=torch.float64, device="cuda")
    fd = torch.zeros(batch, dtype=torch.float64, device="cuda")
    it = torch.zeros(batch, dtype=torch.int32, device="cuda")
    o = dict(o, iters=10, line_search=int(line_search))
    solves = [0]

    def run():
        Zd.copy_(Z0d)
        solves[0] = ops.sqp_dubins(h, batch, o, x0d, xfd, Zd, fp, fd, it)

    ms = time_it(run, steps, warmup, stream)
    by = 64384 if N == 201 else None
    return dict(kind="sqp", n=3, m=2, N=N, batch=batch, kernel=h.last_kernel, ms=ms, kkt_solves=solves[0],
                solves_per_s=solves[0] / ms * 1e3, line_search=int(line_search),
                converged=float(((fp < 1e-5) & (fd < 2e-5)).double().mean()),
                hbm_frac=(by * solves[0] / ms / 1e6 / json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if by else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c2,c3,c4,5aR,5aK,5bR,5bK")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batches (debug)")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to lqrb_set_option (kernel A/B)")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    h = _lib.Handle(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    for kv in args.opt:
        name, val = kv.split("=")
        h.set_option(name, int(val))
    S = args.scale
    out = []
    for w in args.which.split(","):
        if w == "c2":
            r = probe_riccati(h, stream, 4, 1, 101, int(65536 * S), 256, args.steps, args.warmup, peak)
        elif w == "c3":
            r = probe_kkt(h, stream, 3, 2, 201, int(262144 * S), 256, args.steps, args.warmup, peak, dubins=True)
        elif w == "c3p1":
            r = probe_kkt(h, stream, 3, 2, 201, int(262144 * S), 256, args.steps, args.warmup, peak, dubins=True, mid_p=1)
        elif w == "c1b":
            r = probe_kkt(h, stream, 4, 1, 101, int(65536 * S), 256, args.steps, args.warmup, peak)
        elif w == "c4":
            r = probe_sqp(h, stream, int(65536 * S), 201, args.steps, args.warmup, True)
        elif w == "c4fs":
            r = probe_sqp(h, stream, int(65536 * S), 201, args.steps, args.warmup, False)
        elif w == "5aR":
            r = probe_riccati(h, stream, 12, 4, 1001, int(16384 * S), 32, args.steps, args.warmup, peak)
        elif w == "5aK":
            r = probe_kkt(h, stream, 12, 4, 1001, int(16384 * S), 32, args.steps, args.warmup, peak)
        elif w == "5bR":
            r = probe_riccati(h, stream, 64, 16, 101, int(4096 * S), 32, args.steps, args.warmup, peak)
        elif w == "5bK":
            r = probe_kkt(h, stream, 64, 16, 101, int(4096 * S), 32, args.steps, args.warmup, peak)
        elif w in ("5bKp1", "5bKp2", "5bKp4"):  # 5b-K with 1 / 2 / 4 stage rows on every interior knot
            r = probe_kkt(h, stream, 64, 16, 101, int(4096 * S), 32, args.steps, args.warmup, peak, mid_p=int(w[-1]))
        elif w in ("5aKp1", "5aKp3"):
            r = probe_kkt(h, stream, 12, 4, 1001, int(16384 * S), 32, args.steps, args.warmup, peak, mid_p=int(w[-1]))
        elif w.startswith("kkt:"):  # kkt:n:m:N:batch[:mid_p]
            f = [int(x) for x in w.split(":")[1:]]
            r = probe_kkt(h, stream, f[0], f[1], f[2], f[3], 32, args.steps, args.warmup, peak, mid_p=f[4] if len(f) > 4 else 0)
        elif w.startswith("ric:"):  # ric:n:m:N:batch
            f = [int(x) for x in w.split(":")[1:]]
            r = probe_riccati(h, stream, f[0], f[1], f[2], f[3], 32, args.steps, args.warmup, peak)
        else:
            continue
        r["config"] = w
        print(json.dumps(r), flush=True)
        out.append(r)
        torch.cuda.empty_cache()
    h.close()


if __name__ == "__main__":
    main()
