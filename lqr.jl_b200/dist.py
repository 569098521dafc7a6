"""Multi-GPU plumbing: batch slices, no collective on the data path.

Instances are independent (nothing in src/cholesky_solver.jl:166-182 or src/dynamic_programming.jl:54-72
couples them), so G GPUs = G contiguous batch slices.  torch.distributed is used only for the timing
protocol of bench.py (barrier + max over ranks); NCCL never touches problem data.
"""
from __future__ import annotations

import threading

import numpy as np


def batch_slice(batch: int, rank: int, world: int, align: int = 32):
    """Contiguous slice [lo, hi) of a global batch for `rank`; interior boundaries are multiples of
    `align` (the packed tile width) so every rank packs whole tiles."""
    tiles = (batch + align - 1) // align
    lo_t = tiles * rank // world
    hi_t = tiles * (rank + 1) // world
    return min(lo_t * align, batch), min(hi_t * align, batch)


def max_over_ranks(values, device=None):
    """Elementwise max of a list of floats over all ranks (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]


def aggregate_throughput(units_this_rank: float, ms_this_rank: float, device=None) -> float:
    """Whole-job units/s = (units all ranks processed) / (max over ranks of the time)."""
    total = sum_over_ranks([units_this_rank], device)[0]
    ms = max_over_ranks([ms_this_rank], device)[0]
    return total / (ms * 1e-3)


def solve_sharded(solve_slice, batch: int, devices):
    """In-process variant (one host thread + one handle per device): calls
    `solve_slice(device, lo, hi)` for each device's slice concurrently and returns the results in
    device order.  ctypes releases the GIL during the C-ABI call, so the slices overlap."""
    results = [None] * len(devices)
    errors = []

    def work(i, dev):
        lo, hi = batch_slice(batch, i, len(devices))
        try:
            results[i] = solve_slice(dev, lo, hi)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=work, args=(i, d)) for i, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def riccati_multi_gpu(prob: dict, devices):
    """Riccati solve of one host problem dict over several GPUs by batch slice."""
    from . import _lib, ops
    f = ops.riccati_flatten(prob)
    b = f["batch"]

    def solve(dev, lo, hi):
        h = _lib.Handle(dev)
        try:
            sub = dict(f)
            for k in ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0"):
                sub[k] = None if f[k] is None else f[k][lo:hi]
            sub["batch"] = hi - lo
            return ops.riccati_solve_problem(sub, handle=h)
        finally:
            h.close()
    parts = solve_sharded(solve, b, list(devices))
    return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(5))
