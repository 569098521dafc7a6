"""Host->device link probe for the end-to-end (host buffer) call, one process per GPU under torchrun.

Every rank measures, between barriers so that all ranks copy at once:
  1. the plain pinned cudaMemcpyAsync rate of the config-2 A array (839 MB),
  2. the same from write-combined pinned memory (cudaHostAllocWriteCombined),
  3. lqrb_riccati_f64 end to end with pinned host inputs,
  4. the same with write-combined pinned inputs.
Rank 0 prints one JSON line with the per-rank minimum / maximum of each.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from cuda.bindings import runtime as cudart

    from lqr_b200 import _lib, ops, synthetic

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    n, m, N, batch = 4, 1, 101, int(os.environ.get("PROBE_BATCH", 65536))
    names = synthetic.RICCATI_NAMES
    chunks = []
    for ci, first in enumerate(range(0, batch, 16384)):
        f = synthetic.riccati_cartpole_chunk(min(16384, batch - first), 11 + rank, ci, N=N)
        chunks.append({k: f[k].cpu() for k in names})
        del f
    host = {k: torch.cat([c[k] for c in chunks]).pin_memory() for k in names}
    del chunks

    # write-combined pinned copies of the same inputs
    wc_ptrs, wc = [], {}
    for k in names:
        nbytes = host[k].numel() * 8
        err, p = cudart.cudaHostAlloc(nbytes, cudart.cudaHostAllocWriteCombined)
        assert err == cudart.cudaError_t.cudaSuccess, err
        wc_ptrs.append(p)
        dst = np.ctypeslib.as_array((ctypes.c_double * host[k].numel()).from_address(p))
        np.copyto(dst, host[k].numpy().reshape(-1))
        wc[k] = int(p)

    h = _lib.Handle(local)
    NN = _lib.num_vars(n, m, N)
    Zh = torch.empty(batch, NN, dtype=torch.float64).pin_memory()
    infoh = torch.zeros(batch, dtype=torch.int32).pin_memory()
    dev = torch.empty_like(host["A"], device="cuda")
    abytes = host["A"].numel() * 8
    stream = torch.cuda.current_stream().cuda_stream

    def copy_pinned():
        dev.copy_(host["A"], non_blocking=True)

    def copy_wc():
        (err,) = cudart.cudaMemcpyAsync(dev.data_ptr(), wc["A"], abytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice,
                                        stream)
        assert err == cudart.cudaError_t.cudaSuccess, err

    def solve(src):
        ops.riccati(h, n, m, N, batch, 0, *[src[k] for k in names], Zh, None, None, infoh)

    def timed(fn, reps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier()
        return dt

    res = {}
    res["link_pinned_gbs"] = abytes / timed(copy_pinned, 4) / 1e9
    res["link_wc_gbs"] = abytes / timed(copy_wc, 4) / 1e9
    res["e2e_pinned_ms"] = 1e3 * timed(lambda: solve(host), 3)
    ref = Zh[:64].clone()
    res["e2e_wc_ms"] = 1e3 * timed(lambda: solve(wc), 3)
    assert int(infoh.abs().max()) == 0
    assert torch.equal(ref, Zh[:64]), "write-combined inputs changed the result"
    res["link_pinned_gbs_again"] = abytes / timed(copy_pinned, 4) / 1e9

    keys = sorted(res)
    t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = t.cpu().numpy()[None]
    if rank == 0:
        out = {"n_gpus": world, "batch_per_gpu": batch,
               "h2d_bytes_per_call": sum(host[k].numel() * 8 for k in names)}
        for i, k in enumerate(keys):
            out[k] = {"min": float(allv[:, i].min()), "max": float(allv[:, i].max())}
        print(json.dumps(out))
    h.close()
    for p in wc_ptrs:
        cudart.cudaFreeHost(p)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
