#!/usr/bin/env python
"""bench.py — batched LQR KKT solves/sec (FP64) on B200, the BASELINE.json metric.

Workload at every N (weak scaling, per-GPU work fixed): BASELINE.json configs[1] — batched cartpole
LQR n=4 m=1 N=101, 65,536 random LTV-affine instances per GPU, Riccati backward pass + forward rollout
(SURVEY §8d config 2).  A "step" = one pass of the hot path over the whole batch.

  value : solves/s with the packed inputs already resident in HBM (one kernel launch per step),
          CUDA-event timed on the launching stream, max over ranks.
  e2e   : the same metric through the reference-facing C-ABI call lqrb_riccati_f64 with HOST (pinned)
          instance-major buffers: H2D + pack + solve + unpack + D2H inside the timed region.
  roofline / cpu_baseline : see DESIGN.md §Measurement.

  --impl reference : times the reference algorithm's CPU restatement (oracle/, OpenMP over all host
          cores) on a bounded sample of the same workload.  (LQR.jl itself is pure Julia and Julia is
          not installed in this image — see DESIGN.md.)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n, m, N, per-GPU batch)
    "cartpole_riccati_n4_m1_N101_b65536": (4, 1, 101, 65536),
}
DEFAULT_WORKLOAD = "cartpole_riccati_n4_m1_N101_b65536"
METRIC = "batched LQR KKT solves/sec (FP64)"
UNIT = "solves/s"


def algorithmic_bytes_per_solve(n, m, N):
    """unique FP64 words read once + written once per instance (symmetric packed)."""
    tri = lambda k: k * (k + 1) // 2  # noqa: E731
    words_in = (N - 1) * (n * n + n * m + tri(n) + tri(m) + n + m) + tri(n) + 2 * n
    words_out = N * n + (N - 1) * m
    return 8 * (words_in + words_out)


def algorithmic_flops_per_solve(n, m, N):
    return (N - 1) * (4 * n**3 + 8 * n * n * m + 4 * n * m * m + m**3 / 3 + 2 * n * n + 4 * n * m)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ data
def make_host_problem(n, m, N, batch, seed):
    """Seeded synthetic instances in the ABI's instance-major column-major layout (numpy, host)."""
    from lqr_b200 import ops, problems
    assert (n, m, N) == (4, 1, 101)
    prob = problems.riccati_cartpole_batch(batch, seed=seed, N=N)
    return ops.riccati_flatten(prob)


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes read+written per launch of the dominant kernel, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------ CPU leg (oracle port)
def cpu_leg(n, m, N, f, target_seconds=12.0, steps=1, warmup=0):
    """Times the oracle's OpenMP batch driver on a bounded sample.  Returns (solves/s, cores, sample, ms/step)."""
    import oracle
    oracle.build()
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask for them explicitly)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    total = f["x0"].shape[0]

    def run(cnt):
        X = np.zeros((cnt, N, n)); U = np.zeros((cnt, N - 1, m))
        K = np.zeros((cnt, N - 1, n, m)); kff = np.zeros((cnt, N - 1, m))
        info = np.zeros(cnt, dtype=np.int32)
        t0 = time.perf_counter()
        oracle.riccati_raw(n, m, N, 0, cnt, f["A"][:cnt], f["B"][:cnt], f["Q"][:cnt], f["R"][:cnt], f["q"][:cnt],
                           f["r"][:cnt], f["Qf"][:cnt], f["qf"][:cnt], f["x0"][:cnt], X, U, K, kff, info, cores)
        return time.perf_counter() - t0
    probe = min(total, 2048)
    run(probe)
    t_probe = run(probe)
    per_step = max(1, steps + warmup)
    cnt = int(min(total, max(probe, probe * (target_seconds / per_step) / max(t_probe, 1e-6))))
    for _ in range(warmup):
        run(cnt)
    times = [run(cnt) for _ in range(max(1, steps))]
    t = sum(times) / len(times)
    return cnt / t, cores, f"{cnt} of {total} instances per step, {len(times)} step(s)", 1e3 * t


def sparse_leg(n, m, N, seed, seconds=3.0, max_inst=256):
    """The reference's other solver on the same instances: SparseSolver assembles the global KKT pieces and calls
    a sparse factorisation (src/sparse_solver.jl:267-292).  Analogue here: scipy's sparse LU of [H D'; D 0] per
    instance (assembly excluded), one core, bounded sample.  Reported next to the block-recursion port."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from lqr_b200 import problems
    from oracle import dense_kkt
    prob = dense_kkt.riccati_as_kkt(problems.riccati_cartpole_batch(max_inst, seed=seed, N=N))
    t_solve, cnt = 0.0, 0
    while cnt < max_inst and t_solve < seconds:
        H, g, D, d = dense_kkt.assemble(prob, cnt)
        K = sp.bmat([[H, D.T], [D, None]], format="csc")
        rhs = -np.concatenate([g, d])
        t0 = time.perf_counter()
        spla.splu(K).solve(rhs)
        t_solve += time.perf_counter() - t0
        cnt += 1
    return {"value": cnt / t_solve, "unit": UNIT, "cores": 1, "sample": f"{cnt} instances",
            "what": "scipy sparse LU of the assembled KKT system (analogue of src/sparse_solver.jl:267-292)"}


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug only)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: re-launch ourselves under torchrun, one rank per GPU
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    n, m, N, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    config = {"workload": args.workload, "n": n, "m": m, "N": N, "batch_per_gpu": batch,
              "global_batch": batch * max(1, args.gpus), "problem": "LTV affine LQR, Riccati backward pass + forward rollout",
              "l2": "inputs (2.1 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
              "parallelism": f"batch slices over {args.gpus} GPU(s), no collective on the data path"}

    # ------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sample_total = min(batch, 16384)
        f = make_host_problem(n, m, N, sample_total, seed=0)
        v, cores, sample, ms = cpu_leg(n, m, N, f, target_seconds=60.0, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "reference algorithm, C/OpenMP restatement (oracle/lqr_oracle.c); LQR.jl is pure Julia and "
                        "Julia is not installed in this image"}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from lqr_b200 import _lib, ops

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = _lib.Handle(local_rank)
    # a real (non-default) stream: the handle launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)

    f = make_host_problem(n, m, N, batch, seed=rank)
    L = _lib.riccati_layout(n, m, N)
    ldb = _lib.padded_batch(batch)
    names = ("A", "B", "Q", "R", "q", "r", "Qf", "qf", "x0")
    host = {k: torch.from_numpy(f[k]).pin_memory() for k in names}
    dev = {k: host[k].cuda(non_blocking=True) for k in names}
    knots = torch.empty(ldb * L.knot_count * L.rows_per_knot, dtype=torch.float64, device="cuda")
    term = torch.empty(ldb * L.term_rows, dtype=torch.float64, device="cuda")
    Zp = torch.empty(ldb * L.z_rows, dtype=torch.float64, device="cuda")
    gains = torch.empty(ldb * L.gain_rows, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ops.riccati_pack(h, n, m, N, batch, 0, *[dev[k] for k in names], knots, term)
    torch.cuda.synchronize()
    del dev

    def step():
        ops.riccati_solve_packed(h, n, m, N, batch, 0, knots, term, Zp, gains, info)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = h.launches
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    launches = h.launches - launches0
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    assert int(info.abs().max().item()) == 0, "numerical failure flagged in info[]"
    kernel_name = h.last_kernel

    # ---- end to end through the C ABI with host buffers
    NN = L.z_rows
    Zh = torch.empty(batch, NN, dtype=torch.float64).pin_memory()
    infoh = torch.zeros(batch, dtype=torch.int32).pin_memory()

    def e2e_step():
        ops.riccati(h, n, m, N, batch, 0, *[host[k] for k in names], Zh, None, None, infoh)

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    h2d_bytes = sum(host[k].numel() * 8 for k in names)
    d2h_bytes = Zh.numel() * 8 + infoh.numel() * 4

    # ---- parity spot check of what the timed kernel produced (device path == host path, bitwise)
    Zd = torch.empty(batch, NN, dtype=torch.float64, device="cuda")
    ops.unpack_rows(h, NN, batch, Zp, Zd)
    torch.cuda.synchronize()
    same = bool(torch.equal(Zd[:256].cpu(), Zh[:256]))

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t[0]), float(t[1])
    ms_per_step = total_ms_max / args.steps
    value = world * batch / (ms_per_step * 1e-3)
    e2e_value = world * batch / e2e_s_max

    if rank == 0:
        bytes_per = algorithmic_bytes_per_solve(n, m, N)
        kern_ms = statistics.mean(per_launch_ms)
        achieved = bytes_per * batch / (kern_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak_hbm()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e2e_s_max,
                    "api": "lqrb_riccati_f64 (host pinned instance-major buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload), "kernel": kernel_name,
                         "peak_source": peak_src, "algorithmic_bytes_per_solve": bytes_per,
                         "algorithmic_flops_per_solve": algorithmic_flops_per_solve(n, m, N),
                         "kernel_ms": kern_ms},
            "parity_spot_check": "device-resident result == host-path result (bitwise)" if same else "MISMATCH",
        }
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample, _ = cpu_leg(n, m, N, f, target_seconds=12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "sparse_kkt": sparse_leg(n, m, N, seed=0)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


if __name__ == "__main__":
    main()
