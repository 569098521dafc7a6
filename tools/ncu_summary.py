#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[ki][:110])
        for wname in WANT:
            cands = [x for x in hdr if x.endswith(wname)]
            if cands:
                i = hdr.index(cands[0])
                print(f"  {wname:82s} {r[i]:>18s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
            print(f"  traffic (dram read+write) per launch: {tot:.6e} bytes")
        except Exception as e:  # noqa: BLE001
            print("  traffic: n/a", e)


if __name__ == "__main__":
    main()
