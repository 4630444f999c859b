#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep from `ncu --set full`) as the short markdown table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rNN_<kernel>.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active, % of elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active, % of active"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM/tensor-memory path active, % of elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall: long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct", "stall: barrier"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard_per_warp_active.pct", "stall: short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_wait_per_warp_active.pct", "stall: wait"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall: math pipe throttle"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle_per_warp_active.pct", "stall: mio throttle"),
    ("smsp__average_warp_latency_issue_stalled_membar_per_warp_active.pct", "stall: membar"),
    ("smsp__average_warp_latency_issue_stalled_sleeping_per_warp_active.pct", "stall: sleeping"),
]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and pat not in name:
            continue
        print(f"## {name}  (launch id {r[col['ID']]})\n")
        print("| metric | value | unit |")
        print("|---|---|---|")
        for key, label in WANT:
            if key in col:
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        print()


if __name__ == "__main__":
    main()
