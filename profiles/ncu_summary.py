"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into a small JSON list, one entry per profiled launch.
usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv | python profiles/ncu_summary.py > profiles/rN_X.json"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "gpu__time_duration",
    "dram__bytes_read.sum": "dram__bytes_read",
    "dram__bytes_write.sum": "dram__bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard": "stall_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier": "stall_barrier",
    "smsp__pcsamp_warps_issue_stalled_wait": "stall_wait",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard": "stall_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_no_instructions": "stall_no_instruction",
    "smsp__pcsamp_warps_issue_stalled_selected": "issued_selected",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle": "stall_math_pipe",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle": "stall_mio",
}
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][-70:]}
    for h, u, v in zip(hdr, units, r):
        if h in KEEP:
            d[KEEP[h]] = f"{v} {u}".strip()
    out.append(d)
json.dump(out, sys.stdout, indent=1)
