#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (ncu --set full) into the JSON layout used under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "note ..." > profiles/rNN_ncu_x.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "drain", "lg_throttle", "long_scoreboard", "math_pipe_throttle",
          "membar", "mio_throttle", "no_instruction", "not_selected", "selected", "short_scoreboard", "sleeping", "tex_throttle", "wait"]


def main():
    rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        k = {key: (f"{d[key]} {u[key]}".strip() if key in d else None) for key in KEYS}
        st = {}
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in d:
                try:
                    st[s] = round(float(d[key]), 3)
                except ValueError:
                    pass
        k["stalls_per_issue"] = {a: b for a, b in st.items() if b >= 0.01}
        kernels.append(k)
    print(json.dumps({"note": note, "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
