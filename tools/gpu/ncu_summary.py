#!/usr/bin/env python
"""Turns `ncu -i <rep> --page raw --csv` output into a markdown table per launch and profiles/traffic.json
(dram bytes per launch of each kernel, read by bench.py for roofline.traffic).
usage: python tools/gpu/ncu_summary.py <raw.csv> <out.md> [traffic.json]"""
import csv
import json
import sys

METRICS = [
    ("time ms", "gpu__time_duration.sum"),
    ("grid", "Grid Size"), ("block", "Block Size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("threads per warp instruction (32 = no divergence)", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("integer/ALU pipe % of peak", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("shared wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("stall long_scoreboard (per issue)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall no_instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall branch_resolving", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    traffic = {}
    for n, r in enumerate(data):
        name = r[col["Kernel Name"]].split("(")[0]
        out.append("### launch %d: `%s`\n" % (n, name))
        out.append("| metric | value |\n|---|---|")
        for label, m in METRICS:
            if m not in col:
                continue
            v = r[col[m]]
            if v in ("", "nan", "-nan"):
                v = "n/a (not collected in this pass)"
            out.append("| %s | %s %s |" % (label, v, units[col[m]]))
        out.append("")
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
            ms = float(r[col["gpu__time_duration.sum"]])
            if rd == rd and wr == wr:
                traffic.setdefault(name, []).append({"dram_bytes": rd + wr, "ms": ms, "grid": r[col["Grid Size"]]})
        except Exception:
            pass
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
    if len(sys.argv) > 3:
        tj = {}
        for name, lst in traffic.items():
            # the launch that took longest is the one bench.py's dominant-kernel line refers to
            best = max(lst, key=lambda x: x["ms"])
            tj[name] = {"dram_bytes_per_launch": best["dram_bytes"], "ms_under_ncu": best["ms"], "grid": best["grid"],
                        "launches_with_counters": len(lst), "source": sys.argv[1]}
        json.dump(tj, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
