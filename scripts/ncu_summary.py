#!/usr/bin/env python
"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into the committed summaries under profiles/:
   python scripts/ncu_summary.py <workload> <report.ncu-rep> [<report2> ...]  ->  profiles/r02_ncu_<workload>.csv
One row per profiled launch: kernel, duration, DRAM bytes, lane utilisation, issue utilisation, registers, top stalls."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
     "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_instr", "smsp__inst_executed.sum": "warp_instr",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
     "launch__registers_per_thread": "registers", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
     "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
     "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
     "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
     "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
     "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
     "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch"}


def to_float(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "us": 1.0, "ms": 1e3, "ns": 1e-3}
    return v * scale.get(u, 1.0)


def main():
    workload = sys.argv[1]
    rows = []
    for rep in sys.argv[2:]:
        if rep.endswith(".csv"):          # `ncu -i x.ncu-rep --page raw --csv` exported on the GPU box (the reports exceed gpurun's copy-back limit)
            rd = list(csv.reader(open(rep)))
        else:
            r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
            rd = list(csv.reader(io.StringIO(r.stdout)))
        hdr, units = rd[0], rd[1]
        for line in rd[2:]:
            d = dict(zip(hdr, line)); un = dict(zip(hdr, units))
            row = {"kernel": d["Kernel Name"].split("(")[0].replace("void ", "").replace(", ", ";"), "report": os.path.basename(rep)}
            for k, name in M.items():
                if k in d and d[k] != "":
                    row[name] = to_float(d[k], un[k])
            if "duration" in row:
                row["duration_us"] = row.pop("duration")
            rows.append(row)
    cols = ["kernel", "duration_us", "dram_read_bytes", "dram_write_bytes", "threads_per_warp_instr", "issue_active_pct", "warps_active_pct", "warp_instr",
            "registers", "l1_hit_pct", "l2_hit_pct"] + sorted({k for r in rows for k in r if k.startswith("stall_")}) + ["report"]
    path = os.path.join(ROOT, "profiles", f"r02_ncu_{workload}.csv")
    with open(path, "w") as f:
        f.write("# ncu --set full --clock-control none, one launch per row (cold caches, serialised); written by scripts/ncu_summary.py\n")
        f.write(",".join(cols) + "\n")
        for r in rows:
            f.write(",".join(("%.6g" % r[c] if isinstance(r.get(c), float) else str(r.get(c, ""))) for c in cols) + "\n")
    print(open(path).read())


if __name__ == "__main__":
    main()
