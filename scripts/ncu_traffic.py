#!/usr/bin/env python
"""Records the DRAM traffic of the dominant kernel of a workload from an `ncu --set full`
capture: dram__bytes_read.sum + dram__bytes_write.sum of the longest EvaluateKernel launch in
the report, written into profiles/r2_traffic.json (bench.py reads `roofline.traffic` from
there).  Also prints a short summary of that launch for profiles/*_ncu_summary.txt.

  scripts/ncu_traffic.py L gpurun_out/prof_r2_bench_L.ncu-rep [--summary profiles/r2_L_ncu_summary.txt]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def main():
    workload, rep = sys.argv[1], sys.argv[2]
    want = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "EvaluateKernel"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    best, best_ms = None, -1.0
    for r in data:
        if want not in r[col["Kernel Name"]]:
            continue
        ms = float(r[col["gpu__time_duration.sum"]]) * TIME.get(units[col["gpu__time_duration.sum"]], 1.0)
        if ms > best_ms:
            best, best_ms = r, ms
    if best is None:
        raise SystemExit(f"no kernel matching {want!r} in {rep}")

    def nbytes(key):
        return float(best[col[key]]) * UNIT[units[col[key]]]

    traffic = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    rec = json.load(open(path)) if os.path.exists(path) else {}
    rec[workload] = {"dram_bytes_per_launch": traffic,
                     "dram_bytes_read": nbytes("dram__bytes_read.sum"),
                     "dram_bytes_write": nbytes("dram__bytes_write.sum"),
                     "kernel": best[col["Kernel Name"]], "launch_ms_under_ncu": best_ms,
                     "source": os.path.relpath(rep, ROOT) + " (ncu --set full --clock-control none)"}
    json.dump(rec, open(path, "w"), indent=1, sort_keys=True)
    lines = [f"# {os.path.relpath(rep, ROOT)}", f"kernel: {best[col['Kernel Name']]}",
             f"dram traffic per launch: {traffic / 1e9:.4f} GB"]
    for k in KEYS:
        if k in col:
            lines.append(f"{k:90s} {best[col[k]]:>16s} {units[col[k]]}")
    text = "\n".join(lines) + "\n"
    if "--summary" in sys.argv:
        open(sys.argv[sys.argv.index("--summary") + 1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
