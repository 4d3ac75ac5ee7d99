"""Summarise ncu reports / launch lists from gpurun_out/ into profiles/ (text + traffic.json).

    python scripts/ncu_summary.py launches gpurun_out/r1_launches_config4.csv
    python scripts/ncu_summary.py report gpurun_out/r1_config4_full.ncu-rep
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r[-1]))
    total = sum(sum(v) for v in agg.values())
    print("%-60s %6s %12s %12s %7s" % ("kernel", "calls", "avg_ns", "total_ns", "share"))
    for k, v in agg.items():
        print("%-60s %6d %12.0f %12.0f %6.1f%%" % (k[:60], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        print("== %s  grid %s block %s" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                print("   %-82s %14s %s" % (k, r[i], units[i]))
        def to_bytes(k):
            v, u = vals[k]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if "dram__bytes_read.sum" in vals:
            tr = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            print("   traffic (dram read+write) = %.1f MB" % (tr / 1e6))
            out[name.split("<")[0].split("::")[-1]] = int(tr)
    print(json.dumps(out))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
