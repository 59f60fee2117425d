#!/usr/bin/env python3
"""Summarises ncu outputs into the small text files committed under profiles/.

  python profiles/summarize_ncu.py full   gpurun_out/prof.ncu-rep   > profiles/rNN_<name>_full.txt
  python profiles/summarize_ncu.py launch gpurun_out/launches.csv   > profiles/rNN_<name>_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def full(path):
    if path.endswith(".csv"):  # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.csv
        txt = open(path).read()
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[hdr.index("Kernel Name")].split("(")[0])
        d = {}
        for m in FULL:
            if m in hdr:
                d[m] = r[hdr.index(m)]
                print("  %-86s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            ut = units[hdr.index("gpu__time_duration.sum")]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[ut.replace("second", "s") if ut in ("second",) else ut.replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns")]
            def b(k):
                v = float(d[k].replace(",", ""))
                u = units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print("  %-86s %.1f MB  -> %.1f GB/s" % ("dram traffic (read+write) per launch", tr / 1e6, tr / t_s / 1e9))
        except Exception as e:  # keep the raw lines even if a unit is unexpected
            print("  (traffic summary unavailable: %s)" % e)


def launch(path):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-44s %8s %12s %7s" % ("kernel", "launches", "total us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.1f %6.1f%%" % (k[:44], a[0], a[1], 100 * a[1] / tot))
    print("%-44s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2])
