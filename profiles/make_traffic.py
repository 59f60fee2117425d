#!/usr/bin/env python3
"""DRAM traffic per pipeline stage (dram__bytes_read.sum + dram__bytes_write.sum, summed over the
stage's kernels, one step) from an `ncu --set full` raw CSV of profiles/profile_step.py:

  python profiles/make_traffic.py gpurun_out/step_raw.csv > profiles/traffic.json

bench.py copies the dominant stage's figure into `roofline.traffic`.  The capture is cold-cache and
serialised; writes that stay in the 126 MB L2 at the end of a kernel do not show up as DRAM writes.
"""
import csv
import json
import sys

STAGE = [  # first match wins
    ("k_zero_multi", "signal"), ("k_signal", "signal"), ("k_smooth", "smooth"), ("k_tile_prefix", "lists"), ("k_tile_lists", "lists"), ("k_threshold", "threshold"),
    ("k_cand_meta", "cand_meta"), ("k_fixed", "fixed"), ("k_sub_build", "subproblems"), ("k_tint_cov", "coverage"),
    ("k_coverage", "coverage"), ("k_sub_fill", "dp_plan"), ("k_dp_solve", "dp_solve"), ("k_dp", "dp"),
    ("k_final_mark", "refine"), ("k_refine", "refine"), ("k_flag_", "finals"), ("k_final_meta", "finals"), ("k_digit_sizes", "finals"),
    ("k_seg_cuts", "finals"), ("k_digits", "digits"), ("k_gap_count", "runs"), ("k_run_fill", "runs"),
    ("k_gap_prep", "gaps"), ("k_gap_sizes", "gaps"), ("k_poly", "poly"), ("k_gap_finish", "poly"),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    out, per_kernel = {}, {}
    prev = None
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "")
        b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        st = next((s for k, s in STAGE if name.startswith(k)), None)
        if st is None:  # scans / compactions belong to the stage of the kernel before them
            st = prev or "other"
        prev = st
        out[st] = out.get(st, 0) + int(b)
        per_kernel[name] = per_kernel.get(name, 0) + int(b)
    out["_per_kernel"] = per_kernel
    out["_source"] = "ncu --set full --clock-control none, profiles/profile_step.py (cfg2, one step)"
    json.dump(out, sys.stdout, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
