#!/usr/bin/env python3
"""Replaces the rows of re-captured kernels in an `ncu --page raw --csv` export:

    python profiles/merge_raw.py step_raw.csv changed_raw.csv > merged_raw.csv

The whole-step `--set full` capture costs ~6 GPU-minutes; after a change to a few kernels only those are captured
again (`-k regex:...`, same command, same workload) and their rows replace the old ones by kernel name (a kernel
that no longer exists is dropped, a new one is appended after the last replaced row).  Columns are matched by
name."""
import csv
import sys

base = list(csv.reader(open(sys.argv[1])))
new = list(csv.reader(open(sys.argv[2])))
hb, hn = base[0], new[0]
ik, jk = hb.index("Kernel Name"), hn.index("Kernel Name")


def key(name):
    return name.split("(")[0].replace("void ", "").strip()


fresh = {}
for r in new[2:]:
    fresh.setdefault(key(r[jk]), []).append(r)
cols = [hn.index(c) if c in hn else None for c in hb]
out, used, last = [base[0], base[1]], set(), 2
for r in base[2:]:
    k = key(r[ik])
    if k in fresh:
        if k not in used:
            used.add(k)
            for nr in fresh[k]:
                out.append([nr[c] if c is not None else "" for c in cols])
            last = len(out)
    else:
        out.append(r)
extra = [[nr[c] if c is not None else "" for c in cols] for k, rows in fresh.items() if k not in used for nr in rows]
out[last:last] = extra
csv.writer(sys.stdout, quoting=csv.QUOTE_ALL).writerows(out)
