#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass), written to profiles/sass_summary.txt.

Evidence for DESIGN.md: the library holds sm_100a code only; the DP table kernels stage coverage rows with TMA
bulk copies (UBLKCP + SYNCS mbarrier arrive / try_wait); the fp64 Gaussian / variance kernels contain no fused
multiply-add outside the IEEE divide / sqrt sequences (DFMA appears only there), so their sums round like
scipy's / numpy's; the mask / triple phases of the DP are VOTE / LOP3 / POPC code.

    python profiles/sass_summary.py [path/to/libfreddie_b200.so]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "freddie_b200", "libfreddie_b200.so")
WATCH = ["ACQBULK", "PREEXIT", "UBLKCP", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "VOTE", "POPC", "LOP3", "REDUX", "MATCH", "ATOMS", "ATOMG", "RED",
         "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "NANOSLEEP"]
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
fn = None
cnt = collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        cnt[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and fn:
        op = m.group(1)
        cnt[fn]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                cnt[fn][w] += 1
out = ["# %s" % os.path.relpath(lib, ROOT), "# cubin architectures: %s" % ", ".join(archs),
       "# per kernel: instructions in the SASS listing (static counts, not executed counts)", ""]
cols = ["total"] + WATCH
out.append("%-44s" % "kernel" + "".join("%8s" % c[:7] for c in cols))
for fn, c in cnt.items():
    out.append("%-44s" % fn[:43] + "".join("%8d" % c[k] for k in cols))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:8]))
