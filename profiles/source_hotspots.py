#!/usr/bin/env python3
"""Maps the per-instruction samples of an ncu report (--set full --import-source on) to source lines with
nvdisasm's line info of the SAME build:  python profiles/source_hotspots.py report.ncu-rep 'k_dp_warpILi16E' [N]
(the CSV of `--page source --print-source cuda` carries no metric columns in this ncu, hence the detour)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, mangled = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("FRS_LIB") or os.path.join(root, "freddie_b200", "libfreddie_b200.so")
src_root = os.environ.get("FRS_SRC") or os.path.join(root, "freddie_b200", "csrc")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
instrs = None
for cubin in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    start = None
    for i, l in enumerate(txt):
        if ".section" in l and ".text." in l and mangled in l:
            start = i
            break
    if start is None:
        continue
    cur, instrs = None, []
    for l in txt[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            instrs.append((m.group(2).strip(), cur))
    break
assert instrs, "kernel not found in the built library"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
inst_rows, k, hdr = [], None, None
want = os.environ.get("KNAME") or re.sub(r"ILi(\d+)E", r"<(int)\1>", mangled)
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        k, hdr = row[1], None
        if want in k.replace("void ", ""):
            inst_rows.append([])
        continue
    if row and row[0] == "Address":
        hdr = row
        continue
    if k and hdr and len(row) > 6 and want in k.replace("void ", ""):
        inst_rows[-1].append((row[1].strip(), int(row[hdr.index("# Samples")]), int(row[hdr.index("Instructions Executed")])))
# several launches of the kernel: the one that executed most instructions
rows = max(inst_rows, key=lambda rs: sum(r[2] for r in rs))[:len(instrs)]
assert len(rows) == len(instrs) and all(a[0].split()[0] == b[0].split()[0] for a, b in zip(instrs, rows)), "SASS of the report differs from the build"
inst, smp = collections.Counter(), collections.Counter()
for (_, line), (_, s, n) in zip(instrs, rows):
    inst[line] += n
    smp[line] += s
ti, ts = sum(inst.values()), sum(smp.values())
print("%s: %d warp instructions, %d samples" % (want, ti, ts))
src = {}
for line, v in sorted(smp.items(), key=lambda kv: -kv[1])[:top]:
    text = ""
    if line:
        p = os.path.join(src_root, line[0])
        if os.path.exists(p):
            src.setdefault(p, open(p).read().split("\n"))
            text = src[p][line[1] - 1].strip()[:100]
    print("%5.1f %% samples %5.1f %% inst  %s:%s  %s" % (100 * v / ts, 100 * inst[line] / ti, line[0] if line else "?", line[1] if line else "", text))
