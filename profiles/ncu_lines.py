#!/usr/bin/env python3
"""Per-source-line hot spots of a kernel from an ncu report captured with --import-source on.

  python profiles/ncu_lines.py gpurun_out/x.ncu-rep 'regex:k_dp_warp' [top] [launch index]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # sections: ("File Path", p) ("Function Name", f) header, data...
    launches = []  # list of (function, {(file, line): [src, samples, inst, thread_inst]})
    cur = None
    fpath = None
    hdr = None
    last_fn = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            continue
        if r[0] == "Function Name":
            if last_fn != r[1] or (cur is not None and fpath in cur["files"]):
                cur = dict(fn=r[1], lines={}, files=set())
                launches.append(cur)
            last_fn = r[1]
            cur["files"].add(fpath)
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or cur is None or r[0] == "":
            continue
        try:
            cur["lines"][(fpath.split("/")[-1], int(r[0]))] = [r[1], float(r[i_s] or 0), float(r[i_i] or 0), float(r[i_t] or 0)]
        except (ValueError, IndexError):
            pass
    if not launches:
        print("no kernel matched")
        return
    L = launches[min(which, len(launches) - 1)]
    tot_s = sum(v[1] for v in L["lines"].values()) or 1
    tot_i = sum(v[2] for v in L["lines"].values()) or 1
    print("%s  (launch %d of %d)  samples %d  warp-inst %d" % (L["fn"], which, len(launches), tot_s, tot_i))
    for (f, ln), v in sorted(L["lines"].items(), key=lambda kv: -kv[1][1])[:top]:
        print("%6.2f%% smp %6.2f%% inst thr/inst %5.1f | %s:%d | %s" % (
            100 * v[1] / tot_s, 100 * v[2] / tot_i, v[3] / max(v[2], 1), f, ln, v[0].strip()[:100]))


if __name__ == "__main__":
    main()
