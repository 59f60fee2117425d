#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA pipeline with the oracle on the golden sets (dev tool).

    python tests/gpu_debug.py [set ...]     (needs a GPU; run through gpurun)
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from freddie_b200 import synth, _lib  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams, apply_result, format_tint  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402
from oracle import segment_oracle as orc  # noqa: E402


def flags_to_params(flags):
    kw = {}
    it = iter(flags)
    for f in it:
        if f == "--consider-ends":
            kw["consider_ends"] = True
        else:
            v = next(it)
            kw[{"-sd": "sigma", "-tp": "tp", "-vf": "vf", "-mps": "mps", "-lo": "lo"}[f]] = (
                int(v) if f in ("-mps", "-lo") else float(v))
    o = orc.Params(**kw)
    g = SegmentParams(o.sigma, o.tp, o.vf, o.mps, o.lo, o.ignore_ends)
    return o, g


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    if a.dtype.kind == "f":
        d = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
    else:
        d = np.flatnonzero(a != b)
    if len(d) == 0:
        return None
    i = int(d[0])
    return "%d diffs, first at %d: %r vs %r" % (len(d), i, a[i], b[i])


def check_set(eng, name):
    tints, flags = synth.make_golden_set(name)
    oprm, gprm = flags_to_params(flags)
    t0 = time.time()
    batch = pack_tints(tints)
    t_pack = time.time() - t0
    eng.set_profiling(True)
    t0 = time.time()
    res = eng.segment_batch(batch, gprm)
    t_gpu = time.time() - t0
    print("== %s: %s pack %.2fs gpu %.3fs sizes %s" % (name, batch.counts(), t_pack, t_gpu, res.sizes))
    print("   timings:", ["%s %.3fms x%d" % t for t in eng.timings()])
    # oracle
    import copy
    otints = copy.deepcopy(tints)
    t0 = time.time()
    inters = [orc.segment_tint(t, oprm, keep=True) for t in otints]
    print("   oracle %.1fs" % (time.time() - t0))
    bad = 0
    yraw = eng.tap(_lib.TAP_Y_RAW, np.int32)
    y = eng.tap(_lib.TAP_Y, np.float64)
    thr = eng.tap(_lib.TAP_THR, np.float64)
    cand = eng.tap(_lib.TAP_CAND, np.int32)
    fixed = eng.tap(_lib.TAP_FIXED, np.uint8)
    dpf = eng.tap(_lib.TAP_DP_FINAL, np.uint8)
    o_yraw = np.concatenate([np.concatenate(i["Y_raw"]) for i in inters])
    o_y = np.concatenate([np.concatenate(i["Y"]) for i in inters])
    o_thr = np.array([i["thr"] for i in inters])
    iso = batch.arrays["island_sample_off"]
    o_cand, o_fixed, o_dpf = [], [], []
    k = 0
    for i in inters:
        for a in range(len(i["cand"])):
            c = np.array(i["cand"][a])
            o_cand.append(c + iso[k])
            f = np.zeros(len(c), dtype=np.uint8)
            f[i["fixed"][a]] = 1
            o_fixed.append(f)
            f = np.zeros(len(c), dtype=np.uint8)
            f[i["dp_final"][a]] = 1
            o_dpf.append(f)
            k += 1
    for label, g, o in [("Y_raw", yraw, o_yraw.astype(np.int32)), ("Y", y, o_y), ("thr", thr, o_thr),
                        ("cand", cand, np.concatenate(o_cand)), ("fixed", fixed, np.concatenate(o_fixed)),
                        ("dp_final", dpf, np.concatenate(o_dpf))]:
        d = first_diff(g, o)
        print("   %-9s %s" % (label, "ok" if d is None else "MISMATCH " + d))
        bad += d is not None
    # final positions
    fo = res.arrays["tint_final_off"]
    for t, (ot, it) in enumerate(zip(otints, inters)):
        g = res.arrays["final_pos"][fo[t]:fo[t + 1]].tolist()
        o = ot["final_positions"]
        if g != o:
            sg, so = set(g), set(o)
            extra = set()
            for a, (s0, e0) in enumerate(ot["intervals"]):
                extra |= {s0 + x for x in it["refine"][a]}
            print("   tint %d finals: gpu-only %s  oracle-only %s (oracle-only that are refine extras: %s)" % (
                t, sorted(sg - so)[:10], sorted(so - sg)[:10], sorted((so - sg) & extra)[:10]))
            bad += 1
    # end to end text
    n_bad_t = 0
    for t, ot in enumerate(otints):
        txt = format_tint(batch, res, t)
        ref = orc.format_segment(ot)
        if txt != ref:
            n_bad_t += 1
            if n_bad_t <= 3:
                gl, rl = txt.split("\n"), ref.split("\n")
                for ln, (x, z) in enumerate(zip(gl, rl)):
                    if x != z:
                        print("   tint %d line %d:\n     gpu %s\n     ref %s" % (t, ln, x[:300], z[:300]))
                        break
    print("   text: %d/%d tints differ" % (n_bad_t, len(otints)))
    ties = sum(len(i["ties"]) for i in inters)
    if ties:
        print("   note: %d refine tie(s) between equal-height peaks" % ties)
    return bad + n_bad_t


def main():
    names = sys.argv[1:] or ["degenerate", "plateau", "cfg2_flagsA", "cfg2_small", "cfg1", "cfg2_flagsB", "cfg4_mini",
                             "cfg5_mini", "cfg3_mini"]
    eng = Engine(0)
    total = 0
    for n in names:
        try:
            total += check_set(eng, n)
        except Exception:
            traceback.print_exc()
            total += 1
            eng = Engine(0)
    print("TOTAL MISMATCHES", total)


if __name__ == "__main__":
    main()
