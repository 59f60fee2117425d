#!/usr/bin/env python3
"""Pins the oracle against the UNMODIFIED reference and (re)generates ``tests/golden/``.

Runs only in the authoring container (needs ``/root/reference``).  For every named golden set
(``freddie_b200.synth.GOLDEN_SETS``):

1. realise the seeded synthetic SPLIT directory,
2. run ``python /root/reference/py/freddie_segment.py`` on it (the live oracle),
3. run ``oracle/segment_oracle.py`` on it and require a byte-identical SEGMENT directory,
4. record SHA-256 manifests of inputs and reference outputs in ``tests/golden/manifest.json``.

For ``cfg1`` it also imports the reference module and dumps the reference's own intermediates
(``Y_raw``, smoothed ``Y``, variance threshold, candidates, coverage checksum, final positions) into
``tests/golden/cfg1_intermediates.npz`` after checking the oracle's bit-for-bit.  The tiny
``degenerate`` and ``plateau`` sets are committed wholesale (inputs + reference outputs).

Usage:  python oracle/pin_against_reference.py [--sets a,b,...] [--keep DIR]
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/py"

from freddie_b200 import synth  # noqa: E402
from oracle import segment_oracle as orc  # noqa: E402


def sha_dir(d):
    out = {}
    for base, _, files in os.walk(d):
        for fn in sorted(files):
            p = os.path.join(base, fn)
            with open(p, "rb") as fh:
                out[os.path.relpath(p, d)] = hashlib.sha256(fh.read()).hexdigest()
    return dict(sorted(out.items()))


def digest(m):
    h = hashlib.sha256()
    for k, v in sorted(m.items()):
        h.update(k.encode())
        h.update(v.encode())
    return h.hexdigest()


def params_from_flags(flags):
    kw = {}
    it = iter(flags)
    for f in it:
        if f == "--consider-ends":
            kw["consider_ends"] = True
        else:
            v = next(it)
            kw[{"-sd": "sigma", "-tp": "tp", "-vf": "vf", "-mps": "mps", "-lo": "lo"}[f]] = (
                int(v) if f in ("-mps", "-lo") else float(v))
    return orc.Params(**kw)


def reference_intermediates(split_dir, contig, tint_id, prm):
    """Calls the reference's own functions step by step (freddie_segment.py:738-814)."""
    sys.path.insert(0, REF)
    import freddie_segment as ref
    from scipy.ndimage import gaussian_filter1d
    tint = ref.read_split("%s/%s/split_%s_%d.tsv" % (split_dir, contig, contig, tint_id))[0]
    ref.read_sequence(tint, "%s/%s/reads_%s_%d.tsv" % (split_dir, contig, contig, tint_id))
    pos_to, to_pos, Y_raw = ref.process_splicing_data(tint, prm.ignore_ends)
    Y = [gaussian_filter1d(y, prm.sigma, truncate=4.0) for y in Y_raw]
    nz = np.array([v for y in Y for v in y if v > 0])
    thr = nz.mean() + prm.vf * nz.std()
    cands, fixeds, csum = [], [], []
    for a in range(len(Y)):
        c = ref.candidates_from_peaks(Y[a])
        C = ref.get_cumulative_coverage(tint["read_reps"], c, to_pos[a], pos_to, a)
        fx = {0, len(c) - 1} | {i for i, yi in enumerate(c) if Y[a][yi] > thr}
        fx, _ = ref.break_large_problems(c, fx, Y[a], prm.mps)
        cands.append(c)
        fixeds.append(sorted(fx))
        csum.append(int(C.astype(np.uint64).sum()))
    ref.segment(tint, prm.sigma, ref.smooth_threshold(prm.tp), prm.tp, prm.vf, prm.mps, prm.lo,
                prm.ignore_ends)
    return dict(Y_raw=Y_raw, Y=Y, thr=thr, cand=cands, fixed=fixeds, csum=csum,
                final_positions=tint["final_positions"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", default=",".join(synth.GOLDEN_SETS))
    ap.add_argument("--keep", default=None, help="keep work directories here")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    mpath = os.path.join(gold, "manifest.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {}
    work = a.keep or tempfile.mkdtemp(prefix="frs_pin_")
    for name in a.sets.split(","):
        tints, flags = synth.make_golden_set(name)
        sd = os.path.join(work, name, "split")
        rd = os.path.join(work, name, "ref")
        od = os.path.join(work, name, "oracle")
        for d in (sd, rd, od):
            shutil.rmtree(d, ignore_errors=True)
        synth.write_split_dir(tints, sd)
        t0 = time.time()
        subprocess.run([sys.executable, "-W", "ignore", os.path.join(REF, "freddie_segment.py"), "-s", sd,
                        "-o", rd, "-t", str(a.threads)] + flags, check=True, stdout=subprocess.DEVNULL)
        t_ref = time.time() - t0
        prm = params_from_flags(flags)
        t0 = time.time()
        orc.run_dir(sd, od, prm, a.threads)
        t_orc = time.time() - t0
        m_ref, m_orc = sha_dir(rd), sha_dir(od)
        assert m_ref == m_orc, "oracle differs from reference on %s: %s" % (
            name, [k for k in m_ref if m_ref[k] != m_orc.get(k)][:5])
        m_in = sha_dir(sd)
        manifest[name] = dict(flags=flags, describe=synth.describe(tints), input_digest=digest(m_in),
                              output_digest=digest(m_ref), outputs=m_ref,
                              reference_seconds=round(t_ref, 2), oracle_seconds=round(t_orc, 2),
                              threads=a.threads)
        print("%-12s OK  ref %.1fs oracle %.1fs  %s" % (name, t_ref, t_orc, manifest[name]["describe"]))
        if name in ("degenerate", "plateau"):
            dst = os.path.join(gold, name)
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(sd, os.path.join(dst, "split"))
            shutil.copytree(rd, os.path.join(dst, "segment"))
        if name == "cfg1":
            contig, tid = tints[0]["chr"], tints[0]["id"]
            ri = reference_intermediates(sd, contig, tid, prm)
            t = orc.parse_split("%s/%s/split_%s_%d.tsv" % (sd, contig, contig, tid))
            orc.parse_reads(t, "%s/%s/reads_%s_%d.tsv" % (sd, contig, contig, tid))
            oi = orc.segment_tint(t, prm, keep=True)
            for x, y in zip(ri["Y_raw"], oi["Y_raw"]):
                assert np.array_equal(x, y)
            for x, y in zip(ri["Y"], oi["Y"]):
                assert np.array_equal(x, y), "smoothed signal not bit-identical"
            assert ri["thr"] == oi["thr"], (ri["thr"], oi["thr"])
            assert ri["cand"] == oi["cand"] and ri["fixed"] == oi["fixed"]
            assert ri["final_positions"] == t["final_positions"]
            off = np.cumsum([0] + [len(y) for y in ri["Y"]])
            coff = np.cumsum([0] + [len(c) for c in ri["cand"]])
            foff = np.cumsum([0] + [len(c) for c in ri["fixed"]])
            np.savez_compressed(
                os.path.join(gold, "cfg1_intermediates.npz"),
                Y_raw=np.concatenate(ri["Y_raw"]), Y=np.concatenate(ri["Y"]), thr=np.float64(ri["thr"]),
                island_off=off, cand=np.concatenate(ri["cand"]), cand_off=coff,
                fixed=np.concatenate(ri["fixed"]), fixed_off=foff, csum=np.array(ri["csum"], dtype=np.uint64),
                final_positions=np.array(ri["final_positions"], dtype=np.int64))
            print("cfg1 intermediates: oracle bit-identical to reference functions")
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)
    if not a.keep:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
