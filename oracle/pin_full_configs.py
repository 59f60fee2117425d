#!/usr/bin/env python3
"""Full-size golden digests of the BASELINE.json configs from the UNMODIFIED reference.

TEST INFRASTRUCTURE (authoring container only: needs ``/root/reference``).  For a full-size config
(``cfg2`` 200 k reads / 3 k tints, ``cfg3`` 20 x 100 k-read giant tints, ``cfg4`` power-law 1..200 k,
``cfg5`` 10 M reads / 60 k tints) it

1. streams the seeded synthetic SPLIT directory to disk (``freddie_b200.synth.iter_config``),
2. runs ``python /root/reference/py/freddie_segment.py -s SPLIT -o OUT -t <cores>`` on it,
3. records, per tint, a truncated SHA-256 of the input files and of the reference's SEGMENT file,
   plus whole-directory digests, in ``tests/golden/full/<cfg>.json`` (a few hundred KB; the data itself
   is never committed).

The GPU-side check (``tests/full_config_check.py``) regenerates the same SPLIT directory on the GPU
box, runs the drop-in CLI and compares every file hash with this manifest.

Usage:  python oracle/pin_full_configs.py --cfg 3 [--tints 4] [--work /tmp/frs_full] [--threads 8] [--impl reference|oracle]
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/py"

from freddie_b200 import synth  # noqa: E402


def sha_file(p):
    h = hashlib.sha256()
    with open(p, "rb") as fh:
        while True:
            b = fh.read(1 << 22)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def parse_select(text):
    """"0-399,60000,60004-60010" -> list of tint indices."""
    out = []
    for part in text.split(","):
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def write_config(cfg, split_dir, workers, scale=1.0, limit=None, chunk_reads=200000, select=None):
    """Streams the config (or its first ``limit`` tints / the tints ``select``) to ``split_dir``; returns
    ({(contig, id): n_reads}, describe dict)."""
    tints_meta = {}
    tot = dict(tints=0, reads=0, intervals=0, positions=0, islands=0)
    for part in synth.iter_config(cfg, scale=scale, workers=workers, limit=limit, chunk_reads=chunk_reads, select=select):
        synth.write_split_dir(part, split_dir)
        d = synth.describe(part)
        for k in tot:
            tot[k] += d[k]
        for t in part:
            tints_meta[(t["chr"], t["id"])] = len(t["reads"])
    return tints_meta, tot


def hash_inputs(split_dir, keys):
    out = {}
    for c, i in keys:
        a = sha_file("%s/%s/split_%s_%d.tsv" % (split_dir, c, c, i))
        b = sha_file("%s/%s/reads_%s_%d.tsv" % (split_dir, c, c, i))
        out["%s/%d" % (c, i)] = hashlib.sha256((a + b).encode()).hexdigest()[:16]
    return out


def hash_outputs(out_dir, keys):
    out = {}
    for c, i in keys:
        p = "%s/%s/segment_%s_%d.tsv" % (out_dir, c, c, i)
        lg = "%s/%s/segment_%s_%d.log" % (out_dir, c, c, i)
        assert os.path.getsize(lg) == 0, lg
        out["%s/%d" % (c, i)] = sha_file(p)[:16]
    return out


def digest(m):
    h = hashlib.sha256()
    for k in sorted(m):
        h.update(k.encode())
        h.update(m[k].encode())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True)
    ap.add_argument("--work", default="/tmp/frs_full")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--impl", default="reference", choices=["reference", "oracle"])
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--tints", type=int, default=None, help="pin only the first k tints of the config (cfg3: one "
                    "100 k-read tint costs the reference tens of minutes)")
    ap.add_argument("--chunk-reads", type=int, default=200000)
    ap.add_argument("--select", default=None, help="pin only these tint indices, e.g. 0-999,60000-60004")
    ap.add_argument("--name", default=None, help="manifest name (default cfg<N>)")
    ap.add_argument("--merge", action="store_true", help="add the pinned tints to an existing manifest of that name")
    a = ap.parse_args()
    name = a.name or "cfg%d" % a.cfg
    sd = os.path.join(a.work, name, "split")
    od = os.path.join(a.work, name, "ref")
    for d in (sd, od):
        shutil.rmtree(d, ignore_errors=True)
    t0 = time.time()
    meta, desc = write_config(a.cfg, sd, a.threads, limit=a.tints, chunk_reads=a.chunk_reads,
                              select=parse_select(a.select) if a.select else None)
    keys = sorted(meta)
    t_gen = time.time() - t0
    print("%s: generated %s in %.0fs" % (name, desc, t_gen), flush=True)
    m_in = hash_inputs(sd, keys)
    t0 = time.time()
    if a.impl == "reference":
        subprocess.run([sys.executable, "-W", "ignore", os.path.join(REF, "freddie_segment.py"), "-s", sd, "-o", od,
                        "-t", str(a.threads)], check=True, stdout=subprocess.DEVNULL)
    else:
        from oracle import segment_oracle as orc
        orc.run_dir(sd, od, orc.Params(), a.threads)
    t_ref = time.time() - t0
    m_out = hash_outputs(od, keys)
    n_files = sum(len(f) for _, _, f in os.walk(od))
    assert n_files == 2 * len(keys), (n_files, len(keys))
    gold = os.path.join(ROOT, "tests", "golden", "full")
    os.makedirs(gold, exist_ok=True)
    n_reads = {k: meta[tuple([k.split("/")[0], int(k.split("/")[1])])] for k in m_in}
    runs = [dict(tints=len(keys), reads=desc["reads"], threads=a.threads, seconds=round(t_ref, 1), select=a.select,
                 limit=a.tints)]
    if a.merge and os.path.exists(os.path.join(gold, name + ".json")):
        with open(os.path.join(gold, name + ".json")) as fh:
            old = json.load(fh)
        assert old["impl"] == a.impl
        for k in old["inputs"]:
            if k in m_in:
                assert old["inputs"][k] == m_in[k] and old["outputs"][k] == m_out[k], "re-pinned tint %s differs" % k
        m_in = dict(old["inputs"], **m_in)
        m_out = dict(old["outputs"], **m_out)
        n_reads = dict(old["n_reads"], **n_reads)
        runs = old.get("runs", [dict(tints=old["tints_pinned"], reads=old["describe"]["reads"], threads=old["threads"],
                                     seconds=old["seconds"])]) + runs
        desc = dict(tints=len(m_in), reads=sum(n_reads.values()))
    # reads/s of the reference over every pinning run (wall clock, threads as recorded per run)
    rps = round(sum(r["reads"] for r in runs) / max(sum(r["seconds"] for r in runs), 1e-9), 1)
    man = dict(config=name, source_cfg=a.cfg, impl=a.impl, tints_pinned=len(m_in), describe=desc, threads=a.threads,
               seconds=round(sum(r["seconds"] for r in runs), 1), reads_per_sec=rps, runs=runs,
               input_digest=digest(m_in), output_digest=digest(m_out), n_reads=n_reads, inputs=m_in, outputs=m_out)
    with open(os.path.join(gold, name + ".json"), "w") as fh:
        json.dump(man, fh, sort_keys=True, separators=(",", ":"))
    print("%s: %s took %.0fs (%.0f reads/s on %d threads); output digest %s" % (
        name, a.impl, t_ref, desc["reads"] / t_ref, a.threads, man["output_digest"]), flush=True)
    if not a.keep:
        shutil.rmtree(os.path.join(a.work, name), ignore_errors=True)


if __name__ == "__main__":
    main()
