#!/usr/bin/env python3
"""Full-size parity check of a BASELINE.json config on the GPU box (not a pytest: minutes of work).

Regenerates the seeded synthetic SPLIT directory of ``--cfg`` (same generator and seeds as
``oracle/pin_full_configs.py``), verifies the per-tint INPUT hashes against
``tests/golden/full/cfgN.json`` (so both machines segment the same bytes), runs the drop-in CLI path
(native parser -> CUDA pipeline -> native formatter) and compares the SHA-256 of EVERY SEGMENT file
with the digest recorded from the unmodified reference (or, for cfg5, the pinned oracle).

    python tests/full_config_check.py --cfg 3 [--tints 4] [--gpus 1] [--threads 16] [--work /tmp/frs_full]

``--tints k`` checks the first k tints of the config only (full-size tints, byte-identical to the same
tints of the whole config: every tint is generated from its own child seed) -- generating all of cfg3
(2 M reads of 43 intervals, 38 GB of SPLIT text) takes longer than the whole GPU budget of a round.

Prints one JSON line (also appended to gpurun_out/full_config_check.jsonl).
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True)
    ap.add_argument("--work", default="/tmp/frs_full")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--batch-reads", type=int, default=131072)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--tints", type=int, default=None, help="only the first k tints of the config")
    ap.add_argument("--chunk-reads", type=int, default=200000, help="reads generated per pool round")
    a = ap.parse_args()
    name = "cfg%d" % a.cfg
    with open(os.path.join(ROOT, "tests", "golden", "full", name + ".json")) as fh:
        man = json.load(fh)
    sd = os.path.join(a.work, name, "split")
    od = os.path.join(a.work, name, "seg")
    for d in (sd, od):
        shutil.rmtree(d, ignore_errors=True)
    t0 = time.time()
    n_reads = 0
    made = set()
    for part in synth.iter_config(a.cfg, workers=a.threads, limit=a.tints, chunk_reads=a.chunk_reads):
        synth.write_split_dir(part, sd)
        n_reads += sum(len(t["reads"]) for t in part)
        made.update("%s/%d" % (t["chr"], t["id"]) for t in part)
    t_gen = time.time() - t0
    keys = sorted(k for k in man["inputs"] if k in made)
    assert len(keys) == len(made), "generated tints missing from the manifest"
    bad_in = 0
    for k in keys:
        c, i = k.split("/")
        x = sha_file("%s/%s/split_%s_%s.tsv" % (sd, c, c, i))
        y = sha_file("%s/%s/reads_%s_%s.tsv" % (sd, c, c, i))
        bad_in += hashlib.sha256((x + y).encode()).hexdigest()[:16] != man["inputs"][k]
    from freddie_b200.engine import SegmentParams
    from freddie_b200.segment import run_directory
    t0 = time.time()
    stats = run_directory(sd, od, SegmentParams(), threads=a.threads, gpus=a.gpus, batch_reads=a.batch_reads,
                          progress=False)
    t_run = time.time() - t0
    bad_out, missing = 0, 0
    for k in keys:
        c, i = k.split("/")
        p = "%s/%s/segment_%s_%s.tsv" % (od, c, c, i)
        lg = "%s/%s/segment_%s_%s.log" % (od, c, c, i)
        if not os.path.exists(p) or not os.path.exists(lg) or os.path.getsize(lg) != 0:
            missing += 1
            continue
        bad_out += sha_file(p)[:16] != man["outputs"][k]
    n_files = sum(len(f) for _, _, f in os.walk(od))
    line = dict(config=name, pinned_by=man["impl"], tints=len(keys), tints_in_config=len(man["inputs"]), reads=n_reads,
                gpus=stats["gpus"],
                host_threads=a.threads, generate_seconds=round(t_gen, 1), cli_seconds=round(t_run, 2),
                cli_reads_per_sec=round(n_reads / t_run, 1), reference_reads_per_sec=man["reads_per_sec"],
                reference_threads=man["threads"], dp_cells=stats["dp_cells"], input_mismatches=int(bad_in),
                output_mismatches=int(bad_out), missing=int(missing), extra_files=int(n_files - 2 * len(keys)),
                ok=bool(bad_in == 0 and bad_out == 0 and missing == 0 and n_files == 2 * len(keys)))
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "full_config_check.jsonl"), "a") as fh:
        fh.write(json.dumps(line) + "\n")
    if not a.keep:
        shutil.rmtree(os.path.join(a.work, name), ignore_errors=True)
    sys.exit(0 if line["ok"] else 1)


if __name__ == "__main__":
    main()
