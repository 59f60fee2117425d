#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (authoring container only: needs /root/reference).  Pins
``oracle/cluster_prep_oracle.py`` against the UNMODIFIED ``preprocess_ilp`` / ``partition_reads`` of
``/root/reference/py/freddie_cluster.py`` (gurobipy is stubbed: these functions never touch it).

For every tint of a few golden sets: SEGMENT text (the segment oracle's, itself pinned byte-for-byte to
the reference's files) -> the reference's ``read_segment`` -> ``preprocess_ilp`` -> ``partition_reads`` for
three settings -> SHA-256 of the canonical serialisation (``cluster_prep_oracle.canonical``), written to
``tests/golden/cluster_prep.json``.

    python oracle/pin_cluster_prep.py
"""
import contextlib
import copy
import hashlib
import importlib
import io
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/py"
SETS = ["cfg1", "cfg2_small", "dup_heavy", "degenerate", "plateau", "cfg3_mini"]
SETTINGS = [("constant", 1000), ("constant", 7), ("relative", 50)]  # "exons" / "introns" raise in the reference (:194)


def reference_module():
    try:
        importlib.import_module("gurobipy")
    except Exception:
        stub = types.ModuleType("gurobipy")  # `from gurobipy import Model, GRB, quicksum, LinExpr` (:13)
        for attr in ("Model", "GRB", "quicksum", "LinExpr"):
            setattr(stub, attr, None)
        sys.modules["gurobipy"] = stub
    sys.path.insert(0, REF)
    try:
        return importlib.import_module("freddie_cluster")
    finally:
        sys.path.remove(REF)


def segment_texts(name):
    """{(chr, id): SEGMENT text} of a golden set, from the segment oracle."""
    from freddie_b200 import synth
    from oracle import segment_oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import flags_to_kwargs
    tints, flags = synth.make_golden_set(name)
    prm = orc.Params(**flags_to_kwargs(flags))
    out = {}
    for t in copy.deepcopy(tints):
        orc.segment_tint(t, prm)
        out[(t["chr"], t["id"])] = orc.format_segment(t)
    return out


def reference_digest(fc, seg_file, model, max_ilp):
    from oracle import cluster_prep_oracle as cpo
    tints = fc.read_segment(seg_file)
    assert len(tints) == 1
    tint = list(tints.values())[0]
    fc.preprocess_ilp(tint, dict(recycle_model=model))
    with contextlib.redirect_stdout(io.StringIO()):  # partition_reads prints every piece
        fc.partition_reads(tint, max_ilp)
    d = tint["ilp_data"]
    U = len(tint["read_reps"])
    gaps = [tint["reads"][tint["read_reps"][i][0]]["gaps"] for i in range(U)]
    cat = [tint["reads"][tint["read_reps"][i][0]]["poly_tail_category"] for i in range(U)]
    s = cpo.canonical(d["I"], d["C"], d["FL"], cat, d["garbage_cost"], gaps, tint["partitions"])
    return hashlib.sha256(s.encode()).hexdigest(), U, len(tint["segs"]), len(tint["partitions"])


def main():
    fc = reference_module()
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in SETS:
            gold[name] = {}
            for (c, i), text in sorted(segment_texts(name).items()):
                f = os.path.join(tmp, "segment_%s_%d.tsv" % (c, i))
                with open(f, "w") as fh:
                    fh.write(text)
                entry = dict(segment_sha256=hashlib.sha256(text.encode()).hexdigest())
                for model, mx in SETTINGS:
                    dg, U, M, P = reference_digest(fc, f, model, mx)
                    entry["%s/%d" % (model, mx)] = dg
                    entry.update(reps=U, segments=M)
                    entry["partitions/%s/%d" % (model, mx)] = P
                gold[name]["%s/%d" % (c, i)] = entry
            print(name, len(gold[name]), "tints pinned", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "cluster_prep.json"), "w") as fh:
        json.dump(gold, fh, sort_keys=True, indent=0)


if __name__ == "__main__":
    main()
