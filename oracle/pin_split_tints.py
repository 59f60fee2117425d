#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (authoring container only: needs /root/reference).  Pins ``oracle/split_tints_oracle.py``
against the UNMODIFIED ``get_transcriptional_intervals`` / ``break_tint`` of ``/root/reference/py/freddie_split.py``
(pysam is stubbed: these functions never touch it) on the seeded groups of ``split_tints_oracle.GOLDEN_GROUPS`` and
writes ``tests/golden/split_tints.json`` (SHA-256 of the canonical serialisation, tints, reads, intervals per group).

    python oracle/pin_split_tints.py
"""
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/py"


def reference_module():
    try:
        importlib.import_module("pysam")
    except Exception:
        stub = types.ModuleType("pysam")  # `import pysam` (:12); only read_sam / get_intervals use it
        # the CIGAR operation codes of the SAM specification (the module-level tables :64-101 name them)
        for code, name in enumerate(["CMATCH", "CINS", "CDEL", "CREF_SKIP", "CSOFT_CLIP", "CHARD_CLIP", "CPAD", "CEQUAL", "CDIFF", "CBACK"]):
            setattr(stub, name, code)
        sys.modules["pysam"] = stub
    sys.path.insert(0, REF)
    try:
        return importlib.import_module("freddie_split")
    finally:
        sys.path.remove(REF)


def reference_tints(fs, group):
    reads = [dict(id=i, name="r%d" % i, contig="c", strand="+", simple_tints=list(), tint=None,
                  intervals=[(s, e, 0, e - s, [(0, e - s)]) for s, e in ivs]) for i, ivs in enumerate(group)]
    return [(t["intervals"], t["rids"]) for t in fs.get_transcriptional_intervals(reads=reads)]


def main():
    from oracle import split_tints_oracle as sto
    fs = reference_module()
    gold = {}
    for name, kw in sto.GOLDEN_GROUPS.items():
        group = sto.make_group(**kw)
        ref = reference_tints(fs, group)
        mine = sto.transcriptional_intervals(group)
        assert sto.canonical(ref) == sto.canonical(mine), name
        gold[name] = dict(sha256=sto.digest_of(ref), tints=len(ref), reads=len(group), intervals=sum(len(r) for r in group),
                          largest_tint_intervals=max([len(t[0]) for t in ref] + [0]))
        print(name, gold[name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "split_tints.json"), "w") as fh:
        json.dump(gold, fh, sort_keys=True, indent=1)


if __name__ == "__main__":
    main()
