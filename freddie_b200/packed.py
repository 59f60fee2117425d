"""Packed SPLIT side-channel: a SPLIT directory as a few binary batch files.

``freddie_split.py`` writes, and ``freddie_segment.py`` re-parses, ~1 KB of text per read
(freddie_split.py:445-481, freddie_segment.py:121-185).  A pipeline that runs both stages can hand the
tints over in the packed form the GPU consumes instead::

    python -m freddie_b200.packed -s <SPLIT dir> -o <PACKED dir> [-t threads] [--batch-reads N]
    python -m freddie_b200.segment -s <PACKED dir> -o <SEGMENT dir> ...        # detected by its index

A packed directory holds ``index.json`` (format tag, and for every batch its file and the
``(contig, tint id)`` of its tints in order) and one ``batch_<k>.frsb`` per batch (layout:
``csrc/host_io.cpp``, "FRSBATC1").  The SEGMENT output of a packed directory is byte-identical to that of
the SPLIT directory it was made from (tested); text stays the default.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import List, Optional, Tuple

from . import hostio, schedule

FORMAT = "freddie-b200-packed-split-1"
INDEX = "index.json"


def is_packed_dir(path: str) -> bool:
    p = os.path.join(path, INDEX)
    if not os.path.isfile(p):
        return False
    try:
        with open(p) as fh:
            return json.load(fh).get("format") == FORMAT
    except (OSError, ValueError):
        return False


def read_index(path: str) -> List[dict]:
    """Batches of a packed directory: ``[{file, tints: [(contig, id), ...], reads}, ...]``."""
    with open(os.path.join(path, INDEX)) as fh:
        idx = json.load(fh)
    assert idx.get("format") == FORMAT, "not a packed SPLIT directory: %s" % path
    out = []
    for b in idx["batches"]:
        out.append(dict(file=os.path.join(path, b["file"]), tints=[(str(c), int(t)) for c, t in b["tints"]],
                        reads=int(b["reads"])))
    return out


def pack_directory(split_dir: str, packed_dir: str, threads: int = 1, batch_reads: int = 131072) -> dict:
    """Parses a SPLIT directory with the native parser and writes it as packed batches."""
    from .segment import list_tints
    split_dir = split_dir.rstrip("/")
    jobs = list_tints(split_dir)
    costs = [schedule.estimate_cost_from_files(split_dir, c, t) for c, t in jobs]
    os.makedirs(packed_dir, exist_ok=True)
    batches = []
    n_reads = 0
    for k, chunk in enumerate(schedule.batches(jobs, costs, batch_reads)):
        sp, rp, _, _ = hostio._paths(split_dir, "", chunk)
        pb = hostio.ParsedBatch(sp, rp, threads)
        try:
            name = "batch_%05d.frsb" % k
            pb.write_packed(os.path.join(packed_dir, name))
            batches.append(dict(file=name, tints=[[c, t] for c, t in chunk], reads=pb.n_reads))
            n_reads += pb.n_reads
        finally:
            pb.close()
    with open(os.path.join(packed_dir, INDEX), "w") as fh:
        json.dump(dict(format=FORMAT, source=os.path.abspath(split_dir), batches=batches), fh)
    return dict(batches=len(batches), tints=len(jobs), reads=n_reads)


def main(argv: Optional[List[str]] = None):
    ap = argparse.ArgumentParser(description="SPLIT directory -> packed batches for freddie_b200.segment")
    ap.add_argument("-s", "--split-dir", required=True)
    ap.add_argument("-o", "--outdir", required=True)
    ap.add_argument("-t", "--threads", type=int, default=1)
    ap.add_argument("--batch-reads", type=int, default=131072)
    a = ap.parse_args(argv)
    st = pack_directory(a.split_dir, a.outdir, a.threads, a.batch_reads)
    print("[freddie_b200.packed] {batches} batches, {tints} tints, {reads} reads".format(**st))


if __name__ == "__main__":
    main()
