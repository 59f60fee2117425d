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


# ---- the SEGMENT twin -------------------------------------------------------------------------------
_SEG_SECTIONS = [  # name, dtype -- order of frs_packed_write_segment
    ("tint_id", "<i8"), ("tint_chr_off", "<u8"), ("tint_chr", "u1"), ("tint_read_off", "<i4"), ("tint_rep_off", "<i4"),
    ("tint_final_off", "<i4"), ("final_pos", "<i4"), ("tint_digit_off", "<i8"), ("digits", "u1"), ("read_rep", "<i4"),
    ("read_rid", "<i8"), ("read_tint", "<i8"), ("name_off", "<u8"), ("names", "u1"), ("chr_off", "<u8"), ("chrs", "u1"),
    ("read_strand", "u1"), ("read_head", "<i4"), ("read_gap_off", "<i4"), ("gap_rec", "<i4"),
]


class PackedSegment:
    """Reader of a packed SEGMENT batch ("FRSSEGM1", written with ``--packed-segment``): the arrays the
    SEGMENT rows are printed from, memory-mapped.  ``text(t)`` is byte-identical to ``segment_<chr>_<id>.tsv``;
    ``read_segment()`` builds what ``freddie_cluster.read_segment`` (freddie_cluster.py:119-172) builds
    from the text, without the regexes."""

    def __init__(self, path: str):
        import numpy as np
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        if raw.size < 16 or bytes(raw[:8]) != b"FRSSEGM1":
            raise ValueError("not a packed SEGMENT batch (FRSSEGM1): %s" % path)
        ns = int(np.frombuffer(raw[8:16], dtype="<u8")[0])
        if ns != len(_SEG_SECTIONS) or raw.size < 16 + 16 * ns:
            raise ValueError("not a packed SEGMENT batch (FRSSEGM1): %s" % path)
        table = np.frombuffer(raw[16:16 + 16 * ns], dtype="<u8").reshape(ns, 2)
        self.a = {}
        for (name, dt), (off, nbytes) in zip(_SEG_SECTIONS, table.tolist()):
            if off + nbytes > raw.size or nbytes % np.dtype(dt).itemsize:
                raise ValueError("damaged packed SEGMENT batch: %s" % path)
            self.a[name] = np.frombuffer(raw[off:off + nbytes], dtype=dt)
        self.n_tints = len(self.a["tint_id"])
        self.n_reads = len(self.a["read_rep"])
        a = self.a
        ok = (len(a["tint_read_off"]) == self.n_tints + 1 and len(a["tint_final_off"]) == self.n_tints + 1
              and len(a["read_head"]) == 8 * self.n_reads and len(a["read_gap_off"]) == self.n_reads + 1
              and int(a["tint_read_off"][-1]) == self.n_reads and len(a["final_pos"]) == int(a["tint_final_off"][-1])
              and len(a["digits"]) == int(a["tint_digit_off"][-1]) and len(a["gap_rec"]) == 3 * int(a["read_gap_off"][-1]))
        if not ok:
            raise ValueError("damaged packed SEGMENT batch: %s" % path)

    def tints(self) -> List[Tuple[str, int]]:
        a = self.a
        return [(bytes(a["tint_chr"][int(a["tint_chr_off"][t]):int(a["tint_chr_off"][t + 1])]).decode(), int(a["tint_id"][t]))
                for t in range(self.n_tints)]

    def _rows(self, t: int):
        """(rid, name, chr, strand, tint column, digits row as bytes, head[8], gap records [k, 3]) per read of tint t."""
        a = self.a
        f0, f1 = int(a["tint_final_off"][t]), int(a["tint_final_off"][t + 1])
        S = f1 - f0 - 1
        d0 = int(a["tint_digit_off"][t])
        rep0 = int(a["tint_rep_off"][t])
        for i in range(int(a["tint_read_off"][t]), int(a["tint_read_off"][t + 1])):
            rep = int(a["read_rep"][i]) - rep0
            g0, g1 = int(a["read_gap_off"][i]), int(a["read_gap_off"][i + 1])
            yield (int(a["read_rid"][i]),
                   bytes(a["names"][int(a["name_off"][i]):int(a["name_off"][i + 1])]).decode(),
                   bytes(a["chrs"][int(a["chr_off"][i]):int(a["chr_off"][i + 1])]).decode(),
                   "-" if a["read_strand"][i] else "+", int(a["read_tint"][i]),
                   bytes(a["digits"][d0 + rep * S:d0 + (rep + 1) * S]),
                   a["read_head"][8 * i:8 * i + 8], a["gap_rec"][3 * g0:3 * g1].reshape(-1, 3))

    def text(self, t: int) -> str:
        """The SEGMENT file of tint ``t`` (freddie_segment.py:715-731)."""
        from .engine import gap_strings
        a = self.a
        chrom, tid = self.tints()[t]
        f0, f1 = int(a["tint_final_off"][t]), int(a["tint_final_off"][t + 1])
        out = ["#%s\t%d\t%s\n" % (chrom, tid, ",".join(map(str, a["final_pos"][f0:f1].tolist())))]
        for rid, name, rchr, strand, tcol, row, head, recs in self._rows(t):
            gaps = gap_strings(head, recs.reshape(-1))
            out.append("%d\t%s\t%s\t%s\t%d\t%s\t%s\n" % (rid, name, rchr, strand, tcol, row.decode("ascii"),
                                                           "".join("%s," % g for g in gaps)))
        return "".join(out)

    def read_segment(self) -> dict:
        """``{tint id: tint}`` with the fields ``freddie_cluster.read_segment`` fills (freddie_cluster.py:119-172):
        ``segs`` as (start, end, length), ``reads`` with ``data`` / ``gaps`` / ``softclip`` / ``poly_tail``, and
        ``read_reps`` = lists of read indices that share a key of digits (2 -> 0), gap sizes and poly lengths
        (sizes <= 10 count as 0), in first-seen order."""
        a = self.a
        tints = {}
        for t, (chrom, tid) in enumerate(self.tints()):
            f0, f1 = int(a["tint_final_off"][t]), int(a["tint_final_off"][t + 1])
            pos = a["final_pos"][f0:f1].tolist()
            assert all(x < y for x, y in zip(pos[:-1], pos[1:])), pos
            assert tid not in tints, "Transcriptional interval with id {} is repeated!".format(tid)
            tint = dict(id=tid, chr=chrom, segs=[(s, e, e - s) for s, e in zip(pos[:-1], pos[1:])], read_reps=dict(),
                        reads=list())
            tints[tid] = tint
            for rid, name, rchr, strand, tcol, row, head, recs in self._rows(t):
                flags = int(head[0])
                has = bool(flags & 1)
                # the text lists the gap strings sorted; the dicts below do not depend on that order, the
                # read-rep key does (internal gaps first in sorted order, then poly tails in sorted order)
                gap_items = sorted({(int(r[0]), int(r[1]), int(r[2])) for r in recs.tolist()},
                                   key=lambda g: "%d-%d:%d" % g) if has else []
                poly = []
                if has and (flags >> 8) & 3:
                    poly.append(("S" + "AT"[((flags >> 8) & 3) - 1], int(head[1]), int(head[2])))
                if has and (flags >> 16) & 3:
                    poly.append(("E" + "AT"[((flags >> 16) & 3) - 1], int(head[4]), int(head[5])))
                poly.sort(key=lambda p: "%s_%d:%d" % p)
                read = dict(id=rid, name=name, chr=rchr, strand=strand, tint=tcol, data=[c - 48 for c in row],
                            gaps={(g[0], g[1]): g[2] for g in gap_items},
                            softclip=({"SSC": int(head[3]), "ESC": int(head[6])} if has else {}),
                            poly_tail={p[0]: (p[1], p[2]) for p in poly})
                key = row.replace(b"2", b"0").decode("ascii")
                key += "".join(".{}".format(g[2] if g[2] > 10 else 0) for g in gap_items)
                key += "".join(".{}{}".format(p[0][0], p[2] if p[2] > 10 else 0) for p in poly)
                tint["reads"].append(read)
                tint["read_reps"].setdefault(key, []).append(len(tint["reads"]) - 1)
                assert len(read["data"]) == len(tint["segs"]), (read["data"], tint["segs"])
                assert read["chr"] == tint["chr"]
                assert all(0 <= g[0] < g[1] < len(read["data"]) for g in read["gaps"])
        for tint in tints.values():
            tint["read_reps"] = list(tint["read_reps"].values())
        return tints


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
