"""Tint packing: variable-size tints -> one CSR batch (``frs_batch`` of ``include/freddie_b200.h``).

Python implementation used by the in-process seam (``freddie_b200.segment.segment``) and by tests; the
CLI uses the native parser in ``csrc/host_io.cpp`` which produces the same arrays straight from the
SPLIT files.  Mirrors what ``read_split`` builds (freddie_segment.py:121-171): read reps are the
distinct tuples of target intervals in first-seen order with weight = number of reads (:165-170);
islands are iterated ``s..e`` inclusive (:652-659); both ends of an interval must lie in one island
(:666-668).
"""
from __future__ import annotations

import bisect
import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from . import _lib

_OP = {"M": 0, "X": 0, "=": 0, "I": 1, "D": 2}

_DTYPES = dict(
    tint_island_off=np.int32, tint_rep_off=np.int32, tint_read_off=np.int32, island_start=np.int32,
    island_sample_off=np.int32, rep_iv_off=np.int32, rep_weight=np.int32, rep_iv_fs=np.int32, rep_iv_fe=np.int32,
    read_rep=np.int32, read_strand=np.uint8, read_len=np.int32, read_iv_off=np.int32, read_seq_off=np.int64,
    riv_ts=np.int32, riv_te=np.int32, riv_qs=np.int32, riv_qe=np.int32, riv_cig_off=np.int32, cigar=np.uint32,
    seq_is_a=np.uint32, seq_is_t=np.uint32,
)


class PackedBatch:
    """Flat arrays of a batch plus the per-tint host metadata the formatter needs."""

    def __init__(self, arrays: Dict[str, np.ndarray], tints: Sequence[dict]):
        self.arrays = arrays
        self.tints = list(tints)
        self.n_tints = len(arrays["tint_island_off"]) - 1
        self.n_reads = len(arrays["read_rep"])
        self.n_reps = len(arrays["rep_weight"])
        # a read's target intervals are its rep's: do not ship them twice (frs_batch.riv_ts / riv_te NULL)
        self.derive_riv = True

    def counts(self) -> Dict[str, int]:
        a = self.arrays
        return dict(
            n_tints=len(a["tint_island_off"]) - 1, n_islands=len(a["island_start"]), n_reps=len(a["rep_weight"]),
            n_rep_ivs=len(a["rep_iv_fs"]), n_reads=len(a["read_rep"]), n_read_ivs=len(a["riv_ts"]),
            n_cigar_ops=len(a["cigar"]), n_samples=int(a["island_sample_off"][-1]), n_seq_words=len(a["seq_is_a"]),
        )

    EDGE_WORDS = 16  # plane words per side in the edge store (512 bases: ~19 of 20 soft clips of ONT-like reads)

    def build_edge_store(self, words: int = None) -> np.ndarray:
        """``frs_batch.seq_edge``: the first and the last ``words`` plane words of every read, dense
        (``[read][side][plane][word]``, left-aligned, zero-padded).  The soft clips that the poly-A/T scans read
        sit at the two ends of a read; with this store the library copies them in one dense DMA."""
        E = int(words or self.EDGE_WORDS)
        a = self.arrays
        off = a["read_seq_off"].astype(np.int64)
        n = len(off) - 1
        out = np.zeros((n, 2, 2, E), dtype=np.uint32)
        if n and len(a["seq_is_a"]):
            k = np.arange(E, dtype=np.int64)[None, :]
            lo, hi = off[:-1, None], off[1:, None]
            first = lo + k                                  # words [0, E) of the read
            last = np.maximum(lo, hi - E) + k               # words [max(0, nw - E), nw), left-aligned
            top = len(a["seq_is_a"]) - 1
            for side, idx in ((0, first), (1, last)):
                ok = idx < hi
                src = np.minimum(idx, top)
                out[:, side, 0, :] = np.where(ok, a["seq_is_a"][src], 0)
                out[:, side, 1, :] = np.where(ok, a["seq_is_t"][src], 0)
        self.seq_edge = np.ascontiguousarray(out.reshape(-1))
        self.seq_edge_words = E
        self._struct_key = None
        return self.seq_edge

    def compact(self) -> "PackedBatch":
        """Adds the compact encodings ``frs_batch`` accepts for three per-interval arrays (a quarter of the bytes
        of a batch): CIGAR ops as uint16 when every length is < 4096, op counts per interval as uint8 instead
        of int32 offsets, and the query ends left out when they equal ``qs`` + the query bases the interval's
        CIGAR consumes.  Each one only when it is exact for this batch; the full arrays stay available."""
        a = self.arrays
        self.compact_arrays = {}
        cig = a["cigar"]
        if len(cig) == 0 or int(cig.max()) < (1 << 16):
            self.compact_arrays["cigar16"] = np.ascontiguousarray(cig.astype(np.uint16))
        n = np.diff(a["riv_cig_off"])
        if len(n) == 0 or int(n.max()) < 256:
            self.compact_arrays["riv_cig_n"] = np.ascontiguousarray(n.astype(np.uint8))
        ql = np.where((cig & 15) <= 1, cig >> 4, 0).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(ql)])
        off = a["riv_cig_off"].astype(np.int64)
        self.qe_from_cigar = bool(np.array_equal(a["riv_qs"].astype(np.int64) + csum[off[1:]] - csum[off[:-1]],
                                                 a["riv_qe"].astype(np.int64)))
        self._struct_key = None
        return self

    def as_struct(self) -> "_lib.FrsBatch":
        """The ``frs_batch`` view of the arrays (cached: building ~30 ctypes pointers costs more than enqueueing
        the batch; the cache is dropped when ``derive_riv`` or the arrays change identity)."""
        edge = getattr(self, "seq_edge", None)
        comp = getattr(self, "compact_arrays", None) or {}
        qe_c = bool(getattr(self, "qe_from_cigar", False)) and bool(comp)
        key = (self.derive_riv, id(edge), id(comp), qe_c, tuple(id(v) for v in self.arrays.values()))
        if getattr(self, "_struct_key", None) == key:
            return self._struct
        b = _lib.FrsBatch()
        for k, v in self.counts().items():
            setattr(b, k, v)
        for name in _lib.BATCH_ARRAYS:
            if self.derive_riv and name in ("riv_ts", "riv_te"):
                setattr(b, name, None)  # the library derives them from the rep intervals (no copy)
                continue
            if (name == "cigar" and "cigar16" in comp) or (name == "riv_cig_off" and "riv_cig_n" in comp) or (
                    name == "riv_qe" and qe_c):
                setattr(b, name, None)  # travels in its compact form
                continue
            arr = self.arrays[name]
            assert arr.dtype == _DTYPES[name] and arr.flags["C_CONTIGUOUS"], name
            setattr(b, name, arr.ctypes.data_as(C.c_void_p))
        if edge is not None and len(edge):
            b.seq_edge_words = self.seq_edge_words
            b.seq_edge = edge.ctypes.data_as(C.c_void_p)
        for name, arr in comp.items():
            setattr(b, name, arr.ctypes.data_as(C.c_void_p))
        b.qe_from_cigar = int(qe_c)
        b.host_arena = int(bool(getattr(self, "host_arena", False)))
        self._struct, self._struct_key = b, key
        return b

    def nbytes(self) -> int:
        return int(sum(v.nbytes for v in self.arrays.values()))

    # order in which frs_upload copies the arrays of a batch (frs.cu: stage_upload); arrays laid out in this
    # order inside one pinned allocation, each at the next multiple of 256 bytes, cross the bus as one copy
    UPLOAD_ORDER = ["tint_island_off", "tint_rep_off", "tint_read_off", "island_start", "island_sample_off",
                    "rep_iv_off", "rep_weight", "rep_iv_fs", "rep_iv_fe", "read_rep", "read_strand", "read_len",
                    "read_iv_off", "read_seq_off", "riv_ts", "riv_te", "riv_qs", "cigar", "riv_cig_off", "riv_qe"]

    def pin(self, edge_words: int = None, compact: bool = True):
        """Moves the arrays into page-locked host memory (torch is used for buffer management only), adds
        the edge store of the sequence planes (``build_edge_store``; ``edge_words=0``: none) and the compact
        encodings of ``compact()``.  Everything ``frs_upload`` copies is placed in ONE pinned allocation in the
        library's upload order (``frs_batch.host_arena``): the arrays that travel first, in order, then the
        ones that stay behind (full forms of compacted arrays, derivable arrays)."""
        import torch
        self._pinned = {}
        if compact:
            self.compact()
        comp = getattr(self, "compact_arrays", None) or {}
        qe_c = bool(getattr(self, "qe_from_cigar", False)) and bool(comp)
        edge = self.build_edge_store(edge_words) if (edge_words is None or edge_words > 0) else None
        sent, kept = [], []
        for name in self.UPLOAD_ORDER:
            if self.derive_riv and name in ("riv_ts", "riv_te"):
                kept.append(("a", name))
            elif name == "cigar" and "cigar16" in comp:
                sent.append(("c", "cigar16"))
                kept.append(("a", name))
            elif name == "riv_cig_off" and "riv_cig_n" in comp:
                sent.append(("c", "riv_cig_n"))
                kept.append(("a", name))
            elif name == "riv_qe" and qe_c:
                kept.append(("a", name))
            else:
                sent.append(("a", name))
        if edge is not None:
            sent.append(("e", "seq_edge"))
        items = sent + kept
        get = lambda kind, name: (self.arrays if kind == "a" else comp)[name] if kind != "e" else edge  # noqa: E731
        al = lambda n: (max(int(n), 16) + 255) & ~255  # noqa: E731
        total = sum(al(get(k, n).nbytes) for k, n in items)
        arena = torch.empty(total + 256, dtype=torch.uint8).pin_memory()
        self._pinned["arena"] = arena
        base = arena.numpy()
        skew = (-base.ctypes.data) % 256  # the allocation is page-aligned in practice; keep the rule anyway
        off = skew
        for kind, name in items:
            v = get(kind, name)
            view = base[off:off + v.nbytes].view(v.dtype)
            view[...] = v
            if kind == "a":
                self.arrays[name] = view
            elif kind == "c":
                comp[name] = view
            else:
                self.seq_edge = view
            off += al(v.nbytes)
        for k in ("seq_is_a", "seq_is_t"):  # the planes stay where they are read from on demand
            v = self.arrays[k]
            t = torch.from_numpy(v if v.size else np.zeros(1, dtype=v.dtype)).pin_memory()
            self._pinned[k] = t
            self.arrays[k] = t.numpy()[: v.size] if v.size else t.numpy()[:0]
        self.host_arena = True
        self._struct_key = None
        return self


def seq_planes(seq: str):
    """isA / isT bit-planes of a read, 32 bases per little-endian word."""
    a = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    n_words = (len(a) + 31) // 32
    out = []
    for ch in (65, 84):
        bits = np.packbits(a == ch, bitorder="little")
        buf = np.zeros(n_words * 4, dtype=np.uint8)
        buf[: len(bits)] = bits
        out.append(buf.view(np.uint32))
    return out[0], out[1]


def pack_tints(tints: Sequence[dict]) -> PackedBatch:
    L: Dict[str, list] = {k: [] for k in _DTYPES}
    L["tint_island_off"].append(0)
    L["tint_rep_off"].append(0)
    L["tint_read_off"].append(0)
    L["island_sample_off"].append(0)
    L["rep_iv_off"].append(0)
    L["read_iv_off"].append(0)
    L["read_seq_off"].append(0)
    L["riv_cig_off"].append(0)
    n_isl = n_rep = n_read = 0
    n_smp = n_rep_iv = n_riv = n_cig = n_words = 0
    planes_a: List[np.ndarray] = []
    planes_t: List[np.ndarray] = []
    for tint in tints:
        islands = tint["intervals"]
        assert all(a[1] < b[0] for a, b in zip(islands[:-1], islands[1:])), islands  # :138
        assert all(s < e for s, e in islands)  # :140
        starts = [s for s, _ in islands]
        isl_off = []
        for s, e in islands:
            isl_off.append(n_smp)
            L["island_start"].append(s)
            n_smp += e - s + 1
            L["island_sample_off"].append(n_smp)
        n_isl += len(islands)
        # read reps in first-seen order (:165-170)
        rep_index: Dict[tuple, int] = {}
        rep_base = n_rep
        for read in tint["reads"]:
            key = tuple((iv[0], iv[1]) for iv in read["intervals"])
            j = rep_index.get(key)
            if j is None:
                j = len(rep_index)
                rep_index[key] = j
                L["rep_weight"].append(0)
                for ts, te in key:
                    a = bisect.bisect_right(starts, ts) - 1
                    if a < 0 or ts > islands[a][1]:
                        raise KeyError(ts)  # pos_to_Yy_idx[ts] (:666)
                    if te > islands[a][1]:
                        b = bisect.bisect_right(starts, te) - 1
                        if te > islands[b][1]:
                            raise KeyError(te)  # pos_to_Yy_idx[te] (:667)
                        raise AssertionError((a, b))  # assert Y_idx_s == Y_idx_e (:668)
                    L["rep_iv_fs"].append(isl_off[a] + ts - islands[a][0])
                    L["rep_iv_fe"].append(isl_off[a] + te - islands[a][0])
                n_rep_iv += len(key)
                L["rep_iv_off"].append(n_rep_iv)
            L["rep_weight"][rep_base + j] += 1
            L["read_rep"].append(rep_base + j)
            L["read_strand"].append(0 if read["strand"] == "+" else 1)
            seq = read["seq"]
            L["read_len"].append(len(seq))
            for (ts, te, qs, qe, cig) in read["intervals"]:
                L["riv_ts"].append(ts)
                L["riv_te"].append(te)
                L["riv_qs"].append(qs)
                L["riv_qe"].append(qe)
                for cnt, op in cig:
                    L["cigar"].append((cnt << 4) | _OP.get(op, 3))
                n_cig += len(cig)
                L["riv_cig_off"].append(n_cig)
            n_riv += len(read["intervals"])
            L["read_iv_off"].append(n_riv)
            pa, pt = seq_planes(seq)
            planes_a.append(pa)
            planes_t.append(pt)
            n_words += len(pa)
            L["read_seq_off"].append(n_words)
        n_rep += len(rep_index)
        n_read += len(tint["reads"])
        L["tint_island_off"].append(n_isl)
        L["tint_rep_off"].append(n_rep)
        L["tint_read_off"].append(n_read)
    assert n_smp < 2 ** 31 and n_rep_iv < 2 ** 31 and n_cig < 2 ** 31, "batch too large for int32 offsets"
    arrays = {}
    for k, dt in _DTYPES.items():
        if k == "seq_is_a":
            arrays[k] = np.concatenate(planes_a) if planes_a else np.zeros(0, dtype=np.uint32)
        elif k == "seq_is_t":
            arrays[k] = np.concatenate(planes_t) if planes_t else np.zeros(0, dtype=np.uint32)
        else:
            arrays[k] = np.ascontiguousarray(np.array(L[k], dtype=dt))
    return PackedBatch(arrays, tints)
