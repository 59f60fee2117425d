"""Per-GPU engine: owns one ``frs_context`` and runs packed batches through the CUDA pipeline.

Host code only manages buffers (numpy / torch pinned memory) and calls the C ABI; every step of the
hot path runs in the hand-written kernels of ``csrc/``.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .pack import PackedBatch


def smooth_threshold(threshold: float) -> List[float]:
    """Per-length threshold table, same arithmetic as the reference's ``smooth_threshold``
    (freddie_segment.py:277-286); computed on the host in Python so the bits are the reference's."""
    smooth: List[float] = []
    x = 0
    while True:
        y = threshold / (1 + ((threshold - .5) / .5) * math.exp(-0.05 * x))
        if x > 5 and x * (threshold - y) < 0.5:
            return smooth
        smooth.append(round(y, 2))
        x += 1
        assert x < 1000


def gaussian_kernel(sigma: float, truncate: float) -> np.ndarray:
    """scipy's ``_gaussian_kernel1d`` for order 0 (``_filters.py:656-666``, radius ``int(truncate*sd+.5)``,
    ``:745``), evaluated with the same numpy expressions so the weights are bit-identical."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-lw, lw + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    return np.ascontiguousarray(phi_x / phi_x.sum())


class SegmentParams:
    """Flags of the stage (``parse_args``, freddie_segment.py:53-110) with its asserts (:104-109)."""

    def __init__(self, sigma=5.0, threshold_rate=0.90, variance_factor=3.0, max_problem_size=50,
                 min_read_support_outside=3, ignore_ends=True, smoothed_threshold: Optional[Sequence[float]] = None):
        assert 1 >= threshold_rate >= 0.5
        assert 10 > variance_factor > 0
        assert 50 >= sigma > 0
        assert max_problem_size > 3
        assert min_read_support_outside >= 0
        self.sigma = float(sigma)
        self.tp = float(threshold_rate)
        self.vf = float(variance_factor)
        self.mps = int(max_problem_size)
        self.lo = int(min_read_support_outside)
        self.ignore_ends = bool(ignore_ends)
        tbl = smooth_threshold(self.tp) if smoothed_threshold is None else list(smoothed_threshold)
        self.table = np.ascontiguousarray(np.array(tbl if len(tbl) else [self.tp], dtype=np.float64))
        self.table_len = len(tbl)
        self.gw = gaussian_kernel(self.sigma, 4.0)
        self.rw = gaussian_kernel(self.sigma, 1.0)

    def as_struct(self) -> "_lib.FrsParams":
        if getattr(self, "_struct", None) is not None:
            return self._struct
        p = _lib.FrsParams()
        p.sigma, p.tp, p.vf = self.sigma, self.tp, self.vf
        p.mps, p.lo, p.ignore_ends = self.mps, self.lo, int(self.ignore_ends)
        p.thr_table_len = self.table_len
        p.thr_table = self.table.ctypes.data_as(C.c_void_p)
        p.gauss_w = self.gw.ctypes.data_as(C.c_void_p)
        p.refine_w = self.rw.ctypes.data_as(C.c_void_p)
        p.gauss_radius = (len(self.gw) - 1) // 2
        p.refine_radius = (len(self.rw) - 1) // 2
        self._struct = p
        return p


class BatchResult:
    """Flat result arrays of one batch (``frs_result``)."""

    def __init__(self, sizes: "_lib.FrsResultSizes", n_tints: int, n_reads: int, pinned: bool = False):
        self.sizes = {f[0]: getattr(sizes, f[0]) for f in sizes._fields_}
        shapes = dict(
            tint_final_off=(n_tints + 1, np.int32), final_pos=(sizes.n_final, np.int32),
            tint_digit_off=(n_tints + 1, np.int64), digits=(sizes.n_digit_bytes, np.uint8),
            read_head=(n_reads * 8, np.int32), read_gap_off=(n_reads + 1, np.int32),
            gap_rec=(sizes.n_gap_records * 3, np.int32),
        )
        self.arrays: Dict[str, np.ndarray] = {}
        self._keep = []
        for k, (n, dt) in shapes.items():
            if pinned:
                import torch
                t = torch.empty(max(int(n), 1), dtype=getattr(torch, np.dtype(dt).name)).pin_memory()
                self._keep.append(t)
                self.arrays[k] = t.numpy()[: int(n)]
            else:
                self.arrays[k] = np.empty(int(n), dtype=dt)

    def as_struct(self) -> "_lib.FrsResult":
        if getattr(self, "_struct", None) is None:
            r = _lib.FrsResult()
            for k in _lib.RESULT_ARRAYS:
                setattr(r, k, self.arrays[k].ctypes.data_as(C.c_void_p))
            self._struct = r
        return self._struct


class Engine:
    """One CUDA context of the library on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        rc = self.lib.frs_create(device, C.byref(self.ctx))
        if rc != 0:
            raise _lib.FrsError(rc, (self.lib.frs_last_error(None) or b"").decode())
        self.device = device
        self._batch: Optional[PackedBatch] = None

    def close(self):
        if self.ctx:
            self.lib.frs_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = (self.lib.frs_last_error(self.ctx) or b"").decode()
            if msg.startswith("AssertionError"):
                raise AssertionError("[libfreddie_b200 %d] %s" % (rc, msg))
            raise _lib.FrsError(rc, msg)

    # -- the three steps of a batch ---------------------------------------------------------
    def upload(self, batch: PackedBatch):
        b = batch.as_struct()
        self._check(self.lib.frs_upload(self.ctx, C.byref(b)))
        self._batch = batch

    def run(self, prm: SegmentParams) -> "_lib.FrsResultSizes":
        sizes = _lib.FrsResultSizes()
        p = prm.as_struct()
        self._check(self.lib.frs_run(self.ctx, C.byref(p), C.byref(sizes)))
        return sizes

    def download(self, sizes, pinned: bool = False) -> BatchResult:
        res = BatchResult(sizes, self._batch.n_tints, self._batch.n_reads, pinned)
        r = res.as_struct()
        self._check(self.lib.frs_download(self.ctx, C.byref(r)))
        return res

    def segment_batch(self, batch: PackedBatch, prm: SegmentParams, pinned: bool = False) -> BatchResult:
        """upload + run + download: host buffers in, host buffers out."""
        self.upload(batch)
        sizes = self.run(prm)
        return self.download(sizes, pinned)

    # -- pipelined form: two batches in flight, one host thread -------------------------------
    def submit(self, batch: PackedBatch, prm: SegmentParams) -> int:
        """Enqueues the copies and every kernel of ``batch`` and returns a ticket at once; the batch's
        arrays must stay alive until ``wait`` has returned for that ticket."""
        b = batch.as_struct()
        p = prm.as_struct()
        t = C.c_int(-1)
        self._check(self.lib.frs_submit(self.ctx, C.byref(b), C.byref(p), C.byref(t)))
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (batch, prm)
        self._batch = batch
        return t.value

    def wait(self, ticket: int) -> "_lib.FrsResultSizes":
        sizes = _lib.FrsResultSizes()
        self._check(self.lib.frs_wait(self.ctx, ticket, C.byref(sizes)))
        return sizes

    def fetch(self, ticket: int, res: "BatchResult") -> "BatchResult":
        """Copies the results of ``ticket`` into ``res`` (sized from ``wait``'s answer) and frees the ticket."""
        r = res.as_struct()
        try:
            self._check(self.lib.frs_fetch(self.ctx, ticket, C.byref(r)))
        finally:
            self._inflight.pop(ticket, None)
        return res

    def fetch_start(self, ticket: int, res: "BatchResult") -> "BatchResult":
        """First half of ``fetch``: enqueues the copies into ``res`` and returns at once."""
        r = res.as_struct()
        self._check(self.lib.frs_fetch_start(self.ctx, ticket, C.byref(r)))
        return res

    def fetch_finish(self, ticket: int) -> None:
        """Second half: blocks until the copies of ``fetch_start`` have arrived; frees the ticket."""
        try:
            self._check(self.lib.frs_fetch_finish(self.ctx, ticket))
        finally:
            self._inflight.pop(ticket, None)

    def new_result(self, sizes, batch: PackedBatch, pinned: bool = False) -> "BatchResult":
        return BatchResult(sizes, batch.n_tints, batch.n_reads, pinned)

    # -- instrumentation ----------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self.lib.frs_set_profiling(self.ctx, int(on))

    def timings(self):
        names = (C.c_char_p * _lib.FRS_MAX_STAGES)()
        ms = (C.c_float * _lib.FRS_MAX_STAGES)()
        ln = (C.c_int * _lib.FRS_MAX_STAGES)()
        n = self.lib.frs_get_timings(self.ctx, names, ms, ln)
        return [(names[i].decode(), float(ms[i]), int(ln[i])) for i in range(n)]

    def set_option(self, key: int, value: int):
        self._check(self.lib.frs_set_option(self.ctx, int(key), int(value)))

    def stats(self) -> Dict[str, int]:
        """Transfer statistics of the last upload / run (``frs_get_stats``)."""
        buf = (C.c_longlong * len(_lib.STAT_NAMES))()
        self.lib.frs_get_stats(self.ctx, buf, len(_lib.STAT_NAMES))
        return {k: int(buf[i]) for i, k in enumerate(_lib.STAT_NAMES)}

    def launch_count(self) -> int:
        return int(self.lib.frs_last_launch_count(self.ctx))

    def tap(self, which: int, dtype) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self.lib.frs_get_intermediate(self.ctx, which, None, 0, C.byref(n)))
        out = np.empty(n.value // np.dtype(dtype).itemsize, dtype=dtype)
        if n.value:
            self._check(self.lib.frs_get_intermediate(self.ctx, which, out.ctypes.data_as(C.c_void_p), n.value,
                                                      C.byref(n)))
        return out


# ------------------------------------------------------------------------------------------------
# results -> the reference's per-tint / per-read fields
# ------------------------------------------------------------------------------------------------
def gap_strings(head: np.ndarray, recs: np.ndarray) -> List[str]:
    """Gap strings of one read in the reference's order (``sorted(read['gaps'])``, :472)."""
    flags = int(head[0])
    if not flags & 1:
        return []
    out = set()
    sk = (flags >> 8) & 3
    if sk:
        out.add("S%s_%d:%d" % ("AT"[sk - 1], head[1], head[2]))
    out.add("SSC:%d" % head[3])
    ek = (flags >> 16) & 3
    if ek:
        out.add("E%s_%d:%d" % ("AT"[ek - 1], head[4], head[5]))
    out.add("ESC:%d" % head[6])
    for k in range(0, len(recs), 3):
        out.add("%d-%d:%d" % (recs[k], recs[k + 1], recs[k + 2]))
    return sorted(out)


def apply_result(batch: PackedBatch, res: BatchResult) -> None:
    """Writes ``final_positions``, ``segs``, ``read['data']`` and ``read['gaps']`` into the tint
    dicts of the batch, exactly the fields ``segment`` mutates in the reference (:738-844)."""
    a = res.arrays
    ba = batch.arrays
    for t, tint in enumerate(batch.tints):
        f0, f1 = int(a["tint_final_off"][t]), int(a["tint_final_off"][t + 1])
        fp = a["final_pos"][f0:f1].tolist()
        tint["final_positions"] = fp
        tint["segs"] = list(zip(fp[:-1], fp[1:]))
        S = f1 - f0 - 1
        d0 = int(a["tint_digit_off"][t])
        rep0 = int(ba["tint_rep_off"][t])
        r0 = int(ba["tint_read_off"][t])
        rows = {}
        for k, read in enumerate(tint["reads"]):
            i = r0 + k
            rep = int(ba["read_rep"][i]) - rep0
            row = rows.get(rep)
            if row is None:
                row = (a["digits"][d0 + rep * S: d0 + (rep + 1) * S] - 48).tolist()
                rows[rep] = row
            read["data"] = list(row)
            g0, g1 = int(a["read_gap_off"][i]), int(a["read_gap_off"][i + 1])
            read["gaps"] = gap_strings(a["read_head"][i * 8: i * 8 + 8], a["gap_rec"][g0 * 3: g1 * 3])


def format_tint(batch: PackedBatch, res: BatchResult, t: int) -> str:
    """SEGMENT text of tint ``t`` (``run_segment``, :715-731) straight from the flat arrays."""
    a = res.arrays
    ba = batch.arrays
    tint = batch.tints[t]
    f0, f1 = int(a["tint_final_off"][t]), int(a["tint_final_off"][t + 1])
    S = f1 - f0 - 1
    out = ["#%s\t%d\t%s\n" % (tint["chr"], tint["id"], ",".join(map(str, a["final_pos"][f0:f1].tolist())))]
    d0 = int(a["tint_digit_off"][t])
    rep0 = int(ba["tint_rep_off"][t])
    r0 = int(ba["tint_read_off"][t])
    dig = a["digits"]
    for k, read in enumerate(tint["reads"]):
        i = r0 + k
        rep = int(ba["read_rep"][i]) - rep0
        row = dig[d0 + rep * S: d0 + (rep + 1) * S].tobytes().decode("ascii")
        g0, g1 = int(a["read_gap_off"][i]), int(a["read_gap_off"][i + 1])
        gaps = gap_strings(a["read_head"][i * 8: i * 8 + 8], a["gap_rec"][g0 * 3: g1 * 3])
        out.append("%d\t%s\t%s\t%s\t%d\t%s\t%s\n" % (read["id"], read["name"], read["chr"], read["strand"],
                                                      read["tint"], row, "".join("%s," % g for g in gaps)))
    return "".join(out)
