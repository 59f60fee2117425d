"""Host-side mirror of the tint construction of ``freddie_split.py`` over the CUDA path
``frs_split_*`` (``include/freddie_b200.h``, ``csrc/split.cu``, ``csrc/kernels_split.cuh``).

* ``get_transcriptional_intervals(reads)`` has the reference's signature (freddie_split.py:295-364): ``reads`` is the
  list ``read_sam`` yields (dicts with ``id`` = position and ``intervals`` = [(ts, te, qs, qe, cigar), ...]); it returns
  the list of ``dict(intervals=[(s, e), ...], rids=[...])`` in the reference's order, groups with >= 100 intervals or
  >= 1500 reads already broken up like ``break_tint`` (:246-293) does;
* ``SplitTints.run`` is the batch form: many groups (e.g. every group of a contig) in ONE device run.

BAM decoding (pysam) stays with the caller.  There is no CPU fallback: without the CUDA library or a device every
entry point raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


class SplitTints:
    """One split-stage context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        rc = self.lib.frs_split_create(device, C.byref(self.ctx))
        if rc != 0:
            raise _lib.FrsError(rc, "frs_split_create: no usable CUDA device %d; there is no CPU fallback" % device)

    def close(self):
        if self.ctx:
            self.lib.frs_split_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, groups: Sequence[Sequence[Sequence[Tuple[int, int]]]], max_intervals: int = 100, max_reads: int = 1500):
        """``groups[g][r]`` = the (start, end) target intervals of read r of group g.  Returns, per group, the list of
        (intervals, rids) tints, plus the sizes / timing of the run."""
        gro, rio, s, e = [0], [0], [], []
        for reads in groups:
            for ivs in reads:
                for iv in ivs:
                    s.append(iv[0])
                    e.append(iv[1])
                rio.append(len(s))
            gro.append(len(rio) - 1)
        return self.run_arrays(np.asarray(gro, np.int32), np.asarray(rio, np.int32), np.asarray(s, np.int32),
                               np.asarray(e, np.int32), max_intervals, max_reads)

    def run_arrays(self, group_read_off, read_iv_off, iv_s, iv_e, max_intervals: int = 100, max_reads: int = 1500):
        keep = [np.ascontiguousarray(a, dtype=np.int32) for a in (group_read_off, read_iv_off, iv_s, iv_e)]
        G, N = len(keep[0]) - 1, len(keep[1]) - 1
        b = _lib.FrsSplitBatch(n_groups=G, n_reads=N, group_read_off=keep[0].ctypes.data, read_iv_off=keep[1].ctypes.data,
                               iv_s=keep[2].ctypes.data, iv_e=keep[3].ctypes.data, max_intervals=max_intervals, max_reads=max_reads)
        z = _lib.FrsSplitSizes()
        rc = self.lib.frs_split_run(self.ctx, C.byref(b), C.byref(z))
        if rc != 0:
            raise _lib.FrsError(rc, (self.lib.frs_split_last_error(self.ctx) or b"").decode())
        T = int(z.n_tints)
        out = dict(group_tint_off=np.zeros(G + 1, np.int32), tint_iv_off=np.zeros(T + 1, np.int32),
                   tint_iv_s=np.zeros(int(z.n_tint_ivs), np.int32), tint_iv_e=np.zeros(int(z.n_tint_ivs), np.int32),
                   tint_rid_off=np.zeros(T + 1, np.int32), tint_rids=np.zeros(int(z.n_tint_rids), np.int32))
        r = _lib.FrsSplitResult(**{n: out[n].ctypes.data for n in _lib.SPLIT_RESULT_ARRAYS})
        rc = self.lib.frs_split_fetch(self.ctx, C.byref(r))
        if rc != 0:
            raise _lib.FrsError(rc, (self.lib.frs_split_last_error(self.ctx) or b"").decode())
        tints = []
        for g in range(G):
            lst = []
            for t in range(int(out["group_tint_off"][g]), int(out["group_tint_off"][g + 1])):
                a, bb = int(out["tint_iv_off"][t]), int(out["tint_iv_off"][t + 1])
                ra, rb = int(out["tint_rid_off"][t]), int(out["tint_rid_off"][t + 1])
                lst.append((list(zip(out["tint_iv_s"][a:bb].tolist(), out["tint_iv_e"][a:bb].tolist())),
                            out["tint_rids"][ra:rb].tolist()))
            tints.append(lst)
        info = {n: int(getattr(z, n)) for n, _ in z._fields_}
        info["device_ms"] = float(self.lib.frs_split_last_ms(self.ctx))
        return tints, info, out


_ctx: Optional[SplitTints] = None


def get_transcriptional_intervals(reads: List[dict]) -> List[dict]:
    """Drop-in for ``freddie_split.get_transcriptional_intervals`` (:295-364)."""
    global _ctx
    if _ctx is None:
        _ctx = SplitTints(0)
    assert all(read["id"] == i for i, read in enumerate(reads)), "read ids are positions in the list (read_sam, :212)"
    tints, _, _ = _ctx.run([[[(iv[0], iv[1]) for iv in read["intervals"]] for read in reads]])
    return [dict(intervals=ivs, rids=rids) for ivs, rids in tints[0]]
