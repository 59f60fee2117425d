"""Host-side mirror of the Gurobi-free front of ``freddie_cluster.py`` over the CUDA path
``frs_cprep_*`` (``include/freddie_b200.h``, ``csrc/cluster.cu``, ``csrc/kernels_cluster.cuh``).

Same names and argument meaning as the reference for this path:

* ``preprocess_ilp(tint, ilp_settings)``      (freddie_cluster.py:277-328) fills ``tint['ilp_data']`` (``I``, ``C``,
  ``FL``, ``garbage_cost``), ``read['poly_tail_category']`` and the added tail gap in ``read['gaps']``;
* ``partition_reads(tint, maximum_ilp_size)`` (:198-274) fills ``tint['partitions']``;
* ``ClusterPrep.run`` is the batch form both are built on: many tints at once, straight from the arrays the
  segment stage leaves (``FRSSEGM1`` / ``frs_result``), including ``read_segment``'s read-rep merge (:154-164).

There is no CPU fallback: without the CUDA library or a device every entry point raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib


class ClusterPrepResult:
    """Arrays of ``frs_cluster_result`` plus per-tint views in the reference's shapes."""

    def __init__(self, arrays: Dict[str, np.ndarray], sizes: "_lib.FrsClusterSizes", batch: Dict[str, np.ndarray], timings):
        self.a = arrays
        self.batch = batch
        self.sizes = {n: int(getattr(sizes, n)) for n, _ in sizes._fields_}
        self.timings_ms = dict(zip(["dedupe_preprocess", "pair_test", "prune", "components", "incompatible_pairs"], timings))
        self.n_tints = len(arrays["tint_rep_off"]) - 1

    def tint(self, t: int, recycle_model: str = "constant") -> dict:
        """``I``, ``C`` (U x M uint8), ``FL`` (U x 2), ``cat``, ``garbage_cost``, ``gaps`` (dict per rep, the first
        read's gaps plus the added tail gap), ``read_reps`` (read indices per rep, first-seen order) and
        ``partitions`` [(rids, [(rid_1, rid_2), ...])] of tint ``t``."""
        a, b = self.a, self.batch
        u0, u1 = int(a["tint_rep_off"][t]), int(a["tint_rep_off"][t + 1])
        U = u1 - u0
        M = int(b["tint_seg_n"][t])
        o = int(a["tint_row_off"][t])
        I = a["I"][o:o + U * M].reshape(U, M)
        Cm = a["C"][o:o + U * M].reshape(U, M)
        FL = a["rep_fl"][2 * u0:2 * u1].reshape(U, 2).astype(np.int64)
        cat = [chr(c) for c in a["rep_cat"][u0:u1]]
        r0, r1 = int(b["tint_read_off"][t]), int(b["tint_read_off"][t + 1])
        rr = a["read_rep"][r0:r1]
        order = np.argsort(rr, kind="stable")
        bounds = np.searchsorted(rr[order], np.arange(U + 1))
        read_reps = [order[bounds[k]:bounds[k + 1]].tolist() for k in range(U)]
        if recycle_model in ("exons", "introns"):
            # garbage_cost_exons / _introns call .values() on a list (:187-196, :283-287): the reference raises
            raise AttributeError("'list' object has no attribute 'values'")
        cost = {i: int(a["rep_count"][u0 + i]) * 3 for i in range(U)} if recycle_model == "constant" else {}
        gaps = []
        for i in range(U):
            ri = r0 + int(a["rep_first_read"][u0 + i])
            g = {}
            if int(b["read_head"][8 * ri]) & 1:
                g0, g1 = int(b["read_gap_off"][ri]), int(b["read_gap_off"][ri + 1])
                for k in range(g0, g1):
                    g[(int(b["gap_rec"][3 * k]), int(b["gap_rec"][3 * k + 1]))] = int(b["gap_rec"][3 * k + 2])
            if cat[i] != "N":
                g[(int(a["rep_gap"][3 * (u0 + i)]), int(a["rep_gap"][3 * (u0 + i) + 1]))] = int(a["rep_gap"][3 * (u0 + i) + 2])
            gaps.append(g)
        parts = []
        for p in range(int(a["tint_part_off"][t]), int(a["tint_part_off"][t + 1])):
            rids = a["part_rids"][int(a["part_rid_off"][p]):int(a["part_rid_off"][p + 1])].tolist()
            i0, i1 = int(a["part_inc_off"][p]), int(a["part_inc_off"][p + 1])
            inc = a["inc"][2 * i0:2 * i1].reshape(-1, 2)
            parts.append((rids, [tuple(x) for x in inc.tolist()]))
        return dict(I=I, C=Cm, FL=FL, cat=cat, garbage_cost=cost, gaps=gaps, read_reps=read_reps, partitions=parts,
                    edges_before=int(a["tint_edges"][2 * t]), edges_after=int(a["tint_edges"][2 * t + 1]),
                    n_structs=int(a["tint_struct_off"][t + 1] - a["tint_struct_off"][t]))


class ClusterPrep:
    """One cluster-prep context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        rc = self.lib.frs_cprep_create(device, C.byref(self.ctx))
        if rc != 0:
            raise _lib.FrsError(rc, "frs_cprep_create: no usable CUDA device %d; there is no CPU fallback" % device)

    def close(self):
        if self.ctx:
            self.lib.frs_cprep_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = (self.lib.frs_cprep_last_error(self.ctx) or b"").decode()
            if msg.startswith("AssertionError"):
                raise AssertionError("[libfreddie_b200 %d] %s" % (rc, msg))
            if msg.startswith("ZeroDivisionError"):
                raise ZeroDivisionError(msg)
            raise _lib.FrsError(rc, msg)

    def run(self, batch: Dict[str, np.ndarray], maximum_ilp_size: int = 1000) -> ClusterPrepResult:
        """``batch``: the arrays of ``frs_cluster_batch`` (see ``batch_from_segment`` / ``batch_from_tints``)."""
        dts = dict(tint_read_off="i4", tint_seg_n="i4", tint_digit_off="i8", read_row="i4", digits="u1", read_head="i4",
                   read_gap_off="i4", gap_rec="i4")
        keep = {n: np.ascontiguousarray(batch[n], dtype=dts[n]) for n in _lib.CLUSTER_BATCH_ARRAYS}
        T, N = len(keep["tint_seg_n"]), len(keep["read_row"])
        sb = _lib.FrsClusterBatch(n_tints=T, n_reads=N, **{n: keep[n].ctypes.data for n in keep})
        sizes = _lib.FrsClusterSizes()
        self._check(self.lib.frs_cprep_run(self.ctx, C.byref(sb), int(maximum_ilp_size), C.byref(sizes)))
        U, P = int(sizes.n_reps), int(sizes.n_parts)
        shape = dict(tint_rep_off=T + 1, read_rep=N, rep_first_read=U, rep_count=U, rep_fl=2 * U, rep_cat=U, rep_gap=3 * U,
                     tint_row_off=T + 1, I=int(sizes.n_row_bytes), C=int(sizes.n_row_bytes), tint_struct_off=T + 1,
                     rep_struct=U, tint_part_off=T + 1, part_rid_off=P + 1, part_rids=U, part_inc_off=P + 1,
                     inc=2 * int(sizes.n_incomp), tint_edges=2 * T)
        out = {n: np.zeros(shape[n], dtype=dt) for n, dt in _lib.CLUSTER_RESULT_ARRAYS}
        rs = _lib.FrsClusterResult(**{n: out[n].ctypes.data for n in out})
        self._check(self.lib.frs_cprep_fetch(self.ctx, C.byref(rs)))
        ms = (C.c_float * 5)()
        self.lib.frs_cprep_timings(self.ctx, ms, 5)
        return ClusterPrepResult(out, sizes, keep, list(ms))


def batch_from_segment(batch_arrays: Dict[str, np.ndarray], result_arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """``frs_cluster_batch`` arrays from a segment batch (``PackedBatch.arrays``: ``tint_read_off``, ``tint_rep_off``,
    ``read_rep``) and its results (``BatchResult.arrays`` or the sections of a FRSSEGM1 file)."""
    tro = np.asarray(batch_arrays["tint_read_off"], dtype=np.int64)
    rep0 = np.asarray(batch_arrays["tint_rep_off"], dtype=np.int64)
    read_tint = np.repeat(np.arange(len(tro) - 1), np.diff(tro))
    return dict(
        tint_read_off=tro, tint_seg_n=np.diff(np.asarray(result_arrays["tint_final_off"], dtype=np.int64)) - 1,
        tint_digit_off=result_arrays["tint_digit_off"], read_row=np.asarray(batch_arrays["read_rep"], dtype=np.int64) - rep0[read_tint],
        digits=result_arrays["digits"], read_head=result_arrays["read_head"], read_gap_off=result_arrays["read_gap_off"],
        gap_rec=result_arrays["gap_rec"])


def batch_from_tints(tints: Sequence[dict]) -> Dict[str, np.ndarray]:
    """``frs_cluster_batch`` arrays from tints in the shape ``freddie_cluster.read_segment`` returns (:119-172):
    every read gets its own digit row; ``read_head`` is rebuilt from ``softclip`` / ``poly_tail``."""
    tro, segn, doff, rows, digits, head, goff, grec = [0], [], [0], [], [], [], [0], []
    for tint in tints:
        M = len(tint["segs"])
        segn.append(M)
        for k, read in enumerate(tint["reads"]):
            assert len(read["data"]) == M, (read["data"], tint["segs"])
            rows.append(k)
            digits.append(np.asarray(read["data"], dtype=np.uint8) + 48)
            h = [0] * 8
            if read["gaps"] or read["softclip"] or read["poly_tail"]:
                h[0] = 1
                h[3], h[6] = int(read["softclip"].get("SSC", 0)), int(read["softclip"].get("ESC", 0))
                for key, (a, b) in read["poly_tail"].items():
                    kind = 1 + "AT".index(key[1])
                    if key[0] == "S":
                        h[0] |= kind << 8
                        h[1], h[2] = int(a), int(b)
                    else:
                        h[0] |= kind << 16
                        h[4], h[5] = int(a), int(b)
            head.extend(h)
            for (a, b), v in read["gaps"].items():
                if a >= 0 and b < M:  # not the tail gaps a previous preprocess_ilp added
                    grec.extend((int(a), int(b), int(v)))
            goff.append(len(grec) // 3)
        tro.append(tro[-1] + len(tint["reads"]))
        doff.append(doff[-1] + len(tint["reads"]) * M)
    return dict(tint_read_off=np.asarray(tro), tint_seg_n=np.asarray(segn), tint_digit_off=np.asarray(doff),
                read_row=np.asarray(rows, dtype=np.int32), digits=np.concatenate(digits) if digits else np.zeros(0, np.uint8),
                read_head=np.asarray(head, dtype=np.int32), read_gap_off=np.asarray(goff), gap_rec=np.asarray(grec, dtype=np.int32))


_ctx: Optional[ClusterPrep] = None


def _context() -> ClusterPrep:
    global _ctx
    if _ctx is None:
        _ctx = ClusterPrep(0)
    return _ctx


def _apply_preprocess(tint: dict, res: dict, ilp_settings: dict) -> None:
    if "read_reps" in tint and [list(x) for x in tint["read_reps"]] != res["read_reps"]:
        raise AssertionError("read reps of the tint differ from read_segment's merge (freddie_cluster.py:154-164)")
    tint["read_reps"] = res["read_reps"]
    U = len(res["cat"])
    for i, idxs in enumerate(res["read_reps"]):
        first = tint["reads"][idxs[0]]
        first["gaps"] = dict(res["gaps"][i])
        for ridx in idxs:  # (:317-319) every read of the rep shares the first read's category and gaps
            tint["reads"][ridx]["poly_tail_category"] = res["cat"][i]
            tint["reads"][ridx]["gaps"] = first["gaps"]
    tint["ilp_data"] = dict(
        FL={i: (int(res["FL"][i][0]), int(res["FL"][i][1])) for i in range(U)},
        I={i: res["I"][i].tolist() for i in range(U)},
        C={i: res["C"][i].tolist() for i in range(U)},
        garbage_cost=res["garbage_cost"],
    )


def preprocess_ilp(tint: dict, ilp_settings: dict) -> None:
    """Drop-in for ``freddie_cluster.preprocess_ilp`` (:277-328)."""
    res = _context().run(batch_from_tints([tint]), 1000).tint(0, ilp_settings["recycle_model"])
    _apply_preprocess(tint, res, ilp_settings)


def partition_reads(tint: dict, maximum_ilp_size: int) -> None:
    """Drop-in for ``freddie_cluster.partition_reads`` (:198-274); needs ``preprocess_ilp`` first, like the reference."""
    if "ilp_data" not in tint:
        raise KeyError("ilp_data")
    res = _context().run(batch_from_tints([tint]), maximum_ilp_size).tint(0, "relative")
    tint["partitions"] = res["partitions"]


def cluster_prep_tints(tints: List[dict], ilp_settings: dict, maximum_ilp_size: int, device: int = 0) -> ClusterPrepResult:
    """Both steps for many tints in ONE device run (the form a pipeline should use)."""
    ctx = _context() if device == 0 else ClusterPrep(device)
    out = ctx.run(batch_from_tints(tints), maximum_ilp_size)
    for t, tint in enumerate(tints):
        res = out.tint(t, ilp_settings["recycle_model"])
        _apply_preprocess(tint, res, ilp_settings)
        tint["partitions"] = res["partitions"]
    return out
