"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the golden
vectors of the unmodified reference.  Integer / byte / index outputs are compared bit-exactly; the
smoothed fp64 signal is held to bit-identity as well (north_star allows 1e-6 relative)."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, flags_to_kwargs, sha_dir
from oracle import segment_oracle as orc

pytestmark = pytest.mark.gpu

ALL_SETS = ["degenerate", "plateau", "cfg1", "cfg2_small", "cfg2_flagsA", "cfg2_flagsB", "cfg2_sigma50",
            "cfg2_mps11", "dup_heavy", "cfg3_mini", "cfg4_mini", "cfg5_mini", "refine_tie", "cfg2_mps9", "empty_tint"]


@pytest.fixture(scope="module")
def eng():
    from freddie_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _params(flags):
    from freddie_b200.engine import SegmentParams
    o = orc.Params(**flags_to_kwargs(flags))
    return o, SegmentParams(o.sigma, o.tp, o.vf, o.mps, o.lo, o.ignore_ends)


@pytest.mark.parametrize("name", ALL_SETS)
def test_segment_text_equals_oracle(name, golden_set, eng):
    from freddie_b200.engine import format_tint
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    oprm, gprm = _params(flags)
    batch = pack_tints(tints)
    res = eng.segment_batch(batch, gprm)
    assert eng.launch_count() > 0
    otints = copy.deepcopy(tints)
    for t, ot in enumerate(otints):
        orc.segment_tint(ot, oprm)
        assert format_tint(batch, res, t) == orc.format_segment(ot), "tint %d of %s" % (t, name)


@pytest.mark.parametrize("name", ["degenerate", "plateau", "cfg2_small", "cfg2_flagsA", "cfg2_flagsB", "cfg4_mini",
                                  "dup_heavy", "cfg2_sigma50", "cfg2_mps11", "refine_tie", "cfg2_mps9", "empty_tint"])
def test_cli_directory_equals_reference_manifest(name, golden_set, manifest, tmp_path):
    """The drop-in CLI (native parser + kernels + native formatter) against the SHA-256 manifest of the
    SEGMENT directory the unmodified reference wrote for the same SPLIT directory."""
    _, flags, split_dir = golden_set(name)
    out = str(tmp_path / "seg")
    r = subprocess.run([sys.executable, "-m", "freddie_b200.segment", "-s", split_dir, "-o", out, "-t", "4"] + flags,
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[freddie_segment] Done with" in r.stdout
    assert sha_dir(out) == manifest[name]["outputs"]


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("name", ["cfg2_small", "cfg4_mini", "dup_heavy"])
def test_cli_equals_reference_manifest_in_every_launch_mode(name, mode, golden_set, manifest, tmp_path):
    """Programmatic dependent launch never (FRS_PDL=0) and always (FRS_PDL=2, also with several batches in
    flight) must give the bytes of the default policy, i.e. the reference's: the launch attribute only moves
    WHEN a kernel becomes resident, never what it reads."""
    _, flags, split_dir = golden_set(name)
    out = str(tmp_path / "seg")
    env = dict(os.environ, FRS_PDL=mode)
    r = subprocess.run([sys.executable, "-m", "freddie_b200.segment", "-s", split_dir, "-o", out, "-t", "4"] + flags,
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sha_dir(out) == manifest[name]["outputs"]


def test_cfg1_taps_equal_reference_intermediates(golden_set, eng):
    """Per-step taps against the intermediates dumped from the reference's own functions."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set("cfg1")
    _, gprm = _params(flags)
    batch = pack_tints(tints)
    res = eng.segment_batch(batch, gprm)
    z = np.load(os.path.join(GOLDEN, "cfg1_intermediates.npz"))
    assert np.array_equal(eng.tap(_lib.TAP_Y_RAW, np.int32), z["Y_raw"].astype(np.int32))
    y = eng.tap(_lib.TAP_Y, np.float64)
    assert np.allclose(y, z["Y"], rtol=1e-6, atol=0)  # the stated tolerance ...
    assert np.array_equal(y, z["Y"])                  # ... and in fact bit-identical
    assert eng.tap(_lib.TAP_THR, np.float64)[0] == float(z["thr"])
    cand = eng.tap(_lib.TAP_CAND, np.int32)
    off = z["island_off"]
    want = np.concatenate([z["cand"][z["cand_off"][a]:z["cand_off"][a + 1]] + off[a] for a in range(len(off) - 1)])
    assert np.array_equal(cand, want)
    fixed = np.flatnonzero(eng.tap(_lib.TAP_FIXED, np.uint8))
    wantf = np.concatenate([z["fixed"][z["fixed_off"][a]:z["fixed_off"][a + 1]] + z["cand_off"][a]
                            for a in range(len(off) - 1)])
    assert np.array_equal(fixed, wantf)
    assert res.arrays["final_pos"].tolist() == z["final_positions"].tolist()
    # coverage (get_cumulative_coverage, :188-246): P holds one row per candidate of the tint in flat
    # coordinates, so inside an island  P[c0 + c] - P[c0]  must be the reference's C[c]; the checksum of the
    # reference's own (K+1) x R matrix is reproduced from those rows plus the island's total row
    P = eng.tap(_lib.TAP_COVERAGE, np.uint32)
    R = batch.n_reps
    Rp = (R + 3) & ~3
    P = P.reshape(-1, Rp)
    assert P.shape[0] == len(cand)
    assert not P[:, R:].any()  # padding columns
    ot = copy.deepcopy(tints[0])
    keys, _ = orc.build_reps(ot)
    for a, isl in enumerate(ot["intervals"]):
        c0, c1 = int(z["cand_off"][a]), int(z["cand_off"][a + 1])
        rep_iv = [[(ts - isl[0], te - isl[0]) for ts, te in k if isl[0] <= ts <= isl[1]] for k in keys]
        C = orc.coverage_matrix(rep_iv, z["cand"][c0:c1].tolist())
        rows = (P[c0:c1, :R].astype(np.int64) - P[c0, :R].astype(np.int64))
        assert np.array_equal(rows, C[:-1].astype(np.int64)), "coverage rows of island %d" % a
        assert int(rows.sum() + C[-1].astype(np.uint64).sum()) == int(z["csum"][a]), "C.sum() of island %d" % a


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "cfg3_mini", "refine_tie", "dup_heavy"])
def test_refine_tap_equals_oracle(name, golden_set, eng):
    """refine_segmentation (:249-266): the positions refine ADDS to the DP-final breakpoints of every
    island (tap: final flags per sample minus the DP-final candidates) against the oracle's list."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    oprm, gprm = _params(flags)
    n_extra = 0
    for tint in tints[:6]:
        batch = pack_tints([tint])
        eng.segment_batch(batch, gprm)
        flags_final = eng.tap(_lib.TAP_FINAL_FLAGS, np.uint8)
        cand = eng.tap(_lib.TAP_CAND, np.int32)
        dpf = eng.tap(_lib.TAP_DP_FINAL, np.uint8)
        added = sorted(set(np.flatnonzero(flags_final).tolist()) - set(cand[dpf != 0].tolist()))
        it = orc.segment_tint(copy.deepcopy(tint), oprm, keep=True)
        off = np.cumsum([0] + [e - s + 1 for s, e in tint["intervals"]])
        want = sorted(int(off[a]) + p for a, ex in enumerate(it["refine"]) for p in ex)
        assert added == want
        dp_want = sorted(int(off[a]) + it["cand"][a][c] for a, ch in enumerate(it["dp_final"]) for c in ch)
        assert sorted(cand[dpf != 0].tolist()) == dp_want
        n_extra += len(want)
    if name in ("cfg1", "refine_tie"):
        assert n_extra > 0  # the set does exercise refine


def _check_dp_tables(eng, tint, oprm, limit=12):
    """ins / out blocks of the last run (one tint) against the oracle's numpy tables."""
    from freddie_b200 import _lib
    ot = copy.deepcopy(tint)
    it = orc.segment_tint(ot, oprm, keep=True)
    ss, sn = eng.tap(_lib.TAP_SUB_START, np.int32), eng.tap(_lib.TAP_SUB_N, np.int32)
    off, tab = eng.tap(_lib.TAP_SUB_TAB_OFF, np.int64), eng.tap(_lib.TAP_DP_TABLES, np.int32)
    cand_off = np.cumsum([0] + [len(c) for c in it["cand"]])
    keys, members = orc.build_reps(ot)
    checked = 0
    for p in range(len(ss)):
        a = int(np.searchsorted(cand_off, ss[p], side="right") - 1)
        start = int(ss[p] - cand_off[a])
        n = int(sn[p])
        size = n * (n - 1) // 2 + n * (n - 1) * (n - 2) // 6
        isl = ot["intervals"][a]
        rep_iv = [[(ts - isl[0], te - isl[0]) for ts, te in k if isl[0] <= ts <= isl[1]] for k in keys]
        C = orc.coverage_matrix(rep_iv, it["cand"][a])
        oi, oo = orc.dp_tables(it["cand"][a], C, it["W"], start, start + n - 1, oprm.table, oprm.tp)
        blk = tab[off[p]:off[p] + size]
        k = 0
        for i in range(n - 1):
            for j in range(i + 1, n):
                assert -blk[k] == oi[i, j], (p, i, j)
                k += 1
        for j in range(1, n - 1):          # out layout: [j][k-j-1][i]
            for kk in range(j + 1, n):
                for i in range(j):
                    assert blk[k] == oo[i, j, kk], (p, i, j, kk)
                    k += 1
        checked += 1
        if checked >= limit:
            break
    assert checked > 0


def test_dp_tables_equal_oracle(golden_set, eng):
    """ins / out tables of every subproblem of a weighted tint against the oracle's numpy tables: once
    from the on-chip tables of the one-CTA mode, once summed over rep slabs by the multi-CTA mode."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set("dup_heavy")
    oprm, gprm = _params(flags)
    batch = pack_tints(tints[:1])
    try:
        eng.set_option(_lib.OPT_KEEP_DP_TABLES, 1)
        eng.segment_batch(batch, gprm)
        _check_dp_tables(eng, tints[0], oprm)
        eng.set_option(_lib.OPT_KEEP_DP_TABLES, 0)
        eng.set_option(_lib.OPT_SLAB_WORDS, 2)
        eng.segment_batch(batch, gprm)
        _check_dp_tables(eng, tints[0], oprm)
    finally:
        eng.set_option(_lib.OPT_KEEP_DP_TABLES, 0)
        eng.set_option(_lib.OPT_SLAB_WORDS, 64)


@pytest.mark.parametrize("name", ["dup_heavy", "cfg3_mini", "cfg4_mini", "cfg2_flagsB"])
def test_multi_cta_mode_equals_one_cta_mode(name, golden_set, eng):
    """Giant-tint path on small data: with 32-rep slabs every tint above 32 reps runs its subproblems as
    several CTAs + k_dp_solve; all results must equal the default (one CTA per subproblem) run."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    _, gprm = _params(flags)
    batch = pack_tints(tints)
    want = eng.segment_batch(batch, gprm)
    try:
        for words in (1, 3):
            eng.set_option(_lib.OPT_SLAB_WORDS, words)
            got = eng.segment_batch(batch, gprm)
            for k in want.arrays:
                assert np.array_equal(want.arrays[k], got.arrays[k]), (name, words, k)
    finally:
        eng.set_option(_lib.OPT_SLAB_WORDS, 64)


@pytest.mark.parametrize("name", ["cfg1", "degenerate", "cfg2_small", "cfg5_mini"])
def test_warp_poly_scan_equals_thread_scan(name, golden_set, eng):
    """find_longest_poly as a chunked warp scan (used for clips >= 1024 bases) must give the results of
    the serial per-thread scan on every clip: force every scan task through it."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    _, gprm = _params(flags)
    batch = pack_tints(tints)
    try:
        eng.set_option(_lib.OPT_POLY_LONG_CLASS, 95)   # every scan by a single thread
        want = eng.segment_batch(batch, gprm)
        eng.set_option(_lib.OPT_POLY_LONG_CLASS, 1)    # every scan by a warp
        got = eng.segment_batch(batch, gprm)
    finally:
        eng.set_option(_lib.OPT_POLY_LONG_CLASS, 40)
    assert (want.arrays["read_head"].reshape(-1, 8)[:, 0] >> 8).any()  # the set does contain poly-A/T hits
    for k in want.arrays:
        assert np.array_equal(want.arrays[k], got.arrays[k]), (name, k)


@pytest.mark.parametrize("name", ["cfg1", "degenerate", "cfg2_small", "cfg4_mini"])
def test_lazy_clip_fetch_equals_resident_planes(name, golden_set, eng):
    """Default mode with PINNED sequence planes: nothing of them is uploaded; after segmentation a kernel
    fetches only the soft-clip words straight from the caller's host memory.  With the planes fully
    resident (pageable arrays, or FRS_OPT_LAZY_SEQ = 0) the results must be identical."""
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    _, gprm = _params(flags)
    pinned = pack_tints(tints).pin(edge_words=0)  # no edge store: every clip word crosses the bus on demand
    lazy = eng.segment_batch(pinned, gprm)
    st = eng.stats()
    assert 0 < st["clip_words"] <= st["seq_words"] and st["h2d_run"] == 8 * st["clip_words"]
    h_lazy, w_lazy = st["h2d_upload"], st["clip_words"]
    try:
        eng.set_option(_lib.OPT_LAZY_SEQ, 0)
        full = eng.segment_batch(pinned, gprm)
        st = eng.stats()
        assert st["h2d_run"] == 0 and st["h2d_upload"] == h_lazy + 8 * st["seq_words"]
    finally:
        eng.set_option(_lib.OPT_LAZY_SEQ, 1)
    pageable = eng.segment_batch(pack_tints(tints), gprm)  # pageable planes are copied whole
    assert eng.stats()["h2d_run"] == 0
    for k in lazy.arrays:
        assert np.array_equal(lazy.arrays[k], full.arrays[k]), (name, k)
        assert np.array_equal(lazy.arrays[k], pageable.arrays[k]), (name, k)
    # with an edge store (first / last E plane words of every read, copied densely at upload) only the clips
    # longer than 32 E bases are fetched afterwards; E = 1 and 2 leave many of those, the default E few
    for E in (1, 2, None):
        b = pack_tints(tints).pin(edge_words=E)
        got = eng.segment_batch(b, gprm)
        st = eng.stats()
        assert st["clip_words"] <= w_lazy and st["h2d_run"] == 8 * st["clip_words"]
        assert E is not None or name == "degenerate" or st["clip_words"] < w_lazy
        assert st["h2d_upload"] == h_lazy + 16 * b.n_reads * b.seq_edge_words
        for k in lazy.arrays:
            assert np.array_equal(lazy.arrays[k], got.arrays[k]), (name, E, k)


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "dup_heavy", "cfg3_mini", "degenerate"])
def test_compact_interval_encodings_equal_the_full_arrays(name, golden_set, eng):
    """frs_batch.cigar16 / riv_cig_n / qe_from_cigar: CIGAR ops as uint16, op counts as uint8 and query ends derived
    from qs + CIGAR are expanded on the device into the arrays the kernels read; results must equal those with the
    full arrays, and the upload must shrink by exactly the bytes saved."""
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    _, gprm = _params(flags)
    full = pack_tints(tints)
    want = eng.segment_batch(full, gprm)
    h_full = eng.stats()["h2d_upload"]
    comp = pack_tints(tints).compact()
    assert set(comp.compact_arrays) == {"cigar16", "riv_cig_n"} and comp.qe_from_cigar
    got = eng.segment_batch(comp, gprm)
    c = full.counts()
    saved = 2 * c["n_cigar_ops"] + (4 * (c["n_read_ivs"] + 1) - c["n_read_ivs"]) + 4 * c["n_read_ivs"]
    assert h_full - eng.stats()["h2d_upload"] == saved
    for k in want.arrays:
        assert np.array_equal(want.arrays[k], got.arrays[k]), (name, k)
    n_copies = eng.stats()["h2d_copies"]
    assert n_copies >= 20  # separately allocated arrays: one copy each
    # the same batch in ONE pinned arena laid out in the library's upload order: a handful of large copies
    arena = pack_tints(tints).pin()
    got = eng.segment_batch(arena, gprm)
    assert eng.stats()["h2d_copies"] <= 3, eng.stats()
    for k in want.arrays:
        assert np.array_equal(want.arrays[k], got.arrays[k]), (name, k)
    # an interval whose query end is NOT qs + CIGAR keeps its array
    odd = pack_tints(tints)
    odd.arrays["riv_qe"][0] += 1
    assert not odd.compact().qe_from_cigar and odd.as_struct().riv_qe is not None


def test_pipelined_submit_wait_fetch_equals_the_synchronous_calls(golden_set):
    """frs_submit / frs_wait / frs_fetch: up to six batches in flight in ONE context (the copies of one overlap the
    kernels of the other); results must be those of upload + run + download, in any interleaving."""
    from freddie_b200 import _lib
    from freddie_b200.engine import Engine
    from freddie_b200.pack import pack_tints
    names = ["cfg2_small", "cfg4_mini", "cfg5_mini", "dup_heavy", "cfg1"]
    batches, prms, want = [], [], []
    ref = Engine(0)
    for nme in names:
        tints, flags, _ = golden_set(nme)
        _, gprm = _params(flags)
        b = pack_tints(tints).pin()
        batches.append(b)
        prms.append(gprm)
        want.append(ref.segment_batch(b, gprm))
    ref.close()
    e = Engine(0)  # fresh context: its first runs also exercise the grow-and-repeat path of the capacities
    try:
        got = [None] * len(names)
        t_prev = None
        for k in range(len(names) + 1):
            t_new = e.submit(batches[k], prms[k]) if k < len(names) else None
            if t_prev is not None:
                sizes = e.wait(t_prev)
                got[k - 1] = e.fetch(t_prev, e.new_result(sizes, batches[k - 1], pinned=True))
            t_prev = t_new
        for k, nme in enumerate(names):
            for a in want[k].arrays:
                assert np.array_equal(want[k].arrays[a], got[k].arrays[a]), (nme, a)
            assert want[k].sizes == got[k].sizes
        # six batches may be in flight; a seventh submit without a fetch is refused; so is a fetch of a free ticket
        t0 = e.submit(batches[0], prms[0])
        t1 = e.submit(batches[1], prms[1])
        t2 = e.submit(batches[2], prms[2])
        t3 = e.submit(batches[3], prms[3])
        t4 = e.submit(batches[4], prms[4])
        t5 = e.submit(batches[0], prms[0])
        with pytest.raises(_lib.FrsError, match="in flight"):
            e.submit(batches[1], prms[1])
        for tk, kb in ((t4, 4), (t5, 0)):
            rk = e.fetch(tk, e.new_result(e.wait(tk), batches[kb]))
            assert all(np.array_equal(rk.arrays[a], want[kb].arrays[a]) for a in rk.arrays)
        s0, s1, s2 = e.wait(t0), e.wait(t1), e.wait(t2)
        r3 = e.fetch(t3, e.new_result(e.wait(t3), batches[3]))
        assert all(np.array_equal(r3.arrays[a], want[3].arrays[a]) for a in r3.arrays)
        r1 = e.fetch(t1, e.new_result(s1, batches[1]))
        r2 = e.fetch(t2, e.new_result(s2, batches[2]))
        r0 = e.fetch(t0, e.new_result(s0, batches[0]))
        assert all(np.array_equal(r0.arrays[a], want[0].arrays[a]) for a in r0.arrays)
        assert all(np.array_equal(r1.arrays[a], want[1].arrays[a]) for a in r1.arrays)
        assert all(np.array_equal(r2.arrays[a], want[2].arrays[a]) for a in r2.arrays)
        with pytest.raises(_lib.FrsError, match="no batch in flight"):
            e.fetch(t0, r0)
        # fetch in two halves: the copy of one batch is collected behind the submit of another
        ta = e.submit(batches[4], prms[4])
        sa = e.wait(ta)
        ra = e.fetch_start(ta, e.new_result(sa, batches[4], pinned=True))
        tb = e.submit(batches[0], prms[0])
        e.fetch_finish(ta)
        assert all(np.array_equal(ra.arrays[a], want[4].arrays[a]) for a in ra.arrays)
        rb = e.fetch(tb, e.new_result(e.wait(tb), batches[0]))
        assert all(np.array_equal(rb.arrays[a], want[0].arrays[a]) for a in rb.arrays)
        # the synchronous calls still work on the same context afterwards
        again = e.segment_batch(batches[3], prms[3])
        assert all(np.array_equal(again.arrays[a], want[3].arrays[a]) for a in again.arrays)
    finally:
        e.close()


def test_capacity_miss_repeats_the_run(golden_set):
    """Data-dependent buffers (coverage, DP tables, digits, runs, gaps, clip words) are sized from capacities
    that only grow; a fresh context meets a batch whose needs exceed its first guesses, repeats the run with
    larger buffers (no host round trip inside a run) and returns the same results."""
    from freddie_b200 import _lib
    from freddie_b200.engine import Engine
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set("cfg3_mini")
    _, gprm = _params(flags)
    small, _, _ = golden_set("degenerate")
    e = Engine(0)
    try:
        e.set_option(_lib.OPT_SLAB_WORDS, 2)   # multi-CTA DP: global tables, many work items
        e.segment_batch(pack_tints(small), gprm)
        first = e.segment_batch(pack_tints(tints).pin(), gprm)
        assert e.stats()["reruns"] >= 1
        n = e.stats()["reruns"]
        second = e.segment_batch(pack_tints(tints).pin(), gprm)
        assert e.stats()["reruns"] == n           # capacities are warm now
        for k in first.arrays:
            assert np.array_equal(first.arrays[k], second.arrays[k]), k
    finally:
        e.close()
    ref = Engine(0)
    want = ref.segment_batch(pack_tints(tints), gprm)
    ref.close()
    for k in want.arrays:
        assert np.array_equal(first.arrays[k], want.arrays[k]), k


@pytest.mark.parametrize("name", ["cfg1", "degenerate", "dup_heavy", "cfg3_mini", "cfg5_mini"])
def test_derived_read_intervals_equal_uploaded_ones(name, golden_set, eng):
    """frs_batch.riv_ts / riv_te NULL (default of the Python packer): the genomic target intervals of
    every read are derived on the device from its rep's flat-sample intervals (freddie_segment.py:165-170:
    the rep key IS the tuple of target intervals).  Results must equal those with the arrays uploaded,
    and the transfer must shrink by exactly the two arrays."""
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set(name)
    _, gprm = _params(flags)
    batch = pack_tints(tints)
    assert batch.derive_riv
    lean = eng.segment_batch(batch, gprm)
    h_lean = eng.stats()["h2d_upload"]
    batch.derive_riv = False
    full = eng.segment_batch(batch, gprm)
    h_full = eng.stats()["h2d_upload"]
    assert h_full - h_lean == 8 * batch.counts()["n_read_ivs"]
    for k in lean.arrays:
        assert np.array_equal(lean.arrays[k], full.arrays[k]), (name, k)


def test_in_process_seam_has_reference_signature(golden_set):
    """segment(tint, sigma, smoothed_threshold, tp, vf, mps, lo, ignore_ends) mutates the tint like the
    reference (freddie_segment.py:738-844)."""
    from freddie_b200.segment import segment, smooth_threshold
    tints, _, _ = golden_set("plateau")
    t = copy.deepcopy(tints[0])
    ot = copy.deepcopy(tints[0])
    assert segment(t, 5.0, smooth_threshold(0.9), 0.9, 3.0, 50, 3, True) == t["id"]
    orc.segment_tint(ot, orc.Params())
    assert t["final_positions"] == ot["final_positions"] and t["segs"] == ot["segs"]
    for r, o in zip(t["reads"], ot["reads"]):
        assert r["data"] == o["data"] and r["gaps"] == o["gaps"]


def test_output_independent_of_batch_composition(golden_set, eng):
    from freddie_b200.engine import format_tint
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set("cfg4_mini")
    _, gprm = _params(flags)
    whole = pack_tints(tints)
    res = eng.segment_batch(whole, gprm)
    texts = [format_tint(whole, res, t) for t in range(len(tints))]
    again = eng.segment_batch(whole, gprm)
    assert all(np.array_equal(res.arrays[k], again.arrays[k]) for k in res.arrays)  # deterministic
    order = list(range(len(tints)))[::-1]
    for lo in range(0, len(order), 50):
        part = [tints[i] for i in order[lo:lo + 50]]
        b = pack_tints(part)
        r = eng.segment_batch(b, gprm)
        for k, i in enumerate(order[lo:lo + 50]):
            assert format_tint(b, r, k) == texts[i]


def test_full_size_config2_properties(eng):
    """BASELINE config 2 at full size (200k reads, ~3k tints): size-independent properties plus an
    oracle spot check on a sample of tints."""
    from freddie_b200 import synth
    from freddie_b200.engine import SegmentParams, format_tint
    from freddie_b200.pack import pack_tints
    tints = synth.make_config(2, workers=min(8, os.cpu_count() or 1))
    assert sum(len(t["reads"]) for t in tints) > 150000
    batch = pack_tints(tints)
    res = eng.segment_batch(batch, SegmentParams())
    a, ba = res.arrays, batch.arrays
    assert set(np.unique(a["digits"]).tolist()) <= {48, 49, 50}
    for t, tint in enumerate(tints):
        f = a["final_pos"][a["tint_final_off"][t]:a["tint_final_off"][t + 1]]
        assert np.all(np.diff(f) > 0)
        ends = sorted(x for iv in tint["intervals"] for x in iv)
        assert set(ends) <= set(f.tolist()) and f[0] == ends[0] and f[-1] == ends[-1]
        S = len(f) - 1
        R = ba["tint_rep_off"][t + 1] - ba["tint_rep_off"][t]
        assert a["tint_digit_off"][t + 1] - a["tint_digit_off"][t] == S * R
    rng = np.random.default_rng(0)
    for t in rng.choice(len(tints), size=25, replace=False):
        ot = copy.deepcopy(tints[int(t)])
        orc.segment_tint(ot, orc.Params())
        assert format_tint(batch, res, int(t)) == orc.format_segment(ot)


def test_errors_are_loud(golden_set, eng):
    from freddie_b200 import _lib
    from freddie_b200.engine import SegmentParams
    from freddie_b200.pack import pack_tints
    tints, _, _ = golden_set("plateau")
    batch = pack_tints(tints)
    with pytest.raises(_lib.FrsError, match="max_problem_size < 9"):
        eng.segment_batch(batch, SegmentParams(max_problem_size=4))
    bad = pack_tints(tints)
    bad.arrays["read_rep"][0] = 10 ** 6
    with pytest.raises(_lib.FrsError, match="rep"):
        eng.upload(bad)


@pytest.mark.parametrize("name", ["cfg2_small", "cfg4_mini"])
def test_cli_on_a_packed_directory_equals_reference_manifest(name, golden_set, manifest, tmp_path):
    """The packed side-channel: SPLIT directory -> packed batches -> the same CLI.  The
    SEGMENT directory must be the one the unmodified reference wrote from the TSV files."""
    _, flags, split_dir = golden_set(name)
    pk, out = str(tmp_path / "packed"), str(tmp_path / "seg")
    r = subprocess.run([sys.executable, "-m", "freddie_b200.packed", "-s", split_dir, "-o", pk, "-t", "4",
                        "--batch-reads", "700"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "freddie_b200.segment", "-s", pk, "-o", out, "-t", "4"] + flags,
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert sha_dir(out) == manifest[name]["outputs"]
