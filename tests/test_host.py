"""Host logic on CPU: packing, native parser / formatter, scheduler, CLI surface, sharding (gloo)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, flags_to_kwargs, sha_dir
from helpers import oracle_result_arrays
from oracle import segment_oracle as orc


def test_pack_invariants(golden_set):
    from freddie_b200.pack import pack_tints
    tints, _, _ = golden_set("cfg2_flagsA")
    b = pack_tints(tints)
    a = b.arrays
    c = b.counts()
    assert c["n_tints"] == len(tints) and c["n_reads"] == sum(len(t["reads"]) for t in tints)
    assert int(a["rep_weight"].sum()) == c["n_reads"]
    assert np.all(np.diff(a["island_sample_off"]) >= 2)
    assert np.all(a["rep_iv_fs"] < a["rep_iv_fe"])
    # reps are the distinct target-interval tuples in first-seen order (freddie_segment.py:165-170)
    for t, tint in enumerate(tints):
        keys, members = orc.build_reps(tint)
        r0, r1 = a["tint_rep_off"][t], a["tint_rep_off"][t + 1]
        assert r1 - r0 == len(keys)
        assert a["rep_weight"][r0:r1].tolist() == [len(m) for m in members]
    # bit-planes
    r = tints[0]["reads"][0]
    w0 = int(a["read_seq_off"][0])
    bits = [(int(a["seq_is_a"][w0 + i // 32]) >> (i % 32)) & 1 for i in range(len(r["seq"]))]
    assert bits == [int(ch == "A") for ch in r["seq"]]


def test_pack_rejects_what_the_reference_rejects():
    from freddie_b200.pack import pack_tints
    from freddie_b200 import synth
    t = synth.make_degenerate()[0]
    bad = dict(t, intervals=[(1000, 1100), (1500, 1800)])  # first interval ends outside its island
    with pytest.raises(KeyError):
        pack_tints([bad])
    bad = dict(t, intervals=[(1000, 1800), (1700, 1900)])
    with pytest.raises(AssertionError):
        pack_tints([bad])


def _native_batch(split_dir, tints, threads=4):
    from freddie_b200 import hostio
    jobs = [(t["chr"], t["id"]) for t in tints]
    sp, rp, _, _ = hostio._paths(split_dir, "/nonexistent", jobs)
    return hostio.ParsedBatch(sp, rp, threads)


@pytest.mark.parametrize("name", ["cfg1", "cfg2_flagsA", "degenerate", "cfg4_mini"])  # cfg1: single-tint batch (arrays moved)
def test_native_parser_equals_python_packer(name, golden_set, built_lib):
    from freddie_b200 import _lib
    from freddie_b200.pack import pack_tints
    from freddie_b200.segment import _load_tint_py
    tints, _, split_dir = golden_set(name)
    pb = _native_batch(split_dir, tints)
    ref = pack_tints([_load_tint_py(split_dir, t["chr"], t["id"]) for t in tints])
    for k, v in ref.counts().items():
        assert getattr(pb.struct, k) == v, k
    for nm in _lib.BATCH_ARRAYS:
        arr = ref.arrays[nm]
        if arr.size == 0:
            continue
        ptr = C.cast(getattr(pb.struct, nm), C.POINTER(np.ctypeslib.as_ctypes_type(arr.dtype)))
        assert np.array_equal(np.ctypeslib.as_array(ptr, shape=arr.shape), arr), nm
    pb.close()


def test_native_parser_errors_mirror_reference(tmp_path, built_lib):
    from freddie_b200 import hostio, synth, _lib
    t = synth.make_degenerate()[2]
    d = str(tmp_path / "s")
    synth.write_split_dir([t], d)
    sp = os.path.join(d, t["chr"], "split_%s_%d.tsv" % (t["chr"], t["id"]))
    rp = os.path.join(d, t["chr"], "reads_%s_%d.tsv" % (t["chr"], t["id"]))
    good = open(sp).read()
    # wrong read count -> the reference's assert at :164
    open(sp, "w").write(good.replace("\t3\n", "\t4\n", 1))
    with pytest.raises(AssertionError, match="read_count"):
        hostio.ParsedBatch([sp.encode()], [rp.encode()], 1)
    # malformed CIGAR -> no regex match
    open(sp, "w").write(good.replace("M", "Q", 1))
    with pytest.raises(_lib.FrsError, match="read_prog"):
        hostio.ParsedBatch([sp.encode()], [rp.encode()], 1)
    # missing sequence row -> assert at :181
    open(sp, "w").write(good)
    open(rp, "w").write("".join(open(rp).readlines()[:-1]))
    with pytest.raises(AssertionError, match="rid_to_seq"):
        hostio.ParsedBatch([sp.encode()], [rp.encode()], 1)
    with pytest.raises(_lib.FrsError, match="FileNotFoundError"):
        hostio.ParsedBatch([b"/nonexistent/split_x_0.tsv"], [rp.encode()], 1)


@pytest.mark.parametrize("name", ["cfg1", "cfg2_flagsA", "degenerate", "dup_heavy", "cfg3_mini"])
def test_native_parser_chunked_path_equals_sequential(name, golden_set, built_lib, monkeypatch):
    """Giant tints are parsed by all threads inside one file (rows cut into chunks at line ends, appended
    and deduped in row order).  FRS_PARSE_BIG_BYTES=1 forces that path on the golden sets: every array
    must equal the sequential parse, and the first error in row order must be the same one."""
    from freddie_b200 import _lib, hostio
    tints, _, split_dir = golden_set(name)
    seq = _native_batch(split_dir, tints)
    monkeypatch.setenv("FRS_PARSE_BIG_BYTES", "1")
    sp = [("%s/%s/split_%s_%d.tsv" % (split_dir, t["chr"], t["chr"], t["id"])).encode() for t in tints]
    rp = [("%s/%s/reads_%s_%d.tsv" % (split_dir, t["chr"], t["chr"], t["id"])).encode() for t in tints]
    par = hostio.ParsedBatch(sp, rp, 4)
    for k in _lib.BATCH_COUNTS + ["n_seq_words"]:
        assert getattr(par.struct, k) == getattr(seq.struct, k), k
    sizes = dict(tint_island_off="n_tints+1", tint_rep_off="n_tints+1", tint_read_off="n_tints+1", island_start="n_islands",
                 island_sample_off="n_islands+1", rep_iv_off="n_reps+1", rep_weight="n_reps", rep_iv_fs="n_rep_ivs",
                 rep_iv_fe="n_rep_ivs", read_rep="n_reads", read_strand="n_reads", read_len="n_reads",
                 read_iv_off="n_reads+1", read_seq_off="n_reads+1", riv_ts="n_read_ivs", riv_te="n_read_ivs",
                 riv_qs="n_read_ivs", riv_qe="n_read_ivs", riv_cig_off="n_read_ivs+1", cigar="n_cigar_ops",
                 seq_is_a="n_seq_words", seq_is_t="n_seq_words")
    dt = dict(read_strand=C.c_uint8, read_seq_off=C.c_int64, cigar=C.c_uint32, seq_is_a=C.c_uint32, seq_is_t=C.c_uint32)
    for nm, expr in sizes.items():
        base, _, plus = expr.partition("+")
        n = int(getattr(seq.struct, base)) + (1 if plus else 0)
        if n == 0:
            continue
        ct = dt.get(nm, C.c_int32)
        a = np.ctypeslib.as_array(C.cast(getattr(seq.struct, nm), C.POINTER(ct)), shape=(n,))
        b = np.ctypeslib.as_array(C.cast(getattr(par.struct, nm), C.POINTER(ct)), shape=(n,))
        assert np.array_equal(a, b), nm
    seq.close()
    par.close()


def test_native_parser_chunked_path_reports_the_first_error_in_row_order(tmp_path, built_lib, monkeypatch):
    from freddie_b200 import hostio, synth, _lib
    t = synth.make_config(2, scale=0.002, seed=77)[0]
    d = str(tmp_path / "s")
    synth.write_split_dir([t], d)
    sp = os.path.join(d, t["chr"], "split_%s_%d.tsv" % (t["chr"], t["id"]))
    rp = os.path.join(d, t["chr"], "reads_%s_%d.tsv" % (t["chr"], t["id"]))
    lines = open(sp).read().split("\n")
    assert len(lines) > 12
    # an interval outside every island early (KeyError at dedupe, :666) and a malformed row later (read_prog)
    f = lines[3].split("\t")
    f[5] = "1-2:" + f[5].split(":", 1)[1]
    early = "\t".join(f)
    late = lines[-3].replace("M", "Q", 1)
    for big in ("1", None):
        if big:
            monkeypatch.setenv("FRS_PARSE_BIG_BYTES", big)
        else:
            monkeypatch.delenv("FRS_PARSE_BIG_BYTES", raising=False)
        open(sp, "w").write("\n".join(lines[:3] + [early] + lines[4:-3] + [late] + lines[-2:]))
        with pytest.raises(_lib.FrsError, match="KeyError: 1 "):
            hostio.ParsedBatch([sp.encode()], [rp.encode()], 4)
        open(sp, "w").write("\n".join(lines[:-3] + [late] + lines[-2:]))
        with pytest.raises(_lib.FrsError, match="read_prog"):
            hostio.ParsedBatch([sp.encode()], [rp.encode()], 4)
        open(sp, "w").write("\n".join(lines[:6] + [lines[0]] + lines[6:]))
        with pytest.raises(AssertionError, match="repeated"):
            hostio.ParsedBatch([sp.encode()], [rp.encode()], 4)


@pytest.mark.parametrize("chunked", [False, True])
@pytest.mark.parametrize("name", ["cfg2_flagsA", "degenerate", "plateau"])
def test_native_formatter_writes_reference_bytes(name, chunked, golden_set, manifest, tmp_path, built_lib, monkeypatch):
    """Formatter fed with the oracle's results (as frs_result arrays) must emit the reference's files;
    `chunked` forces the path of giant tints (rows formatted in chunks by all threads) on every tint."""
    from freddie_b200 import _lib
    if chunked:
        monkeypatch.setenv("FRS_FORMAT_BIG_ROWS", "0")
    tints, flags, split_dir = golden_set(name)
    pb = _native_batch(split_dir, tints)
    _, arrays = oracle_result_arrays(tints, orc.Params(**flags_to_kwargs(flags)))
    res = _lib.FrsResult()
    for k in _lib.RESULT_ARRAYS:
        setattr(res, k, arrays[k].ctypes.data_as(C.c_void_p))

    class R:
        def as_struct(self):
            return res
    out = str(tmp_path / "seg")
    ops, lps = [], []
    for t in tints:
        os.makedirs(os.path.join(out, t["chr"]), exist_ok=True)
        ops.append(os.path.join(out, t["chr"], "segment_%s_%d.tsv" % (t["chr"], t["id"])).encode())
        lps.append(os.path.join(out, t["chr"], "segment_%s_%d.log" % (t["chr"], t["id"])).encode())
    pb.format(R(), ops, lps, 3)
    assert sha_dir(out) == manifest[name]["outputs"]
    pb.close()


def test_python_formatter_and_gap_strings(golden_set, manifest):
    from freddie_b200.engine import BatchResult, format_tint, apply_result
    from freddie_b200.pack import pack_tints
    tints, flags, _ = golden_set("cfg2_flagsA")
    otints, arrays = oracle_result_arrays(tints, orc.Params(**flags_to_kwargs(flags)))
    batch = pack_tints(tints)
    res = BatchResult.__new__(BatchResult)
    res.arrays, res.sizes = arrays, {}
    for t, ot in enumerate(otints):
        assert format_tint(batch, res, t) == orc.format_segment(ot)
    apply_result(batch, res)
    for t, ot in zip(batch.tints, otints):
        assert t["final_positions"] == ot["final_positions"] and t["segs"] == ot["segs"]
        for r, o in zip(t["reads"], ot["reads"]):
            assert r["data"] == o["data"] and r["gaps"] == o["gaps"]


def test_params_tables_match_scipy_and_reference():
    from scipy.ndimage import _filters
    from freddie_b200.engine import SegmentParams, smooth_threshold, gaussian_kernel
    assert smooth_threshold(0.9) == orc.smooth_threshold(0.9) and smooth_threshold(1.0) == orc.smooth_threshold(1.0)
    for sd in (0.7, 5.0, 12.5, 50.0):
        for tr in (4.0, 1.0):
            lw = int(tr * sd + 0.5)
            assert np.array_equal(gaussian_kernel(sd, tr), _filters._gaussian_kernel1d(sd, 0, lw)[::-1])
    for bad in (dict(threshold_rate=0.4), dict(variance_factor=10), dict(sigma=0), dict(sigma=51),
                dict(max_problem_size=3), dict(min_read_support_outside=-1)):
        with pytest.raises(AssertionError):
            SegmentParams(**bad)


def test_cli_surface_matches_reference():
    from freddie_b200.segment import parse_args
    a = parse_args(["-s", "x"])
    assert (a.outdir, a.threads, a.sigma, a.threshold_rate, a.variance_factor, a.max_problem_size,
            a.min_read_support_outside, a.consider_ends) == ("freddie_segment/", 1, 5.0, 0.9, 3.0, 50, 3, False)
    a = parse_args(["-s", "x", "--consider-ends", "-sd", "2.5", "-tp", "0.8", "-vf", "1.5", "-mps", "12", "-lo", "1",
                    "-t", "4", "-o", "y"])
    assert a.consider_ends is True and a.max_problem_size == 12
    assert parse_args(["-s", "x", "--consider-ends", "no"]).consider_ends is False
    with pytest.raises(AssertionError):
        parse_args(["-s", "x", "-tp", "0.3"])


def test_read_split_equals_oracle_parser(golden_set):
    from freddie_b200.segment import read_split, read_sequence
    tints, _, d = golden_set("cfg2_flagsA")
    t = tints[3]
    sp = os.path.join(d, t["chr"], "split_%s_%d.tsv" % (t["chr"], t["id"]))
    rp = os.path.join(d, t["chr"], "reads_%s_%d.tsv" % (t["chr"], t["id"]))
    a = read_split(sp)[0]
    read_sequence(a, rp)
    b = orc.parse_split(sp)
    orc.parse_reads(b, rp)
    assert a["intervals"] == b["intervals"] and a["id"] == b["id"]
    for x, y in zip(a["reads"], b["reads"]):
        assert x == y


def test_lpt_partition_and_batches():
    from freddie_b200 import schedule
    costs = [(float(c), float(c)) for c in [100, 1, 1, 50, 49, 2, 3, 98]]
    bins = schedule.lpt_partition(costs, 2)
    assert sorted(i for b in bins for i in b) == list(range(8))
    loads = [sum(costs[i][0] for i in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 4
    assert schedule.lpt_partition(costs, 2) == bins  # deterministic
    groups = list(schedule.batches(list(range(8)), costs, 100))
    assert [i for g in groups for i in g] == list(range(8))
    assert all(sum(costs[i][1] for i in g) <= 100 or len(g) == 1 for g in groups)


_GLOO = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch.distributed as dist
from freddie_b200 import schedule, synth
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, ws = dist.get_rank(), dist.get_world_size()
plan = synth.config_plan(4, scale=0.002)
costs = [(schedule.estimate_cost(float(n)), float(n)) for n in plan["sizes"]]
mine = schedule.lpt_partition(costs, ws)[rank]
got = [None] * ws
dist.all_gather_object(got, dict(rank=rank, idx=mine, reads=int(sum(plan["sizes"][i] for i in mine)),
                                 cost=sum(costs[i][0] for i in mine)))
if rank == 0:
    print(json.dumps(dict(shards=got, n=len(costs), total=int(plan["sizes"].sum()))))
dist.destroy_process_group()
'''


def test_sharding_world_size_2_gloo(tmp_path):
    """The multi-GPU path is a partition with a host-side gather and no data-path collective: two ranks
    must cover every tint exactly once with balanced estimated cost."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    info = json.loads(outs[0][0].strip().splitlines()[-1])
    idx = sorted(i for s in info["shards"] for i in s["idx"])
    assert idx == list(range(info["n"]))
    assert sum(s["reads"] for s in info["shards"]) == info["total"]
    c = [s["cost"] for s in info["shards"]]
    assert abs(c[0] - c[1]) / max(c) < 0.25


def test_batch_feed_hands_every_batch_to_exactly_one_lane():
    """The CLI driver's lanes (host threads of one GPU) drain a shared feed: no batch twice, none lost,
    and stop() -- an error in any lane -- ends the feed for everybody (the reference aborts the whole
    run on an exception, freddie_segment.py:871-885)."""
    import threading
    from freddie_b200 import schedule
    jobs = [("chr1", i) for i in range(1000)]
    costs = [(1.0, 37.0 + (i % 11)) for i in range(1000)]
    want = list(schedule.batches(jobs, costs, 500))
    feed = schedule.BatchFeed(jobs, costs, 500)
    got, lock = [], threading.Lock()

    def lane():
        for chunk in feed:
            with lock:
                got.append(chunk)

    ths = [threading.Thread(target=lane) for _ in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert sorted(got) == sorted(want) and sum(len(c) for c in got) == len(jobs)
    feed = schedule.BatchFeed(jobs, costs, 500)
    assert feed.next() == want[0]
    feed.stop()
    assert feed.next() is None and list(feed) == []


def test_packed_batch_ships_read_intervals_once(golden_set):
    """frs_batch.riv_ts / riv_te are NULL by default (a read's target intervals are its rep's,
    freddie_segment.py:165-170, and the library derives them on the device); with derive_riv off the
    arrays are passed.  The rep intervals really do determine them (what k_derive_riv computes)."""
    from freddie_b200.pack import pack_tints
    tints, _, _ = golden_set("cfg2_small")
    b = pack_tints(tints)
    s = b.as_struct()
    assert b.derive_riv and not s.riv_ts and not s.riv_te and s.riv_qs and s.rep_iv_fs
    b.derive_riv = False
    s = b.as_struct()
    assert s.riv_ts and s.riv_te
    a = b.arrays
    iso, ist = a["island_sample_off"], a["island_start"]
    for r in range(0, b.n_reads, 7):
        k0, k1 = a["read_iv_off"][r], a["read_iv_off"][r + 1]
        q0 = a["rep_iv_off"][a["read_rep"][r]]
        fs, fe = a["rep_iv_fs"][q0:q0 + k1 - k0], a["rep_iv_fe"][q0:q0 + k1 - k0]
        isl = np.searchsorted(iso, fs, side="right") - 1
        assert np.array_equal(ist[isl] + fs - iso[isl], a["riv_ts"][k0:k1])
        assert np.array_equal(ist[isl] + fe - iso[isl], a["riv_te"][k0:k1])


# ---- packed side-channel --------------------------------------------------------
def _batch_arrays(pb):
    """Every array of a ParsedBatch's frs_batch as numpy copies, keyed by name."""
    from freddie_b200 import _lib
    st = pb.struct
    n = {k: int(getattr(st, k)) for k in _lib.BATCH_COUNTS + ["n_seq_words"]}
    size = dict(tint_island_off=n["n_tints"] + 1, tint_rep_off=n["n_tints"] + 1, tint_read_off=n["n_tints"] + 1,
                island_start=n["n_islands"], island_sample_off=n["n_islands"] + 1, rep_iv_off=n["n_reps"] + 1,
                rep_weight=n["n_reps"], rep_iv_fs=n["n_rep_ivs"], rep_iv_fe=n["n_rep_ivs"], read_rep=n["n_reads"],
                read_strand=n["n_reads"], read_len=n["n_reads"], read_iv_off=n["n_reads"] + 1, read_seq_off=n["n_reads"] + 1,
                riv_ts=n["n_read_ivs"], riv_te=n["n_read_ivs"], riv_qs=n["n_read_ivs"], riv_qe=n["n_read_ivs"],
                riv_cig_off=n["n_read_ivs"] + 1, cigar=n["n_cigar_ops"], seq_is_a=n["n_seq_words"], seq_is_t=n["n_seq_words"])
    ct = dict(read_strand=C.c_uint8, read_seq_off=C.c_int64, cigar=C.c_uint32, seq_is_a=C.c_uint32, seq_is_t=C.c_uint32)
    out = dict(n)
    for nm, k in size.items():
        out[nm] = (np.ctypeslib.as_array(C.cast(getattr(st, nm), C.POINTER(ct.get(nm, C.c_int32))), shape=(k,)).copy()
                   if k else np.zeros(0))
    return out


@pytest.mark.parametrize("name", ["cfg1", "cfg2_flagsA", "degenerate", "cfg4_mini"])
def test_packed_batch_round_trip(name, golden_set, manifest, tmp_path, built_lib):
    """frs_packed_write -> frs_packed_read gives back every array of the batch, and the formatter fed from
    the re-read batch (names, chr and tint columns travel in the file) emits the reference's bytes."""
    from freddie_b200 import _lib, hostio
    tints, flags, split_dir = golden_set(name)
    pb = _native_batch(split_dir, tints)
    path = str(tmp_path / "b.frsb")
    pb.write_packed(path)
    rd = hostio.ParsedBatch.from_packed(path)
    a, b = _batch_arrays(pb), _batch_arrays(rd)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    _, arrays = oracle_result_arrays(tints, orc.Params(**flags_to_kwargs(flags)))
    res = _lib.FrsResult()
    for k in _lib.RESULT_ARRAYS:
        setattr(res, k, arrays[k].ctypes.data_as(C.c_void_p))

    class R:
        def as_struct(self):
            return res
    out = str(tmp_path / "seg")
    ops, lps = [], []
    for t in tints:
        os.makedirs(os.path.join(out, t["chr"]), exist_ok=True)
        ops.append(os.path.join(out, t["chr"], "segment_%s_%d.tsv" % (t["chr"], t["id"])).encode())
        lps.append(os.path.join(out, t["chr"], "segment_%s_%d.log" % (t["chr"], t["id"])).encode())
    rd.format(R(), ops, lps, 2)
    assert sha_dir(out) == manifest[name]["outputs"]
    pb.close()
    rd.close()


def test_packed_reader_rejects_damaged_files(golden_set, tmp_path, built_lib):
    from freddie_b200 import _lib, hostio
    tints, _, split_dir = golden_set("degenerate")
    pb = _native_batch(split_dir, tints)
    path = str(tmp_path / "b.frsb")
    pb.write_packed(path)
    pb.close()
    raw = open(path, "rb").read()
    for damaged in (raw[:100], b"NOTPACKD" + raw[8:], raw[:-40], raw[:16] + b"\xff" * 8 + raw[24:]):
        bad = str(tmp_path / "bad.frsb")
        open(bad, "wb").write(damaged)
        with pytest.raises(_lib.FrsError, match="not a packed batch"):
            hostio.ParsedBatch.from_packed(bad)
    with pytest.raises(_lib.FrsError, match="FileNotFoundError"):
        hostio.ParsedBatch.from_packed(str(tmp_path / "missing.frsb"))


@pytest.mark.parametrize("mode", ["tsv", "packed"])
def test_directory_driver_with_an_oracle_backed_engine(mode, golden_set, manifest, tmp_path, built_lib, monkeypatch):
    """run_directory's plumbing without a GPU: batching, lanes, the packed index, the native parser and
    formatter are the real ones; only the context that would run the kernels is replaced by one that
    answers with the oracle's results for the batch it is handed.  Output must be the reference's."""
    from freddie_b200 import _lib, packed, schedule, segment
    from freddie_b200.engine import SegmentParams
    name = "cfg2_flagsA"
    tints, flags, split_dir = golden_set(name)
    by_key = {(t["chr"], t["id"]): t for t in tints}
    oprm = orc.Params(**flags_to_kwargs(flags))
    src = split_dir
    if mode == "packed":
        src = str(tmp_path / "packed")
        st = packed.pack_directory(split_dir, src, threads=2, batch_reads=400)
        assert st["tints"] == len(tints) and st["batches"] > 1 and packed.is_packed_dir(src)
        assert not packed.is_packed_dir(split_dir)

    class Result:
        def __init__(self, arrays, cells):
            self.arrays, self.sizes = arrays, dict(dp_cells=cells)

        def as_struct(self):
            r = _lib.FrsResult()
            for k in _lib.RESULT_ARRAYS:
                setattr(r, k, self.arrays[k].ctypes.data_as(C.c_void_p))
            return r

    class OracleEngine:
        """Finds the tints of the batch by its read ids... the batch carries no names, so match on sizes."""
        def segment_batch(self, pb, prm):
            st = pb.struct
            tro = np.ctypeslib.as_array(C.cast(st.tint_read_off, C.POINTER(C.c_int32)), shape=(st.n_tints + 1,))
            ist = np.ctypeslib.as_array(C.cast(st.island_start, C.POINTER(C.c_int32)), shape=(st.n_islands,))
            tio = np.ctypeslib.as_array(C.cast(st.tint_island_off, C.POINTER(C.c_int32)), shape=(st.n_tints + 1,))
            chosen = []
            for k in range(st.n_tints):
                first_island = int(ist[tio[k]])
                n_reads = int(tro[k + 1] - tro[k])
                hit = [t for t in tints if t["intervals"][0][0] == first_island and len(t["reads"]) == n_reads]
                assert len(hit) == 1
                chosen.append(hit[0])
            _, arrays = oracle_result_arrays(chosen, oprm)
            return Result(arrays, 0)

    lib = _lib.load()
    monkeypatch.setattr(lib, "frs_device_count", lambda: 1, raising=False)
    monkeypatch.setattr(segment, "get_engine", lambda dev=0, lane=0: OracleEngine())
    out = str(tmp_path / "seg")
    prm = SegmentParams(oprm.sigma, oprm.tp, oprm.vf, oprm.mps, oprm.lo, oprm.ignore_ends)
    stats = segment.run_directory(src, out, prm, threads=2, gpus=1, batch_reads=400, progress=False, lanes=2,
                                  packed_segment=(mode == "packed"))
    assert stats["tints"] == len(tints) and stats["reads"] == sum(len(t["reads"]) for t in tints)
    assert len(by_key) == len(tints)
    if mode == "packed":  # --packed-segment: the binary twin next to the TSV files, same bytes when printed
        import json
        import shutil
        idx = json.load(open(os.path.join(out, "packed_segment", "index.json")))
        seen = []
        for b in idx["batches"]:
            ps = packed.PackedSegment(os.path.join(out, "packed_segment", b["file"]))
            assert ps.tints() == [(c, t) for c, t in b["tints"]]
            for k, (c, t) in enumerate(ps.tints()):
                assert ps.text(k) == open(os.path.join(out, c, "segment_%s_%d.tsv" % (c, t))).read()
                seen.append((c, t))
        assert sorted(seen) == sorted(by_key)
        shutil.rmtree(os.path.join(out, "packed_segment"))
    assert sha_dir(out) == manifest[name]["outputs"]


def _reference_read_segment():
    """freddie_cluster.read_segment of the unmodified reference (authoring container only; gurobipy and
    other solver-side imports are stubbed: read_segment is plain Python + re)."""
    ref = "/root/reference/py"
    if not os.path.isdir(ref):
        return None
    import importlib
    import types
    try:
        importlib.import_module("gurobipy")
    except Exception:
        stub = types.ModuleType("gurobipy")  # `from gurobipy import Model, GRB, quicksum, LinExpr` (:13)
        for attr in ("Model", "GRB", "quicksum", "LinExpr"):
            setattr(stub, attr, None)
        sys.modules["gurobipy"] = stub
    sys.path.insert(0, ref)
    try:
        return importlib.import_module("freddie_cluster").read_segment  # a failure here must be loud
    finally:
        sys.path.remove(ref)


@pytest.mark.parametrize("name", ["cfg1", "cfg2_flagsA", "degenerate", "plateau"])
def test_packed_segment_twin(name, golden_set, manifest, tmp_path, built_lib):
    """frs_packed_write_segment -> PackedSegment: text(t) is the reference's SEGMENT file byte for byte, and
    read_segment() equals what the reference's own freddie_cluster.read_segment parses from that text."""
    import hashlib
    from freddie_b200 import _lib, packed
    tints, flags, split_dir = golden_set(name)
    pb = _native_batch(split_dir, tints)
    _, arrays = oracle_result_arrays(tints, orc.Params(**flags_to_kwargs(flags)))
    res = _lib.FrsResult()
    for k in _lib.RESULT_ARRAYS:
        setattr(res, k, arrays[k].ctypes.data_as(C.c_void_p))

    class R:
        def as_struct(self):
            return res
    path = str(tmp_path / "seg.frsg")
    pb.write_packed_segment(R(), path)
    pb.close()
    ps = packed.PackedSegment(path)
    assert ps.tints() == [(t["chr"], t["id"]) for t in tints]
    ref_read_segment = _reference_read_segment()
    want_all = {}
    for k, (c, i) in enumerate(ps.tints()):
        text = ps.text(k)
        assert hashlib.sha256(text.encode()).hexdigest() == manifest[name]["outputs"]["%s/segment_%s_%d.tsv" % (c, c, i)]
        if ref_read_segment is not None:
            f = str(tmp_path / ("segment_%s_%d.tsv" % (c, i)))
            open(f, "w").write(text)
            want_all.update(ref_read_segment(f))
    if ref_read_segment is not None:
        got = ps.read_segment()
        assert got.keys() == want_all.keys()
        for tid in got:
            g, w = got[tid], want_all[tid]
            assert g["segs"] == w["segs"] and g["chr"] == w["chr"] and g["read_reps"] == w["read_reps"], tid
            assert g["reads"] == w["reads"], tid
    with pytest.raises(ValueError):
        bad = str(tmp_path / "bad.frsg")
        open(bad, "wb").write(open(path, "rb").read()[:200])
        packed.PackedSegment(bad)


@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_native_parser_on_fresh_random_tints(seed, tmp_path, built_lib, monkeypatch):
    """Seeds that are not golden sets: the native parser (sequential and chunked paths) against the Python
    packer on freshly generated tints, incl. a heavy-duplication tint (weighted read reps)."""
    from freddie_b200 import _lib, hostio, synth
    from freddie_b200.pack import pack_tints
    from freddie_b200.segment import _load_tint_py
    tints = synth.make_config(2, scale=0.004, seed=seed) + synth.make_config(3, scale=0.004, seed=seed)[:1]
    for k, t in enumerate(tints):  # unique (contig, id) pairs inside one directory
        t["id"] = k
        for r in t["reads"]:
            r["tint"] = k
    d = str(tmp_path / "s")
    synth.write_split_dir(tints, d)
    ref = pack_tints([_load_tint_py(d, t["chr"], t["id"]) for t in tints])
    for big in (None, "1"):
        if big:
            monkeypatch.setenv("FRS_PARSE_BIG_BYTES", big)
        pb = _native_batch(d, tints, threads=3)
        for k, v in ref.counts().items():
            assert getattr(pb.struct, k) == v, k
        for nm in _lib.BATCH_ARRAYS:
            arr = ref.arrays[nm]
            if arr.size == 0:
                continue
            ptr = C.cast(getattr(pb.struct, nm), C.POINTER(np.ctypeslib.as_ctypes_type(arr.dtype)))
            assert np.array_equal(np.ctypeslib.as_array(ptr, shape=arr.shape), arr), (nm, big)
        pb.close()
