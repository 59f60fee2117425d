"""SURVEY.md 8f-3 on the GPU: the CUDA path of the Gurobi-free front of freddie_cluster.py (frs_cprep_*, through the
C ABI) against the digests the UNMODIFIED reference functions produced (tests/golden/cluster_prep.json: 71 tints x
3 settings) and against the oracle on seeded random structures.  Everything here is integer / list work: bit-exact."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, flags_to_kwargs
from oracle import cluster_prep_oracle as cpo
from oracle import segment_oracle as orc
from test_oracle_cluster_prep import SETTINGS, _read_segment_text

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "cluster_prep.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def eng():
    from freddie_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def cprep():
    from freddie_b200.cluster_prep import ClusterPrep
    c = ClusterPrep(0)
    yield c
    c.close()


def _digest(res):
    return hashlib.sha256(cpo.canonical(res["I"], res["C"], res["FL"], res["cat"], res["garbage_cost"], res["gaps"],
                                        res["partitions"]).encode()).hexdigest()


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "dup_heavy", "degenerate", "plateau", "cfg3_mini"])
def test_segment_then_cluster_prep_equals_reference_digests(name, gold, eng, cprep):
    """SPLIT tints -> CUDA segment stage -> its result arrays -> CUDA cluster prep, all tints of a golden set in
    one batch; every tint's canonical serialisation must hash to what the reference's preprocess_ilp +
    partition_reads gave on the reference's own SEGMENT file (whose digest is checked as well)."""
    from freddie_b200 import synth
    from freddie_b200.cluster_prep import batch_from_segment
    from freddie_b200.engine import SegmentParams, format_tint
    from freddie_b200.pack import pack_tints
    tints, flags = synth.make_golden_set(name)
    o = orc.Params(**flags_to_kwargs(flags))
    batch = pack_tints(tints)
    res = eng.segment_batch(batch, SegmentParams(o.sigma, o.tp, o.vf, o.mps, o.lo, o.ignore_ends))
    cb = batch_from_segment(batch.arrays, res.arrays)
    keys = ["%s/%d" % (t["chr"], t["id"]) for t in tints]
    assert sorted(keys) == sorted(gold[name])
    for t, key in enumerate(keys):
        assert hashlib.sha256(format_tint(batch, res, t).encode()).hexdigest() == gold[name][key]["segment_sha256"], key
    launches = 0
    for model, mx in SETTINGS:
        out = cprep.run(cb, mx)
        launches += out.sizes["launches"]
        for t, key in enumerate(keys):
            g = gold[name][key]
            r = out.tint(t, model)
            assert len(r["cat"]) == g["reps"] and r["I"].shape[1] == g["segments"], key
            assert len(r["partitions"]) == g["partitions/%s/%d" % (model, mx)], (key, model, mx)
            assert _digest(r) == g["%s/%d" % (model, mx)], (key, model, mx)
            assert r["edges_after"] <= r["edges_before"]
    assert launches > 0


def test_mirror_functions_fill_the_reference_fields(cprep):
    """preprocess_ilp / partition_reads with the reference's signatures on read_segment-style dicts: the fields
    they leave equal what the oracle computes from the same dicts."""
    from freddie_b200 import synth
    from freddie_b200 import cluster_prep as cp
    tints, flags = synth.make_golden_set("cfg2_small")
    prm = orc.Params(**flags_to_kwargs(flags))
    for t in copy.deepcopy(tints)[:6]:
        orc.segment_tint(t, prm)
        tint = _read_segment_text(orc.format_segment(t))
        want = cpo.cluster_prep(copy.deepcopy(tint), "constant", 7)
        cp.preprocess_ilp(tint, dict(recycle_model="constant"))
        cp.partition_reads(tint, 7)
        U = len(tint["read_reps"])
        got = dict(I=np.array([tint["ilp_data"]["I"][i] for i in range(U)], dtype=np.uint8).reshape(U, -1),
                   C=np.array([tint["ilp_data"]["C"][i] for i in range(U)], dtype=np.uint8).reshape(U, -1),
                   FL=np.array([tint["ilp_data"]["FL"][i] for i in range(U)]).reshape(U, 2),
                   cat=[tint["reads"][idxs[0]]["poly_tail_category"] for idxs in tint["read_reps"]],
                   garbage_cost=tint["ilp_data"]["garbage_cost"],
                   gaps=[tint["reads"][idxs[0]]["gaps"] for idxs in tint["read_reps"]], partitions=tint["partitions"])
        assert _digest(got) == cpo.digest_of(want)
        for idxs in tint["read_reps"]:
            for r in idxs:
                assert tint["reads"][r]["gaps"] is tint["reads"][idxs[0]]["gaps"]
        with pytest.raises(AttributeError):
            cp.preprocess_ilp(tint, dict(recycle_model="exons"))


def _random_tint(rng, M, n_reads, dup):
    """A read_segment-style tint with random rows (all-zero rows, every poly-tail category, gaps around the
    threshold of 10, duplicated keys)."""
    reads = []
    protos = []
    for _ in range(max(1, n_reads // dup)):
        row = (rng.random(M) < rng.choice([0.0, 0.2, 0.5, 0.8, 1.0])).astype(int)
        data = [int(x) if x else int(rng.choice([0, 2])) for x in row]
        ones = np.flatnonzero(row)
        gaps = {}
        for a, b in zip(ones[:-1], ones[1:]):
            if rng.random() < 0.4:
                gaps[(int(a), int(b))] = int(rng.choice([0, 3, 10, 11, 25, 140]))
        poly = {}
        k = rng.random()
        if k < 0.25:
            poly["S" + str(rng.choice(["A", "T"]))] = (int(rng.choice([5, 10, 11, 30])), int(rng.choice([0, 10, 11, 40])))
        elif k < 0.5:
            poly["E" + str(rng.choice(["A", "T"]))] = (int(rng.choice([5, 10, 11, 30])), int(rng.choice([0, 10, 11, 40])))
        elif k < 0.6:
            poly["SA"] = (30, 12)
            poly["ET"] = (30, 12)
        protos.append((data, gaps, poly))
    for i in range(n_reads):
        data, gaps, poly = protos[int(rng.integers(len(protos)))]
        reads.append(dict(id=i, name="r%d" % i, chr="c", strand="+", tint=0, data=list(data), gaps=dict(gaps),
                          softclip={"SSC": int(rng.integers(0, 50)), "ESC": int(rng.integers(0, 50))}, poly_tail=dict(poly)))
    tint = dict(id=0, chr="c", segs=[(10 * j, 10 * j + 10, 10) for j in range(M)], reads=reads, read_reps={})
    for i, r in enumerate(reads):  # read_segment's key (:154-160); gaps in the file's (sorted string) order
        internal = sorted(("%d-%d:%d" % (a, b, v), v) for (a, b), v in r["gaps"].items())
        poly = sorted(("%s_%d:%d" % (k, v[0], v[1]), k, v) for k, v in r["poly_tail"].items())
        key = "".join(str(d) for d in r["data"]).replace("2", "0")
        key += "".join(".{}".format(v if v > 10 else 0) for _, v in internal)
        key += "".join(".{}{}".format(k[0], v[1] if v[1] > 10 else 0) for _, k, v in poly)
        tint["read_reps"].setdefault(key, []).append(i)
    tint["read_reps"] = list(tint["read_reps"].values())
    return tint


@pytest.mark.parametrize("seed,M,n_reads,dup,mx", [(1, 1, 40, 2, 1000), (2, 3, 200, 3, 5), (3, 4, 300, 2, 1000), (4, 33, 400, 2, 50),
                                                   (5, 70, 600, 4, 16), (6, 12, 1500, 1, 100), (7, 64, 64, 1, 3)])
def test_random_structures_equal_oracle(seed, M, n_reads, dup, mx, cprep):
    from freddie_b200.cluster_prep import batch_from_tints
    rng = np.random.default_rng(seed)
    tints = [_random_tint(rng, M, n_reads, dup), _random_tint(rng, max(1, M // 2), 1, 1), _random_tint(rng, M + 3, n_reads // 2, dup)]
    out = cprep.run(batch_from_tints(tints), mx)
    for t, tint in enumerate(tints):
        want = cpo.cluster_prep(copy.deepcopy(tint), "constant", mx)
        got = out.tint(t, "constant")
        assert got["read_reps"] == [list(x) for x in tint["read_reps"]], (seed, t)
        assert got["n_structs"] == len(want["structs"])
        assert (got["edges_before"], got["edges_after"]) == (want["edges_before"], want["edges_after"])
        assert _digest(got) == cpo.digest_of(want), (seed, t)


def test_errors_and_empty_batches(cprep):
    from freddie_b200 import _lib
    from freddie_b200.cluster_prep import batch_from_tints
    rng = np.random.default_rng(9)
    tint = _random_tint(rng, 5, 10, 1)
    with pytest.raises(ZeroDivisionError):
        cprep.run(batch_from_tints([tint]), 0)
    bad = batch_from_tints([tint])
    bad["digits"] = bad["digits"].copy()
    bad["digits"][3] = ord("7")
    with pytest.raises(_lib.FrsError):
        cprep.run(bad, 10)
    empty = dict(id=1, chr="c", segs=[(0, 5, 5)], reads=[], read_reps=[])
    out = cprep.run(batch_from_tints([empty, tint, empty]), 10)
    assert out.tint(0)["partitions"] == [] and out.tint(2)["cat"] == []
    assert _digest(out.tint(1)) == cpo.digest_of(cpo.cluster_prep(copy.deepcopy(tint), "constant", 10))
    out = cprep.run(batch_from_tints([]), 10)
    assert out.sizes["n_reps"] == 0 and out.sizes["n_parts"] == 0
