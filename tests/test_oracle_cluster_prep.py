"""The oracle of the NEXT hot-path row (freddie_cluster.py's Gurobi-free front) against the
digests the unmodified reference functions produced (oracle/pin_cluster_prep.py -> tests/golden/
cluster_prep.json).  CPU only; there is no CUDA path for this row yet."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, flags_to_kwargs
from oracle import cluster_prep_oracle as cpo
from oracle import segment_oracle as orc

SETTINGS = [("constant", 1000), ("constant", 7), ("relative", 50)]


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "cluster_prep.json")) as fh:
        return json.load(fh)


def _segment_tints(name, tmp_path):
    """{key: read_segment-style tint} of a golden set, through the packed SEGMENT twin's reader (itself held
    to the reference's read_segment in tests/test_host.py) -- no regexes, no reference needed."""
    from freddie_b200 import synth
    tints, flags = synth.make_golden_set(name)
    prm = orc.Params(**flags_to_kwargs(flags))
    out = {}
    for t in copy.deepcopy(tints):
        orc.segment_tint(t, prm)
        out["%s/%d" % (t["chr"], t["id"])] = (orc.format_segment(t), t)
    return out


def _read_segment_text(text):
    """Plain parser of a SEGMENT file into the structure freddie_cluster.read_segment builds (:119-172)."""
    import re
    tint = None
    for line in text.splitlines():
        if line.startswith("#"):
            c, i, pos = line[1:].split("\t")
            pos = [int(x) for x in pos.split(",")]
            tint = dict(id=int(i), chr=c, segs=[(s, e, e - s) for s, e in zip(pos[:-1], pos[1:])], read_reps={}, reads=[])
            continue
        rid, name, c, strand, cid, data, gaps = line.split("\t")
        internal = re.findall(r"(\d+)-(\d+):(\d+),", gaps)
        poly = re.findall(r"([ES][AT])_(\d+):(\d+),", gaps)
        read = dict(id=int(rid), name=name, chr=c, strand=strand, tint=int(cid), data=[int(d) for d in data],
                    gaps={(int(g[0]), int(g[1])): int(g[2]) for g in internal},
                    softclip={s[0]: int(s[1]) for s in re.findall(r"([ES]SC):(\d+),", gaps)},
                    poly_tail={p[0]: (int(p[1]), int(p[2])) for p in poly})
        key = data.replace("2", "0") + "".join(".{}".format(g[2] if int(g[2]) > 10 else 0) for g in internal)
        key += "".join(".{}{}".format(p[0][0], p[2] if int(p[2]) > 10 else 0) for p in poly)
        tint["reads"].append(read)
        tint["read_reps"].setdefault(key, []).append(len(tint["reads"]) - 1)
    tint["read_reps"] = list(tint["read_reps"].values())
    return tint


SLOW = [] if not os.environ.get("FRS_SLOW_TESTS") else ["cfg3_mini"]  # 2 x 3 000 reps x 1 000 segments: minutes


@pytest.mark.parametrize("name", ["cfg1", "cfg2_small", "dup_heavy", "degenerate", "plateau"] + SLOW)
def test_cluster_prep_oracle_equals_reference_digests(name, gold, tmp_path):
    segs = _segment_tints(name, tmp_path)
    assert sorted(segs) == sorted(gold[name])
    for key, (text, _) in segs.items():
        g = gold[name][key]
        assert hashlib.sha256(text.encode()).hexdigest() == g["segment_sha256"], key
        tint = _read_segment_text(text)
        assert len(tint["read_reps"]) == g["reps"] and len(tint["segs"]) == g["segments"]
        structs = pruned = None
        for model, mx in SETTINGS:
            res = cpo.preprocess(copy.deepcopy(tint), model)
            if structs is None:  # the graph does not depend on the recycle model or the piece size
                structs = cpo.unique_structures(res["I"], res["FL"], res["cat"])
                pruned = cpo.prune(cpo.compatibility_matrix(structs))
            res["partitions"] = cpo.partitions(structs, pruned, mx)
            assert len(res["partitions"]) == g["partitions/%s/%d" % (model, mx)], (key, model, mx)
            assert cpo.digest_of(res) == g["%s/%d" % (model, mx)], (key, model, mx)


def test_vectorised_pair_test_equals_the_scalar_one(tmp_path):
    """numpy compatibility matrix == the reference's expressions pair by pair, incl. rows without a 1
    (f = -1: Python's negative slice start) and every poly-tail category."""
    rng = np.random.default_rng(5)
    for M in (1, 2, 3, 4, 7, 12):
        structs = []
        for _ in range(60):
            row = tuple(int(x) for x in (rng.random(M) < rng.choice([0.0, 0.3, 0.7, 1.0])))
            lo, hi = cpo.find_segment_read(np.array(row))
            cat = str(rng.choice(["N", "S", "E"]))
            if cat == "S":
                lo = 0
            if cat == "E":
                hi = M - 1
            structs.append(((row, (lo, hi, cat)), [len(structs)]))
        assert np.array_equal(cpo.compatibility_matrix(structs), cpo.compatibility_matrix_scalar(structs)), M


def test_broken_recycle_models_raise_like_the_reference():
    tint = dict(segs=[(0, 5, 5)], read_reps=[[0]], reads=[dict(data=[1], gaps={}, poly_tail={})])
    for model in ("exons", "introns"):
        with pytest.raises(AttributeError):
            cpo.preprocess(tint, model)
