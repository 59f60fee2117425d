"""SURVEY.md 8f-4 on the GPU: the CUDA path of freddie_split.py's tint construction (frs_split_*, through the C ABI)
against the digests of the UNMODIFIED reference functions (tests/golden/split_tints.json) and against the oracle on
fresh seeded groups.  Integer / list work: bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import split_tints_oracle as sto

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "split_tints.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def ctx():
    from freddie_b200.split_tints import SplitTints
    c = SplitTints(0)
    yield c
    c.close()


def test_every_golden_group_equals_reference_digest_one_by_one(gold, ctx):
    for name, kw in sorted(sto.GOLDEN_GROUPS.items()):
        group = sto.make_group(**kw)
        tints, info, _ = ctx.run([group])
        assert len(tints[0]) == gold[name]["tints"], name
        assert sto.digest_of(tints[0]) == gold[name]["sha256"], name
        assert info["launches"] > 0


def test_all_golden_groups_in_one_batch(gold, ctx):
    names = sorted(sto.GOLDEN_GROUPS)
    groups = [sto.make_group(**sto.GOLDEN_GROUPS[n]) for n in names]
    tints, info, _ = ctx.run(groups + [[]] + groups[::-1])  # an empty group in the middle
    assert tints[len(names)] == []
    for k, n in enumerate(names):
        assert sto.digest_of(tints[k]) == gold[n]["sha256"], n
        assert sto.digest_of(tints[2 * len(names) - k]) == gold[n]["sha256"], n
    assert info["n_big"] >= 6


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_fresh_groups_equal_oracle(seed, ctx):
    rng = np.random.default_rng(seed)
    groups = []
    for k in range(12):
        groups.append(sto.make_group(seed=int(rng.integers(1 << 30)), n_loci=int(rng.integers(1, 9)),
                                     reads_per_locus=int(rng.integers(1, 120)), chain=float(rng.random()),
                                     big=bool(k % 4 == 3)))
    for thresholds in ((100, 1500), (5, 40), (1, 1)):
        tints, _, _ = ctx.run(groups, *thresholds)
        for g, group in enumerate(groups):
            want = sto.transcriptional_intervals(group, *thresholds)
            assert sto.canonical(tints[g]) == sto.canonical(want), (seed, g, thresholds)


def test_mirror_has_the_reference_signature_and_errors(ctx):
    from freddie_b200 import _lib
    from freddie_b200 import split_tints as st
    group = sto.make_group(**sto.GOLDEN_GROUPS["small"])
    reads = [dict(id=i, name="r%d" % i, contig="c", strand="+", simple_tints=[], tint=None,
                  intervals=[(s, e, 0, e - s, [(0, e - s)]) for s, e in ivs]) for i, ivs in enumerate(group)]
    got = st.get_transcriptional_intervals(reads)
    want = sto.transcriptional_intervals(group)
    assert [(t["intervals"], t["rids"]) for t in got] == want
    with pytest.raises(_lib.FrsError):
        ctx.run([[[(10, 5)]]])
    with pytest.raises(_lib.FrsError):
        ctx.run([[[]]])
    assert ctx.run([])[0] == []
