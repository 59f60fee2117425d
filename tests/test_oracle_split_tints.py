"""The oracle of SURVEY.md 8f-4 (freddie_split.py's tint construction) against the digests the unmodified reference
functions produced (oracle/pin_split_tints.py -> tests/golden/split_tints.json).  CPU only."""
import json
import os

import pytest

from conftest import GOLDEN
from oracle import split_tints_oracle as sto


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "split_tints.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", sorted(sto.GOLDEN_GROUPS))
def test_split_tints_oracle_equals_reference_digests(name, gold):
    group = sto.make_group(**sto.GOLDEN_GROUPS[name])
    g = gold[name]
    assert (len(group), sum(len(r) for r in group)) == (g["reads"], g["intervals"])
    tints = sto.transcriptional_intervals(group)
    assert len(tints) == g["tints"]
    assert sto.digest_of(tints) == g["sha256"]


def test_golden_groups_reach_break_tint(gold):
    assert any(g["largest_tint_intervals"] >= 100 for g in gold.values())
    assert sum(g["tints"] for g in gold.values()) > 30


def test_touching_intervals_merge_and_thresholds():
    # s == running end merges (freddie_split.py:303 tests s > end); fewer than three reads: no tint (:345)
    reads = [[(10, 20), (30, 40)], [(20, 30)], [(40, 50)]]
    assert sto.transcriptional_intervals(reads) == [([(10, 50)], [0, 1, 2])]
    assert sto.transcriptional_intervals(reads[:2]) == []
    reads = [[(10, 20)], [(21, 30)], [(10, 15)], [(12, 18)]]
    assert sto.transcriptional_intervals(reads) == [([(10, 20)], [0, 2, 3])]
