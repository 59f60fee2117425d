"""The C-ABI library: loads, exports every symbol include/freddie_b200.h declares, fails loudly
without a device.  No compute calls here (CPU-only suite)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, have_gpu


def test_header_symbols_are_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "freddie_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(frs_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(built_lib, name), "missing export: " + name
    from freddie_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared


def test_abi_version_and_struct_sizes(built_lib):
    from freddie_b200 import _lib
    assert built_lib.frs_abi_version() == 2
    # layouts the header promises (x86-64 SysV): 3 doubles + 4 int32 + 3 pointers + 2 int32
    assert C.sizeof(_lib.FrsParams) == 3 * 8 + 4 * 4 + 3 * 8 + 2 * 4
    # 8 counts + n_seq_words + 22 arrays + (seq_edge_words, host_arena) + seq_edge, cigar16, riv_cig_n + (qe_from_cigar, reserved1)
    assert C.sizeof(_lib.FrsBatch) == 8 * 4 + 8 + 22 * 8 + 8 + 3 * 8 + 8
    assert C.sizeof(_lib.FrsResultSizes) == 7 * 8 + 2 * 4
    assert C.sizeof(_lib.FrsResult) == 7 * 8


@pytest.mark.skipif(have_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device(built_lib):
    from freddie_b200 import _lib
    from freddie_b200.engine import Engine
    assert built_lib.frs_device_count() == 0
    with pytest.raises(_lib.FrsError) as e:
        Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "freddie_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, fn), errors="replace").read()
                assert "oracle" not in src, os.path.join(base, fn)


def test_neighbour_stage_structs_and_no_fallback(built_lib):
    """frs_cluster_* / frs_split_* structs have the header's layout; without a device their contexts refuse to exist."""
    from freddie_b200 import _lib
    assert C.sizeof(_lib.FrsClusterBatch) == 2 * 4 + 8 * 8
    assert C.sizeof(_lib.FrsClusterSizes) == 7 * 8 + 2 * 4
    assert C.sizeof(_lib.FrsClusterResult) == 18 * 8
    assert C.sizeof(_lib.FrsSplitBatch) == 2 * 4 + 4 * 8 + 2 * 4
    assert C.sizeof(_lib.FrsSplitSizes) == 4 * 8 + 2 * 4
    assert C.sizeof(_lib.FrsSplitResult) == 6 * 8
    if not have_gpu():
        from freddie_b200.cluster_prep import ClusterPrep
        from freddie_b200.split_tints import SplitTints
        for cls in (ClusterPrep, SplitTints):
            with pytest.raises(_lib.FrsError) as e:
                cls(0)
            assert "no CPU fallback" in str(e.value)


def test_cluster_batch_builders_agree():
    """batch_from_tints (read_segment-style dicts) and batch_from_segment (arrays of the segment stage) describe the
    same batch: same digit rows per read, same heads, same gap records."""
    import copy
    import numpy as np
    from freddie_b200 import synth
    from freddie_b200.cluster_prep import batch_from_segment, batch_from_tints
    from freddie_b200.pack import pack_tints
    from helpers import oracle_result_arrays
    from oracle import segment_oracle as orc
    from test_oracle_cluster_prep import _read_segment_text
    tints, _ = synth.make_golden_set("cfg2_small")
    tints = tints[:5]
    otints, arrays = oracle_result_arrays(tints, orc.Params())
    batch = pack_tints(copy.deepcopy(tints))
    a = batch_from_segment(batch.arrays, arrays)
    b = batch_from_tints([_read_segment_text(orc.format_segment(t)) for t in otints])
    assert np.array_equal(a["tint_seg_n"], b["tint_seg_n"]) and np.array_equal(a["tint_read_off"], b["tint_read_off"])
    N = len(b["read_row"])
    for i in range(N):
        t = int(np.searchsorted(np.asarray(a["tint_read_off"]), i, side="right")) - 1
        M = int(a["tint_seg_n"][t])
        ra = a["digits"][int(a["tint_digit_off"][t]) + int(a["read_row"][i]) * M:][:M]
        rb = b["digits"][int(b["tint_digit_off"][t]) + int(b["read_row"][i]) * M:][:M]
        assert np.array_equal(ra, rb), i
        assert list(a["read_head"][8 * i:8 * i + 7]) == list(b["read_head"][8 * i:8 * i + 7]), i
        ga = sorted(map(tuple, np.asarray(a["gap_rec"][3 * int(a["read_gap_off"][i]):3 * int(a["read_gap_off"][i + 1])]).reshape(-1, 3).tolist()))
        gb = sorted(map(tuple, np.asarray(b["gap_rec"][3 * int(b["read_gap_off"][i]):3 * int(b["read_gap_off"][i + 1])]).reshape(-1, 3).tolist()))
        assert ga == gb, i
