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
