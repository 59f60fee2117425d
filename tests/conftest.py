"""Shared fixtures.  ``-m "not gpu"`` covers the oracle against the golden vectors, the host logic and
the C-ABI surface; ``-m gpu`` holds the parity tests proper (CUDA path vs oracle through the C ABI)."""
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def sha_dir(d):
    out = {}
    for base, _, files in os.walk(d):
        for fn in sorted(files):
            p = os.path.join(base, fn)
            with open(p, "rb") as fh:
                out[os.path.relpath(p, d)] = hashlib.sha256(fh.read()).hexdigest()
    return dict(sorted(out.items()))


def digest(m):
    h = hashlib.sha256()
    for k, v in sorted(m.items()):
        h.update(k.encode())
        h.update(v.encode())
    return h.hexdigest()


def flags_to_kwargs(flags):
    kw = {}
    it = iter(flags)
    for f in it:
        if f == "--consider-ends":
            kw["consider_ends"] = True
        else:
            v = next(it)
            kw[{"-sd": "sigma", "-tp": "tp", "-vf": "vf", "-mps": "mps", "-lo": "lo"}[f]] = (
                int(v) if f in ("-mps", "-lo") else float(v))
    return kw


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


_SETS = {}


@pytest.fixture(scope="session")
def golden_set(tmp_path_factory, manifest):
    """golden_set(name) -> (tints, flags, split_dir): the seeded inputs, checked against the digest the
    reference outputs in the manifest were produced from."""
    from freddie_b200 import synth

    def get(name):
        if name not in _SETS:
            tints, flags = synth.make_golden_set(name)
            d = str(tmp_path_factory.mktemp("split_" + name))
            synth.write_split_dir(tints, d)
            assert digest(sha_dir(d)) == manifest[name]["input_digest"], (
                "generated inputs of %s differ from the ones the golden outputs were made from" % name)
            _SETS[name] = (tints, flags, d)
        return _SETS[name]
    return get


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from freddie_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def have_gpu():
    try:
        from freddie_b200 import _lib
        return _lib.load().frs_device_count() > 0
    except Exception:
        return False
