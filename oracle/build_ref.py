#!/usr/bin/env python3
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference segment stage as a runnable artefact.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product (``freddie_b200/``).

The reference is a Python script (``/root/reference/py/freddie_segment.py``; it needs only numpy and
scipy, both in the image).  ``/root/reference`` exists in the authoring container only, so ``build()``
byte-compiles the script *where it lies* into ``oracle/_ref/freddie_segment.bin`` (a ``.pyc`` image under a
name the snapshot keeps) -- a git-ignored
build output that travels to the GPU box with the snapshot like the built ``.so`` -- and records the
SHA-256 of the source it was compiled from.  No reference source is copied into the repository.

``python oracle/_ref/freddie_segment.bin -s SPLIT -o OUT -t N`` then runs the reference CLI exactly as
``python /root/reference/py/freddie_segment.py`` does (same argv, same multiprocessing pool); it is what
``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` time on the GPU box's host cores, and
what ``tests/`` may use as a second checker beside the oracle port.
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/py/freddie_segment.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_PYC = os.path.join(REF_DIR, "freddie_segment.bin")  # CPython bytecode (runs by its magic number; *.pyc does not travel)
REF_META = os.path.join(REF_DIR, "MANIFEST.json")
# the cluster stage's Gurobi-free front: read_segment / preprocess_ilp / partition_reads are
# called as functions of the compiled module (gurobipy is stubbed when absent: they never touch it)
REF_CLUSTER_SRC = "/root/reference/py/freddie_cluster.py"
REF_CLUSTER_PYC = os.path.join(REF_DIR, "freddie_cluster.bin")
# the split stage's tint construction: get_transcriptional_intervals / break_tint as functions of
# the compiled module (pysam is stubbed when absent: they never touch it)
REF_SPLIT_SRC = "/root/reference/py/freddie_split.py"
REF_SPLIT_PYC = os.path.join(REF_DIR, "freddie_split.bin")


def build_ref(quiet: bool = True) -> bool:
    """Compiles the reference script into ``oracle/_ref`` when ``/root/reference`` is present.
    Returns True if the artefact exists afterwards."""
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        py_compile.compile(REF_SRC, cfile=REF_PYC, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        with open(REF_SRC, "rb") as fh:
            sha = hashlib.sha256(fh.read()).hexdigest()
        meta = dict(source=REF_SRC, sha256=sha, python="%d.%d.%d" % sys.version_info[:3],
                    artefact="freddie_segment.bin (py_compile of the unmodified source, optimize=0)")
        if os.path.exists(REF_CLUSTER_SRC):
            import warnings
            with warnings.catch_warnings():  # the source has '\d' in plain strings (SyntaxWarning on 3.12)
                warnings.simplefilter("ignore")
                py_compile.compile(REF_CLUSTER_SRC, cfile=REF_CLUSTER_PYC, doraise=True, optimize=0,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            with open(REF_CLUSTER_SRC, "rb") as fh:
                meta["cluster"] = dict(source=REF_CLUSTER_SRC, sha256=hashlib.sha256(fh.read()).hexdigest(),
                                       artefact="freddie_cluster.bin")
        if os.path.exists(REF_SPLIT_SRC):
            py_compile.compile(REF_SPLIT_SRC, cfile=REF_SPLIT_PYC, doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            with open(REF_SPLIT_SRC, "rb") as fh:
                meta["split"] = dict(source=REF_SPLIT_SRC, sha256=hashlib.sha256(fh.read()).hexdigest(),
                                     artefact="freddie_split.bin")
        with open(REF_META, "w") as fh:
            json.dump(meta, fh)
        if not quiet:
            print("oracle/_ref: compiled %s (sha256 %s)" % (REF_SRC, sha[:16]))
    return available()


def available() -> bool:
    return os.path.exists(REF_PYC)


def cluster_available() -> bool:
    return os.path.exists(REF_CLUSTER_PYC)


_cluster_mod = None


def reference_cluster_module():
    """The unmodified freddie_cluster module from its compiled artefact (functions only; main() is not run)."""
    global _cluster_mod
    if _cluster_mod is None:
        import importlib.machinery
        import importlib.util
        import types
        try:
            import gurobipy  # noqa: F401
        except Exception:
            stub = types.ModuleType("gurobipy")  # `from gurobipy import Model, GRB, quicksum, LinExpr` (:13)
            for attr in ("Model", "GRB", "quicksum", "LinExpr"):
                setattr(stub, attr, None)
            sys.modules["gurobipy"] = stub
        import warnings
        loader = importlib.machinery.SourcelessFileLoader("freddie_cluster_ref", REF_CLUSTER_PYC)
        spec = importlib.util.spec_from_loader("freddie_cluster_ref", loader)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loader.exec_module(mod)
        _cluster_mod = mod
    return _cluster_mod


def split_available() -> bool:
    return os.path.exists(REF_SPLIT_PYC)


_split_mod = None


def reference_split_module():
    """The unmodified freddie_split module from its compiled artefact (functions only; main() is not run)."""
    global _split_mod
    if _split_mod is None:
        import importlib.machinery
        import importlib.util
        import types
        try:
            import pysam  # noqa: F401
        except Exception:
            stub = types.ModuleType("pysam")  # the module-level tables (:64-101) name the SAM CIGAR operation codes
            for code, name in enumerate(["CMATCH", "CINS", "CDEL", "CREF_SKIP", "CSOFT_CLIP", "CHARD_CLIP", "CPAD", "CEQUAL",
                                         "CDIFF", "CBACK"]):
                setattr(stub, name, code)
            sys.modules["pysam"] = stub
        loader = importlib.machinery.SourcelessFileLoader("freddie_split_ref", REF_SPLIT_PYC)
        spec = importlib.util.spec_from_loader("freddie_split_ref", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _split_mod = mod
    return _split_mod


def command(split_dir: str, out_dir: str, threads: int, flags=()):
    """argv of the reference CLI (freddie_segment.py:53-110) on the compiled artefact."""
    return [sys.executable, "-W", "ignore", REF_PYC, "-s", split_dir, "-o", out_dir, "-t", str(threads)] + list(flags)


if __name__ == "__main__":
    ok = build_ref(quiet=False)
    sys.exit(0 if ok else 1)
