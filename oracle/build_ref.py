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


def build_ref(quiet: bool = True) -> bool:
    """Compiles the reference script into ``oracle/_ref`` when ``/root/reference`` is present.
    Returns True if the artefact exists afterwards."""
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        py_compile.compile(REF_SRC, cfile=REF_PYC, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        with open(REF_SRC, "rb") as fh:
            sha = hashlib.sha256(fh.read()).hexdigest()
        with open(REF_META, "w") as fh:
            json.dump(dict(source=REF_SRC, sha256=sha, python="%d.%d.%d" % sys.version_info[:3],
                           artefact="freddie_segment.bin (py_compile of the unmodified source, optimize=0)"), fh)
        if not quiet:
            print("oracle/_ref: compiled %s (sha256 %s)" % (REF_SRC, sha[:16]))
    return available()


def available() -> bool:
    return os.path.exists(REF_PYC)


def command(split_dir: str, out_dir: str, threads: int, flags=()):
    """argv of the reference CLI (freddie_segment.py:53-110) on the compiled artefact."""
    return [sys.executable, "-W", "ignore", REF_PYC, "-s", split_dir, "-o", out_dir, "-t", str(threads)] + list(flags)


if __name__ == "__main__":
    ok = build_ref(quiet=False)
    sys.exit(0 if ok else 1)
