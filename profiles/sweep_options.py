#!/usr/bin/env python3
"""Dev tool (GPU box): stage times of the bench workload for a few values of a library option.

  python profiles/sweep_options.py poly_long 24 28 32 36 40
  python profiles/sweep_options.py slab_words 16 32 64 128
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402,F401

from freddie_b200 import _lib, synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

KEYS = {"poly_long": (_lib.OPT_POLY_LONG_CLASS, ("poly",)), "slab_words": (_lib.OPT_SLAB_WORDS, ("dp", "dp_solve"))}
name = sys.argv[1]
values = [int(v) for v in sys.argv[2:]]
key, stages = KEYS[name]
workload = os.environ.get("WORKLOAD", "cfg2")
tints = synth.make_config({"cfg2": 2, "cfg3": 3, "cfg4": 4}[workload], scale=float(os.environ.get("SCALE", "1")), seed=2,
                          workers=16)
batch = pack_tints(tints).pin()
eng = Engine(0)
eng.set_option(_lib.OPT_LAZY_SEQ, 0)
prm = SegmentParams()
eng.segment_batch(batch, prm, pinned=True)
eng.upload(batch)
eng.set_profiling(True)
for v in values:
    eng.set_option(key, v)
    best = None
    for _ in range(4):
        eng.run(prm)
        tm = {k: ms for k, ms, _ in eng.timings()}
        tot = sum(tm.values())
        s = sum(tm.get(k, 0.0) for k in stages)
        best = (s, tot) if best is None or s < best[0] else best
    print("%s=%d: %s %.3f ms, all stages %.3f ms, stats %s" % (name, v, "+".join(stages), best[0], best[1],
                                                               {k: x for k, x in eng.stats().items() if "poly" in k}))
