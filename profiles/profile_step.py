#!/usr/bin/env python3
"""One step of the bench workload between cudaProfilerStart/Stop, for ncu (run under gpurun):

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -o gpurun_out/step python profiles/profile_step.py

Same workload, parameters and options as bench.py's `value` leg (resident batch, sequence planes in
HBM); the L2 is flushed before the profiled step as bench.py does between timed steps.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from freddie_b200 import _lib, synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

workload = os.environ.get("WORKLOAD", "cfg2")
scale = float(os.environ.get("SCALE", "1"))
steps = int(os.environ.get("STEPS", "1"))
tints = synth.make_config({"cfg2": 2, "cfg3": 3, "cfg4": 4}[workload], scale=scale, seed=2, workers=16)
batch = pack_tints(tints).pin()
eng = Engine(0)
if not os.environ.get("E2E"):
    eng.set_option(_lib.OPT_LAZY_SEQ, 0)
prm = SegmentParams()
for _ in range(3):
    eng.segment_batch(batch, prm, pinned=True)
eng.upload(batch)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
for _ in range(steps):
    if os.environ.get("E2E"):  # host buffers in, host results out: the prep kernels and the clip fetch are in the list
        eng.segment_batch(batch, prm, pinned=True)
    else:
        eng.run(prm)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("profiled %d step(s): %d reads, %d launches per step" % (steps, batch.n_reads, eng.launch_count()))
print("stats", eng.stats())
