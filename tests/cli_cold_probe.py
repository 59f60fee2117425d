#!/usr/bin/env python3
"""Dev tool (GPU box): writes the whole cfg2 SPLIT directory and times a cold CLI process over it with the
per-batch / per-phase profile on stderr (FRS_CLI_PROFILE, FRS_HOST_PROFILE)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freddie_b200 import synth  # noqa: E402

work = tempfile.mkdtemp(prefix="frs_cli_probe_")
sd = os.path.join(work, "split")
t0 = time.time()
made = synth.write_jobs(synth.config_jobs(2), sd, workers=min(32, os.cpu_count() or 1))
print("wrote %d tints / %d reads in %.1f s" % (len(made), sum(m[2] for m in made), time.time() - t0), flush=True)
env = dict(os.environ, FRS_CLI_PROFILE="1", FRS_HOST_PROFILE="1")
for k in range(2):
    t0 = time.time()
    r = subprocess.run([sys.executable, "-X", "importtime", "-c", "import freddie_b200.segment"], cwd=ROOT, capture_output=True, text=True)
    t_imp = time.time() - t0
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "freddie_b200.segment", "-s", sd, "-o", os.path.join(work, "out%d" % k), "-t",
                        str(os.cpu_count() or 1)], cwd=ROOT, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    dt = time.time() - t0
    print("== run %d: import alone %.2f s; CLI %.2f s (rc %d)" % (k, t_imp, dt, r.returncode))
    print("\n".join(l for l in r.stderr.splitlines() if "profile" in l or "Error" in l)[:6000], flush=True)
