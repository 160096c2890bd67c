#!/usr/bin/env python
"""Short render-only run of BASELINE config 2 for ncu / A-B timing (needs a GPU).

    python profiles/render_probe.py [--reps 3] [--match-rounds 24576] [--size 1024] [--frames 64]

Matches the table (untimed), renders `frames` frames `reps` times device-resident and prints the per-kernel device times
(CUDA event pairs, amx_kernel_times) and frames/s.  Environment switches of the library (AMX_TILE_V1, AMX_BIN_V1, ...)
select kernel generations for comparisons."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomorph_b200 import engine as eng   # noqa: E402
from atomorph_b200 import scenes          # noqa: E402
import torch                              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--match-rounds", type=int, default=24576)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--tag", default="")
a = ap.parse_args()
e = eng.Engine(0, seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=100000)
e.load_images(scenes.square_to_disc(a.size))
e.step(8)
e.swap_rounds(a.match_rounds, want_stats=False)
e.render_prepare()
times = np.arange(a.frames) / float(a.frames)
out = torch.empty((a.frames, a.size, a.size), dtype=torch.int32, device="cuda:0")
e.render_into(times, out.data_ptr(), True)
e.sync()
t0 = time.perf_counter()
for _ in range(a.reps):
    e.render_into(times, out.data_ptr(), True)
e.sync()
dt = (time.perf_counter() - t0) / a.reps
e.kernel_times(True)
for _ in range(a.reps):
    e.render_into(times, out.data_ptr(), True)
kt = e.kernel_times(False)
chk = int(out.to(torch.int64).sum().item())
print("%s %.1f us/frame (%.0f frames/s)  bin %.1f us/launch  tile %.1f us/launch  checksum %d  stats %s %s" % (
    a.tag, 1e6 * dt / a.frames, a.frames / dt, 1000 * kt[0]["ms"] / max(1, kt[0]["launches"]), 1000 * kt[1]["ms"] / max(1, kt[1]["launches"]),
    chk, e.render_stats(), e.render_tiled_stats()), flush=True)
