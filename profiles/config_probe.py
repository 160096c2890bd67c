#!/usr/bin/env python
"""Render throughput of BASELINE.json configs 3, 4 and 5 (device-resident frames/s), tiled path against the general A-buffer path.

    python profiles/config_probe.py > profiles/r01b_configs.txt      (needs a GPU)

C4: 512x512, 2500 flat rectangles per key frame (one chain per blob group), density 2, linear motion + cosine fading,
128 frames.  C3: 1024x1024, 2 key frames, fluid (10 MPM steps per frame) + perlin fading + feather 2, 16 frames (stateful
particle path, one frame at a time; the tiled / general switch does not apply).  C5: 4096x4096, 8 cyclic key frames, 16.7 M atoms, spline motion, 64 of the 512 frames."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomorph_b200 import engine as eng   # noqa: E402
from atomorph_b200 import scenes          # noqa: E402
import torch                              # noqa: E402


def run(name, images, params, times, prep):
    for tiled in ("1", "0"):
        os.environ["AMX_RENDER_TILED"] = tiled
        e = eng.Engine(0, **params)
        e.load_images(images)
        prep(e)
        e.render_prepare()
        n, size = len(times), images[0].shape[0]
        out = torch.empty((n, size, size), dtype=torch.int32, device="cuda:0")
        for _ in range(2):
            e.render_into(times, out.data_ptr(), True)
        e.sync()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            e.render_into(times, out.data_ptr(), True)
        e.sync()
        dt = (time.perf_counter() - t0) / reps
        print("%s  %s path: %8.1f frames/s (%.1f us/frame)  paths %s  tiled %s" % (name, "tiled  " if tiled == "1" else "general", n / dt, 1e6 * dt / n,
              e.render_path_frames(), e.render_tiled_stats()), flush=True)
        del out, e
        torch.cuda.empty_cache()


def prep_c4(e):
    e.blobify(); e.match_init(); e.match_rounds(2000); e.init_chains(); e.swap_rounds(400)


def prep_c5(e):
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    e.swap_rounds(256)


run("C4 512^2 2500 blobs d2 ", scenes.rect_blobs(512, 2500, frames=2, seed=11, min_side=2, max_side=20),
    dict(seed=1, motion=eng.LINEAR, fading=eng.COSINE, density=2, blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3, threads=0, cycle_length=1000),
    np.array([f / 128.0 for f in range(128)]), prep_c4)
run("C5 4096^2 8 key frames ", scenes.rotating_shapes(4096, 8), dict(seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=1000),
    np.array([f / 512.0 for f in range(64)]), prep_c5)


def prep_c3(e):
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    e.swap_rounds(512)


os.environ["AMX_RENDER_TILED"] = "1"
e = eng.Engine(0, seed=1, motion=eng.SPLINE, fading=eng.PERLIN, feather=2, fluid=10, threads=0, cycle_length=1000)
e.load_images(scenes.square_to_disc(1024))
prep_c3(e)
out = torch.empty((16, 1024, 1024), dtype=torch.int32, device="cuda:0")
times = np.array([f / 64.0 for f in range(16)])
e.render_into(times[:2], out.data_ptr(), True)
e.sync()
t0 = time.perf_counter()
e.render_into(times, out.data_ptr(), True)
e.sync()
dt = time.perf_counter() - t0
print("C3 1024^2 fluid 10 + perlin + feather 2: %8.1f frames/s (%.1f us/frame)" % (16 / dt, 1e6 * dt / 16), flush=True)
