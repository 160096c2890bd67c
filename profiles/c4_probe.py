#!/usr/bin/env python
"""BASELINE config 4 (512^2, 2 500 rectangles per key frame, density 2: several chains, the general A-buffer path) -- frames/s over
the 128 frames with a checksum of all pixels (A/B runs of the general path must agree on it), then per segment of the morph the
time per frame, the listed (generic) positions and overflow records per frame and the device time of the scatter group
(k_scatter + k_ovf_alloc + k_ovf_place) and of the gather group (k_gather_pixel + k_resolve_list + k_resolve_heavy) per batch.

    python profiles/c4_probe.py            (needs a GPU; AMX_LIB=... selects a tuning build)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomorph_b200 import engine as eng, scenes
import torch
os.environ["AMX_RENDER_TILED"] = os.environ.get("AMX_RENDER_TILED", "1")
e = eng.Engine(0, seed=1, motion=eng.LINEAR, fading=eng.COSINE, density=2, blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3, threads=0, cycle_length=1000)
imgs = scenes.rect_blobs(512, 2500, frames=2, seed=11, min_side=2, max_side=20)
e.load_images(imgs)
e.blobify(); e.match_init(); e.match_rounds(2000); e.init_chains(); e.swap_rounds(400)
e.render_prepare()
times = np.array([f / 128.0 for f in range(128)])
out = torch.empty((128, 512, 512), dtype=torch.int32, device="cuda:0")
for _ in range(2): e.render_into(times, out.data_ptr(), True)
e.sync(); t0 = time.perf_counter()
for _ in range(3): e.render_into(times, out.data_ptr(), True)
e.sync(); dt = (time.perf_counter() - t0) / 3
print("C4 %.1f frames/s (%.1f us/frame) checksum %d stats %s paths %s" % (128 / dt, 1e6 * dt / 128, int(out.to(torch.int64).sum().item()), e.render_stats(), e.render_path_frames()), flush=True)
for lo, hi in ((0, 8), (8, 16), (24, 32), (48, 56), (56, 64), (64, 72), (72, 80), (96, 104), (120, 128)):
    tt = times[lo:hi]
    e.render_into(tt, out.data_ptr(), True); e.sync(); t0 = time.perf_counter()
    for _ in range(5): e.render_into(tt, out.data_ptr(), True)
    e.sync(); dtm = time.perf_counter() - t0
    st1 = e.render_stats(); e.render_into(tt, out.data_ptr(), True); e.sync(); st2 = e.render_stats()
    e.kernel_times(True); e.render_into(tt, out.data_ptr(), True); kt = e.kernel_times(False)
    print("  frames %d-%d: %.1f us/frame  generic/frame %d overflow/frame %d  scatter %.1f gather+list %.1f us/batch" % (lo, hi, 1e6 * dtm / 5 / len(tt), (st2["generic"] - st1["generic"]) / len(tt), (st2["overflow"] - st1["overflow"]) / len(tt), 1000 * kt[0]["ms"] / max(1, kt[0]["launches"]), 1000 * kt[1]["ms"] / max(1, kt[1]["launches"])), flush=True)
