#!/usr/bin/env python
"""Cost of the C2 chain table (1024^2 square -> disc, 1 048 576 atoms) vs proposals per atom, uniform partners only
against every n-th epoch pairing spatial neighbours (amx_set_swap_locality).  Prints one line per checkpoint.

    python profiles/locality_probe.py [size] > profiles/r01b_locality.txt      (needs a GPU)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomorph_b200 import engine as eng   # noqa: E402
from atomorph_b200 import scenes          # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
images = scenes.square_to_disc(size)
print("# %dx%d square -> disc, cost after N rounds of %d proposals (column 1); c_opt(1024^2) ~ 6.8e14 (SURVEY.md section 8c)" % (size, size, size * size // 2))
for every in (0, 4, 2, 1):
    e = eng.Engine(0, seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=1000)
    e.load_images(images)
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    e.set_swap_locality(every)
    c0 = e.cost()
    done = 0
    line = ["locality every %d:" % every if every else "uniform only:    ", "start %.4g" % c0]
    t0 = time.perf_counter()
    for target in (512, 1024, 2048, 4096, 8192, 16384, 24576):
        e.swap_rounds(target - done, column=1, want_stats=False)
        done = target
        line.append("%d: %.5g" % (target, e.cost()))
    e.sync()
    line.append("(%.2f s)" % (time.perf_counter() - t0))
    print("  ".join(line), flush=True)
