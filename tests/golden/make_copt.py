#!/usr/bin/env python
"""BASELINE.md section 3 cost protocol: the reference's converged transport cost for BASELINE config 2.

    python tests/golden/make_copt.py [--size 1024] [--out tests/golden/c2_copt.json]

Runs the UNMODIFIED reference (oracle/_ref/libamref.so, built from /root/reference by oracle/Makefile)
on the C2 scene (atomorph_b200.scenes.square_to_disc, the same arrays bench.py feeds the GPU) through
its OWN pipeline -- blobify, unify, match, init_morph (thread.cpp:225-891) -- with seed 1 and
threads = 0 (bit-reproducible), then drives morph_asynch (thread.cpp:990-1041)
with iterate() to 1000*W, 2000*W and 4000*W proposals.  After each leg the cost is recomputed from the
chain table (thread.cpp:1109-1125 restated in ref_harness.cpp: the reference's own get_energy() is an
uninitialised sum).  With uniform partners the reference follows  cost(ppa) = c_opt (1 + k / ppa)
(ppa = proposals per atom), so c_opt is the intercept of a least-squares line through (1/ppa, cost);
that fitted c_opt is the "reference converged cost" the GPU matcher must reach within 1 %.

CPU only (about 10 minutes at 1024^2); the result is committed as a small JSON fixture.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c2_copt.json"))
    ap.add_argument("--legs", default="1000,2000,4000")
    args = ap.parse_args()
    from atomorph_b200 import scenes
    from oracle import amref

    images = scenes.square_to_disc(args.size)
    m = amref.RefMorph(seed=1, motion=amref.SPLINE, fading=amref.COSINE)
    for k, im in enumerate(images):
        m.add_image(k, im)
    m.set_resolution(args.size, args.size)
    t0 = time.time()
    st = m.run_until(amref.STATE_ATOM_MORPHING, 20000, 0)
    assert st == amref.STATE_ATOM_MORPHING, st
    t_pre = time.time() - t0
    ch = m.chains()
    assert len(ch) == 1 and ch[0]["height"] == 2
    W = int(ch[0]["width"])
    cost0 = m.true_cost()
    print("pre-stages %.1f s, W = %d, initial cost %.6e" % (t_pre, W, cost0), flush=True)
    m.set(threads=0, cycle_length=W)          # one iterate step = W proposals (thread.cpp:1051-1058)
    m.sync()
    legs = [int(v) for v in args.legs.split(",")]
    done = 0
    rows = []
    for ppa in legs:
        t0 = time.time()
        m.iterate(ppa - done)
        done = ppa
        m.sync()                               # refresh the morph-side mirror the cost is read from
        c = m.true_cost()
        rows.append(dict(proposals_per_atom=ppa, proposals=ppa * W, cost=c, seconds=time.time() - t0))
        print("ppa %d: cost %.6e (%.1f s)" % (ppa, c, rows[-1]["seconds"]), flush=True)
    x = np.array([1.0 / r["proposals_per_atom"] for r in rows])
    y = np.array([r["cost"] for r in rows])
    slope, c_opt = np.polyfit(x, y, 1)
    out = dict(scene="scenes.square_to_disc(%d), seed 1, threads 0, reference's own pipeline" % args.size,
               size=args.size, atoms=W, cost_initial=cost0, legs=rows, c_opt=float(c_opt), k=float(slope / c_opt),
               fit="least squares of cost against 1/ppa over the legs; c_opt = intercept, k = slope / c_opt",
               source="tests/golden/make_copt.py on oracle/_ref/libamref.so (unmodified reference)")
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
