"""Generates the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libamref.so, built from /root/reference by oracle/Makefile).  The reference ships no golden
vectors of its own, so these are the known answers that pin oracle/am_oracle.cpp
and, through it and directly, the CUDA path.

    python tests/golden/make_golden.py          # only where /root/reference exists (build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import amref                      # noqa: E402
from atomorph_b200 import scenes              # noqa: E402
from atomorph_b200 import engine as eng       # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def pure():
    rng = np.random.default_rng(20241017)
    L = amref.lib()
    cols = rng.integers(0, 2 ** 32, 3000, dtype=np.uint64).astype(np.uint32)
    cols[:6] = [0, 0xffffffff, 0xff0000ff, 0xff00ff00, 0xffff0000, 0xff808080]
    hsp = np.array([L.amref_rgb_to_hsp(int(c)) for c in cols], dtype=np.uint32)
    back = np.array([L.amref_hsp_to_rgb(int(c)) for c in cols], dtype=np.uint32)      # arbitrary HSP bytes -> RGB
    cd = np.array([L.amref_color_distance(int(a), int(b)) for a, b in zip(cols[:-1], cols[1:])])
    pts = (rng.integers(0, 2 ** 48, 2000, dtype=np.uint64))
    pd = np.array([L.amref_point_distance(int(a), int(b)) for a, b in zip(pts[:-1], pts[1:])], dtype=np.uint64)
    seeds = np.array([0, 1, 2, 7, 12345], dtype=np.uint32)
    nxy = rng.uniform(-3.0, 9.0, size=(400, 2))
    noise = np.array([[L.amref_octave_noise(int(s), x, y, 8) for x, y in nxy] for s in seeds])
    sp_n = rng.integers(2, 7, 300)
    sp_ctrl = rng.uniform(0, 200, size=(300, 6, 2))
    sp_t = rng.uniform(0, 0.999999, 300)
    sp_out = np.zeros((300, 2))
    for i in range(300):
        sp_out[i] = amref.spline_point(sp_ctrl[i, :sp_n[i], 0], sp_ctrl[i, :sp_n[i], 1], sp_t[i])
    m = amref.RefMorph(fading=amref.COSINE)
    w = rng.uniform(size=2000); lag = rng.uniform(size=2000); slope = rng.uniform(size=2000)
    w[:4] = [0.0, 1.0, 0.5, 1.0 / 3.0]
    ic_plain = np.array([L.amref_interpolate_color(m.h, int(a), int(b), 0.0, 0.0, float(x), 1) for a, b, x in zip(cols[:2000], cols[1000:3000], w)], dtype=np.uint32)
    ic_eased = np.array([L.amref_interpolate_color(m.h, int(a), int(b), float(l), float(s), float(x), 0)
                         for a, b, l, s, x in zip(cols[:2000], cols[1000:3000], lag, slope, w)], dtype=np.uint32)
    ip = np.array([L.amref_interpolate_point(m.h, int(a), int(b), float(x)) for a, b, x in zip(pts[:1999], pts[1:2000], w)], dtype=np.uint64)
    save("pure", cols=cols, hsp=hsp, back=back, cd=cd, pts=pts, pd=pd, seeds=seeds, nxy=nxy, noise=noise, sp_n=sp_n, sp_ctrl=sp_ctrl,
         sp_t=sp_t, sp_out=sp_out, w=w, lag=lag, slope=slope, ic_plain=ic_plain, ic_eased=ic_eased, ip=ip)


RENDER_CASES = [
    ("render_spline_cosine", lambda: scenes.ellipses(40, 2, seed=8), dict(motion=amref.SPLINE, fading=amref.COSINE, density=2), 0),
    ("render_linear_perlin_k3", lambda: scenes.ellipses(36, 3, seed=10, alpha_noise=True), dict(motion=amref.LINEAR, fading=amref.PERLIN, feather=2), 0),
    ("render_blobs_bg", lambda: scenes.random_cloud(32, 3, seed=3, margin=4), dict(motion=amref.SPLINE, fading=amref.COSINE, keep_background=1, density=2, feather=1), 150),
    ("render_blobs_average", lambda: scenes.rect_blobs(48, 10, frames=2, seed=5, min_side=4, max_side=12), dict(motion=amref.LINEAR, fading=amref.LINEAR, show_blobs=amref.AVERAGE), 100),
]
TIMES = np.array([0.0, 0.13, 1.0 / 3.0, 0.5, 0.77, 0.999])


def render_case(name, scene, params, match_steps):
    images = scene()
    H, W = images[0].shape[:2]
    m = amref.RefMorph(seed=3, **params)
    for k, im in enumerate(images):
        m.add_image(k, im)
    m.set_resolution(W, H)
    m.run_until(amref.STATE_ATOM_MORPHING, match_steps=match_steps)
    m.set(cycle_length=300)
    m.sync()
    m.iterate(30)
    m.sync()
    out = dict(images=np.stack(images), times=TIMES, bbox=np.array(m.bbox()), keys=np.array(m.frame_keys(), dtype=np.uint64),
               frames=np.stack([m.render(t) for t in TIMES]), cost=np.array(m.true_cost()))
    for k, v in dict(seed=3, **params).items():
        out["param_" + k] = np.array(v)
    for i, key in enumerate(m.frame_keys()):
        out["fetch_%d" % i] = m.fetch_image(key)
        out["labels_%d" % i] = m.blob_labels(key)
        bl = m.blobs(key)
        out["bstats_%d" % i] = np.array([b["stats"] for b in bl]).reshape(-1, 6)
        out["bgroup_%d" % i] = np.array([b["group"] for b in bl], dtype=np.uint64)
    ch = m.chains()
    out["chain_keys"] = np.array([c["key"] for c in ch], dtype=np.uint64)
    out["chain_ms"] = np.array([c["max_surface"] for c in ch], dtype=np.uint64)
    for i, c in enumerate(ch):
        out["chain_%d" % i] = c["words"]
    save(name, **out)


def swap():
    images = scenes.ellipses(40, 3, seed=33)
    m = amref.RefMorph(seed=5)
    for k, im in enumerate(images):
        m.add_image(k, im)
    m.set_resolution(40, 40)
    m.run_until(amref.STATE_ATOM_MORPHING)
    before = m.chains()[0]["words"].copy()
    e1 = m.e1_state()
    m.set(cycle_length=777)
    m.sync()
    m.iterate(40)
    m.sync()
    after = m.chains()[0]["words"].copy()
    save("swap", before=before, after=after, e1=np.array(e1, dtype=np.uint64), steps=np.array(40), cycle_length=np.array(777),
         cost_before=np.array(amref_cost(before)), cost_after=np.array(m.true_cost()), e1_after=np.array(m.e1_state(), dtype=np.uint64))


def amref_cost(words):
    h, w = words.shape
    e = 0.0
    for x in range(w):
        for j in range(h):
            e += float(amref.point_distance(words[j, x], words[(j + 1) % h, x]))
    return e


def blobs():
    for name, images in (("blobs_rects", scenes.rect_blobs(64, 40, frames=2, seed=3, min_side=2, max_side=9)),
                         ("blobs_cloud", scenes.random_cloud(48, 2, fill=0.62, seed=5, margin=1))):
        H, W = images[0].shape[:2]
        m = amref.RefMorph(seed=1, blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3)
        for k, im in enumerate(images):
            m.add_image(k, im)
        m.set_resolution(W, H)
        m.run_until(amref.STATE_BLOB_MATCHING)
        m.iterate(500)                                             # some accepted swaps so the groups are not the identity
        m.sync()
        out = dict(images=np.stack(images), bbox=np.array(m.bbox()))
        for i, key in enumerate(m.frame_keys()):
            st, pres = m.stored_image(key)
            out["stored_%d" % i] = st
            out["present_%d" % i] = pres
            out["labels_%d" % i] = m.blob_labels(key)
            bl = m.blobs(key)
            out["bstats_%d" % i] = np.array([b["stats"] for b in bl]).reshape(-1, 6)
            out["bsize_%d" % i] = np.array([len(b["surface"]) for b in bl], dtype=np.uint64)
            out["bfirst_%d" % i] = np.array([int(b["surface"][0]) if len(b["surface"]) else 2 ** 63 for b in bl], dtype=np.uint64)
            out["bgroup_%d" % i] = np.array([b["group"] for b in bl], dtype=np.uint64)
        wv = m.worker_values()
        out["energy_best"] = np.array(wv["best_blob_map_e"])      # energy of the map the blob groups describe
        out["weights"] = np.array([wv["w_xy"], wv["w_rgba"], wv["w_size"], wv["bbox_d"]])
        save(name, **out)


def fluid():
    rng = np.random.default_rng(7)
    n, gx, gy = 1500, 70, 60
    rec = np.zeros((n, amref.FP_STRIDE))
    rec[:, 0] = rng.uniform(12, gx - 12, n); rec[:, 1] = rng.uniform(12, gy - 12, n)
    rec[:, 2] = rng.normal(0, 0.2, n); rec[:, 3] = rng.normal(0, 0.2, n)
    rec[:, 4] = np.clip(rec[:, 0] + rng.normal(0, 3, n), 1, gx - 2); rec[:, 5] = np.clip(rec[:, 1] + rng.normal(0, 3, n), 1, gy - 2)
    rec[:, 6] = 1.0
    rec[:, 7] = rng.uniform(size=n) > 0.1; rec[:, 8] = rng.uniform(size=n) > 0.2
    rec[:, 9:17] = rng.uniform(size=(n, 8))
    rec[:, 17] = rng.choice([1.0, 0.1], size=n, p=[0.8, 0.2])
    f = amref.RefFluid(gx, gy, n)
    f.set_particles(rec)
    steps = [(3, 12.0), (0, 1.5)]
    outs, nodes = [], []
    for sl, rad in steps:
        f.step(sl, rad, 0.3)
        outs.append(f.get_particles())
        nodes.append(f.nodes())
    save("fluid", rec=rec, dims=np.array([gx, gy, n]), steps=np.array(steps), after=np.stack(outs), nodes=np.stack(nodes))


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "pure" in only:
        pure()
    if not only or "render" in only:
        for case in RENDER_CASES:
            render_case(*case)
    if not only or "swap" in only:
        swap()
    if not only or "blobs" in only:
        blobs()
    if not only or "fluid" in only:
        fluid()
