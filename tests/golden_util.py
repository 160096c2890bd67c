"""Loaders for the golden vectors of tests/golden/ (generated from the reference by make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RENDER_CASES = ["render_spline_cosine", "render_linear_perlin_k3", "render_blobs_bg", "render_blobs_average"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def render_params(g):
    return {k[6:]: int(g[k]) for k in g.files if k.startswith("param_")}


def render_tables(g):
    """-> images, keys, bbox, per-frame (fetch, labels, stats, groups), chains."""
    images = [im for im in g["images"]]
    n = len(images)
    frames = []
    for i in range(n):
        frames.append(dict(fetch=g["fetch_%d" % i], labels=g["labels_%d" % i], stats=g["bstats_%d" % i], groups=g["bgroup_%d" % i]))
    chains = [dict(key=int(k), words=g["chain_%d" % i], max_surface=int(ms)) for i, (k, ms) in enumerate(zip(g["chain_keys"], g["chain_ms"]))]
    return images, [int(k) for k in g["keys"]], [int(v) for v in g["bbox"]], frames, chains


def oracle_scene(amoracle, g):
    images, keys, bbox, frames, chains = render_tables(g)
    p = render_params(g)
    H, W = images[0].shape[:2]
    has = [im[..., 3] != 0 for im in images]
    blobs = [[dict(group=int(gr), stats=st) for gr, st in zip(f["groups"], f["stats"])] for f in frames]
    return amoracle.RenderScene(W, H, bbox, keys, [f["fetch"] for f in frames], has, blobs, chains,
                                motion=p.get("motion", 4), fading=p.get("fading", 6), density=p.get("density", 1),
                                feather=p.get("feather", 0), show_blobs=p.get("show_blobs", 0),
                                keep_background=p.get("keep_background", 0), blob_delimiter=p.get("blob_delimiter", 1), seed=p.get("seed", 0))


def engine_from_golden(eng, g, device=0):
    images, keys, bbox, frames, chains = render_tables(g)
    p = render_params(g)
    e = eng.Engine(device, **p)
    e.load_images(images)
    for i, f in enumerate(frames):
        e.import_blobs(i, f["labels"], f["stats"], f["groups"])
    e.import_chains(chains)
    return e


def canonical(labels):
    out = np.full(labels.shape, -1, dtype=np.int64)
    flat = labels.reshape(-1)
    idx = np.arange(flat.size)
    ok = flat >= 0
    if ok.any():
        mins = np.full(int(flat.max()) + 1, flat.size, dtype=np.int64)
        np.minimum.at(mins, flat[ok], idx[ok])
        out.reshape(-1)[ok] = mins[flat[ok]]
    return out
