"""The CUDA path against the committed golden vectors (generated from the unmodified reference) and
against the CPU oracle on the same seeded inputs.  Needs only the GPU box: no reference tree."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from oracle import amoracle
import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gu.RENDER_CASES)
def test_render_golden_bit_exact(case):
    g = gu.load(case)
    e = gu.engine_from_golden(eng, g)
    assert e.cost() == float(g["cost"])
    got = e.render(g["times"])
    for i, t in enumerate(g["times"]):
        assert np.array_equal(got[i], g["frames"][i]), "%s t=%g: %d pixels differ" % (case, t, int((got[i] != g["frames"][i]).sum()))


@pytest.mark.parametrize("case", ["blobs_rects", "blobs_cloud"])
def test_blobs_golden(case):
    g = gu.load(case)
    images = [im for im in g["images"]]
    e = eng.Engine(0, seed=1)
    e.load_images(images)
    e.blobify()
    for i in range(len(images)):
        labels, stats, meta = e.export_blobs(i)
        assert np.array_equal(gu.canonical(labels), gu.canonical(g["labels_%d" % i]))
        real = g["bsize_%d" % i] > 0
        order = np.argsort(g["bfirst_%d" % i][real])
        assert np.array_equal(meta[:, 1], g["bsize_%d" % i][real][order])          # our blob order = ascending canonical label
        assert np.allclose(stats, g["bstats_%d" % i][real][order], rtol=0, atol=1e-9)
        pres = g["present_%d" % i]
        assert np.array_equal(e.stored_image(i)[pres], g["stored_%d" % i][pres])


def test_fluid_golden():
    g = gu.load("fluid")
    gx, gy, n = (int(v) for v in g["dims"])
    e = eng.Engine(0)
    e.fluid_create(gx, gy, n)
    e.fluid_set_particles(g["rec"])
    for (sl, rad), ref_rec, ref_nodes in zip(g["steps"], g["after"], g["nodes"]):
        e.fluid_step(int(sl), float(rad), 0.3)
        got = e.fluid_get_particles()
        act = ref_rec[:, 7] != 0
        for col in (0, 1, 2, 3, 13, 14, 15, 16):
            scale = np.maximum(np.abs(ref_rec[act, col]), 1e-3 if col in (2, 3) else 1.0)
            assert np.max(np.abs(ref_rec[act, col] - got[act, col]) / scale) < 1e-5     # north star: fluid fields within 1e-5 relative
        nodes = e.fluid_nodes()
        for k in range(13):
            scale = max(np.abs(ref_nodes[..., k]).max(), 1e-9)
            assert np.max(np.abs(ref_nodes[..., k] - nodes[..., k])) / scale < 1e-5


def test_swap_golden_cost():
    """Same table, same proposal budget as the reference's serial run: the parallel matcher must end no
    worse than 1% above it (the serial RNG stream itself cannot be reproduced in parallel)."""
    g = gu.load("swap")
    before = g["before"]
    W = before.shape[1]
    images = scenes.ellipses(40, 3, seed=33)
    e = eng.Engine(0, seed=5)
    e.load_images(images)
    e.import_chains([dict(key=0, words=before, max_surface=W)])
    assert e.cost() == float(g["cost_before"])
    budget = int(g["steps"]) * int(g["cycle_length"])
    st0 = e.swap_stats()
    while int(e.swap_stats()[0] - st0[0]) < budget:
        e.swap_rounds(8)
    assert e.cost() <= 1.01 * float(g["cost_after"])
    after = e.chains()[0]["words"]
    for j in range(before.shape[0]):
        assert np.array_equal(np.sort(before[j]), np.sort(after[j]))


def test_render_matches_oracle_large():
    """Full-size property check (BASELINE config 2 geometry at 512^2): CUDA == oracle on sampled rows,
    after real matching rounds, spline + cosine."""
    n = 192
    images = scenes.square_to_disc(n)
    params = dict(seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=0)
    e = eng.Engine(0, **params)
    e.load_images(images)
    e.step(8)
    e.swap_rounds(300)
    chains = e.chains()
    blobs = []
    for i in range(2):
        labels, stats, meta = e.export_blobs(i)
        blobs.append([dict(group=int(meta[b, 0]), stats=stats[b]) for b in range(len(meta))])
    fetch = [e.fetch_image(i) for i in range(2)]
    has = [im[..., 3] != 0 for im in images]
    S = amoracle.RenderScene(n, n, e.bbox, [0, 1], fetch, has, blobs, chains, motion=eng.SPLINE, fading=eng.COSINE, seed=1)
    assert amoracle.cost(chains) == e.cost()
    for t in (0.0, 0.37, 0.5):
        assert np.array_equal(e.render([t])[0], S.render(t))
