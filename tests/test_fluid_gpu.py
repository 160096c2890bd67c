"""K5 parity: one FluidModel::step from identical state, fields within 1e-5 relative
(SURVEY.md row a-F, hard part 5)."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from helpers import build_ref, diff_stats, engine_from_ref

pytestmark = pytest.mark.gpu


def _particles(n, gx, gy, seed, immature_frac=0.2, inactive_frac=0.1):
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, eng.FP_STRIDE))
    rec[:, 0] = rng.uniform(12, gx - 12, n)
    rec[:, 1] = rng.uniform(12, gy - 12, n)
    rec[:, 2] = rng.normal(0, 0.2, n)
    rec[:, 3] = rng.normal(0, 0.2, n)
    rec[:, 4] = np.clip(rec[:, 0] + rng.normal(0, 3, n), 1, gx - 2)
    rec[:, 5] = np.clip(rec[:, 1] + rng.normal(0, 3, n), 1, gy - 2)
    rec[:, 6] = 1.0
    rec[:, 7] = rng.uniform(size=n) > inactive_frac
    rec[:, 8] = rng.uniform(size=n) > immature_frac
    rec[:, 9:13] = rng.uniform(size=(n, 4))
    rec[:, 13:17] = rng.uniform(size=(n, 4))
    rec[:, 17] = rng.choice([1.0, 0.1], size=n, p=[0.8, 0.2])
    return rec


@pytest.mark.parametrize("n,gx,gy,seed", [(500, 52, 52, 1), (4000, 84, 70, 2)])
def test_single_step_matches_reference(reflib, n, gx, gy, seed):
    rec = _particles(n, gx, gy, seed)
    rf = reflib.RefFluid(gx, gy, n)
    rf.set_particles(rec)
    e = eng.Engine(0)
    e.fluid_create(gx, gy, n)
    e.fluid_set_particles(rec)
    for step, (steps_left, radius) in enumerate([(3, 12.0), (2, 8.0), (0, 1.5)]):
        rf.step(steps_left, radius, 0.3)
        e.fluid_step(steps_left, radius, 0.3)
        a, b = rf.get_particles(), e.fluid_get_particles()
        act = a[:, 7] != 0
        for col in (0, 1, 2, 3, 13, 14, 15, 16):
            ref, got = a[act, col], b[act, col]
            scale = np.maximum(np.abs(ref), 1e-3 if col in (2, 3) else 1.0)
            assert np.max(np.abs(ref - got) / scale) < 1e-5, (step, col)
        # untouched fields
        assert np.array_equal(a[~act, :2], b[~act, :2])
        na, nb = rf.nodes(), e.fluid_nodes()
        for k in range(13):
            scale = max(np.abs(na[..., k]).max(), 1e-9)
            assert np.max(np.abs(na[..., k] - nb[..., k])) / scale < 1e-5, (step, "node", k)
    rf.close()


# ---- the fluid render path: morph::step_fluid / update_particle / draw_fluid (morph.cpp:680-1300)
FLUID_CASES = [
    ("linear", dict(motion=eng.LINEAR, fading=eng.LINEAR, fluid=3)),
    ("spline_cosine_feather", dict(motion=eng.SPLINE, fading=eng.COSINE, fluid=2, feather=2)),
    ("perlin_bg", dict(motion=eng.LINEAR, fading=eng.PERLIN, fluid=2, keep_background=1)),
]


@pytest.mark.parametrize("name,params", FLUID_CASES, ids=[c[0] for c in FLUID_CASES])
def test_fluid_first_frame_matches_reference(reflib, name, params):
    """t = 0 of an interval: every source gets exactly one particle, placed on its attractor, so the frame is the
    reference's up to the order of floating-point sums (which particle slot holds which source is the reference's RNG)."""
    images = scenes.ellipses(40, 2, seed=61)
    m = build_ref(reflib, images, seed=1, **params)
    m.sync()
    m.fluid_sanitize()
    e = engine_from_ref(m, images, seed=1, **params)
    ref = m.render(0.0)
    got = e.render([0.0])[0]
    ndiff, mx = diff_stats(ref, got)
    assert mx <= 2, (name, mx)
    assert ndiff <= 0.02 * ref.size, (name, ndiff)
    # every source carries one active particle
    W = e.chains()[0]["width"]
    rec = e.fluid_get_particles(W)
    assert int((rec[:, 7] != 0).sum()) == int((images[0][..., 3] != 0).sum())


def test_fluid_sequence_follows_reference(reflib):
    """Across an interval the particle population follows round((1-t)|before| + t|after|) exactly; the images agree
    with the reference statistically (the reference draws its creations / deletions from its host RNG)."""
    images = scenes.ellipses(40, 2, seed=62)
    params = dict(motion=eng.LINEAR, fading=eng.LINEAR, fluid=3)
    m = build_ref(reflib, images, seed=1, **params)
    m.sync()
    m.fluid_sanitize()
    e = engine_from_ref(m, images, seed=1, **params)
    n0, n1 = int((images[0][..., 3] != 0).sum()), int((images[1][..., 3] != 0).sum())
    W = e.chains()[0]["width"]
    for t in (0.0, 0.1, 0.2, 0.3, 0.45):
        ref = eng.unpack_rgba(m.render(t)).astype(np.float64)
        got = eng.unpack_rgba(e.render([t])[0]).astype(np.float64)
        tl = t * 2.0                                    # two key frames: local t of interval 0
        rec = e.fluid_get_particles(W)
        assert int((rec[:, 7] != 0).sum()) == int(round((1.0 - tl) * n0 + tl * n1)), t
        cover_ref, cover_got = (ref[..., 3] > 0), (got[..., 3] > 0)
        assert abs(int(cover_ref.sum()) - int(cover_got.sum())) <= 0.10 * cover_ref.sum(), t
        both = cover_ref & cover_got
        assert both.sum() >= 0.85 * cover_ref.sum(), t
        assert np.abs(ref[both] - got[both]).mean() < 6.0, t
    # a new interval invalidates the particles (morph.cpp:868-878) and repopulates from the second key frame
    e.render([0.5])
    rec = e.fluid_get_particles(W)
    assert int((rec[:, 7] != 0).sum()) == n1
