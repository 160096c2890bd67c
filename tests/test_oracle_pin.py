"""Pins the CPU oracle against the LIVE reference (oracle/_ref/libamref.so) on fresh random cases, beyond
the committed golden vectors.  Skipped where the reference build is absent."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from oracle import amoracle
from helpers import build_ref

pytestmark = pytest.mark.skipif(not amoracle.available(), reason="oracle/libamoracle.so not built")


def _scene_from_ref(m, images, **params):
    H, W = images[0].shape[:2]
    keys = m.frame_keys()
    fetch = [m.fetch_image(k) for k in keys]
    has = [im[..., 3] != 0 for im in images]
    blobs = [[dict(group=b["group"], stats=b["stats"]) for b in m.blobs(k)] for k in keys]
    return amoracle.RenderScene(W, H, m.bbox(), keys, fetch, has, blobs, m.chains(), **params)


CASES = [
    (lambda: scenes.ellipses(36, 4, seed=41), dict(motion=eng.SPLINE, fading=eng.PERLIN, feather=1), 0),
    (lambda: scenes.random_cloud(28, 2, seed=13, margin=3), dict(motion=eng.LINEAR, fading=eng.COSINE, density=3, keep_background=1), 80),
    (lambda: scenes.rect_blobs(40, 8, frames=3, seed=17, min_side=3, max_side=10), dict(motion=eng.SPLINE, fading=eng.LINEAR, show_blobs=eng.DISTINCT, feather=2), 60),
    (lambda: scenes.ellipses(30, 2, seed=19), dict(motion=eng.NONE, fading=eng.NONE, blob_delimiter=eng.RGB), 0),
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_render_oracle_equals_reference(reflib, idx):
    scene, params, match_steps = CASES[idx]
    images = scene()
    m = build_ref(reflib, images, seed=7, match_steps=match_steps, **params)
    m.set(cycle_length=200)
    m.sync()
    m.iterate(25)
    m.sync()
    S = _scene_from_ref(m, images, seed=7, **params)
    for t in (0.0, 0.21, 0.5, 2.0 / 3.0, 0.93):
        assert np.array_equal(S.render(t), m.render(t))
    assert amoracle.cost(m.chains()) == m.true_cost()


def test_matcher_replay_long(reflib):
    images = scenes.ellipses(32, 2, seed=23)
    m = build_ref(reflib, images, seed=11)
    before = m.chains()[0]["words"].copy()
    e1 = m.e1_state()
    m.set(cycle_length=1234)
    m.sync()
    m.iterate(100)
    m.sync()
    words, e1b, gain = amoracle.morph_steps(before, 100, 1234, e1)
    assert np.array_equal(words, m.chains()[0]["words"])
    assert e1b == m.e1_state()


def test_pure_functions_exhaustive_slices(reflib):
    L = reflib.lib()
    for c in range(0, 2 ** 24, 4099):
        v = c | 0x7f000000
        assert amoracle.rgb_to_hsp(v) == L.amref_rgb_to_hsp(v)
        assert amoracle.hsp_to_rgb(v) == L.amref_hsp_to_rgb(v)
