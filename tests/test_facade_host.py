"""Host logic of the am::morph facade that needs no GPU: ingest bookkeeping, time mapping,
interpolation helpers, colour round trip -- against the reference built from source."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from atomorph_b200.morph import Morph


def _both(reflib, images, **params):
    m = Morph()
    r = reflib.RefMorph(**params)
    for k, v in params.items():
        getattr(m, "set_" + k)(v)
    for k, im in enumerate(images):
        m.add_image(k, im)
        r.add_image(k, im)
    H, W = images[0].shape[:2]
    m.set_resolution(W, H)
    r.set_resolution(W, H)
    return m, r


@pytest.mark.parametrize("delimiter", [eng.HSP, eng.RGB])
def test_ingest_matches_reference(reflib, delimiter):
    images = scenes.random_cloud(24, 3, seed=5, margin=1)
    m, r = _both(reflib, images, blob_delimiter=delimiter, seed=3)
    assert m.get_frame_count() == 3
    for k in range(3):
        assert m.get_pixel_count(k) == r.L.amref_pixel_count(r.h, k)
        assert m.get_average_pixel(k) == r.average_pixel(k)
        for pos in range(0, 24 * 65536, 65536 + 1):
            assert m.get_pixel(k, pos) == r.L.amref_get_pixel(r.h, k, pos)
    assert m.get_pixel(7, 0) == 0 and m.get_pixel_count(7) == 0


def test_time_mapping_and_interpolation(reflib):
    images = scenes.ellipses(16, 3, seed=2)
    for finite in (0, 1):
        m, r = _both(reflib, images, finite=finite, fading=eng.COSINE, seed=9)
        for total in (1, 2, 7, 64):
            for f in range(total):
                assert m.get_time(f, total) == r.get_time(f, total)
        for t in (0.0, 0.1, 1 / 3, 0.5, 0.999, 1.25, -0.2):
            assert m.get_frame_key(t) == r.L.amref_get_frame_key(r.h, t)
        rng = np.random.default_rng(0)
        for _ in range(200):
            c1, c2 = (int(v) for v in rng.integers(0, 2 ** 32, 2))
            w, lag, slope = rng.uniform(size=3)
            assert m.interpolate_color(c1, c2, w) == r.L.amref_interpolate_color(r.h, c1, c2, 0, 0, w, 1)
            assert m.interpolate_color(c1, c2, w, lag, slope) == r.L.amref_interpolate_color(r.h, c1, c2, lag, slope, w, 0)
            p1, p2 = (int(v) & 0xffffffffffff for v in rng.integers(0, 2 ** 62, 2))
            assert m.interpolate_point(p1, p2, w) == r.L.amref_interpolate_point(r.h, p1, p2, w)


@pytest.mark.parametrize("fading", [eng.LINEAR, eng.COSINE, eng.PERLIN])
def test_get_background(reflib, fading):
    images = scenes.random_cloud(20, 3, seed=8, margin=1)
    m, r = _both(reflib, images, fading=fading, seed=4)
    for t in (0.0, 0.2, 0.5, 0.9):
        for y in range(0, 20, 3):
            for x in range(0, 20, 3):
                assert m.get_background(x, y, t) == r.get_background(x, y, t)
