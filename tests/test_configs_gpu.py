"""BASELINE.json configs 2-5 at (close to) full size through the device pipeline, checked with size-independent
properties (the reference needs minutes to hours per frame at these sizes):

* exact integer bookkeeping of the matcher: cost(before) - cost(after) == the accumulated gain
* a frame rendered at the key time of the LARGEST key frame (no duplicate atoms in that column) is that key frame
* rendering a batch of frames equals rendering them one at a time; rendering is deterministic
* C4: the partition equals the 4-connected components (the deterministic regime of blobify)
* C3: the particle population follows round((1-t)|before| + t|after|)
The small-size, bit-for-bit comparisons against the reference live in test_render_gpu / test_swap_gpu / test_blobs_gpu /
test_fluid_gpu / test_golden_gpu."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes

pytestmark = pytest.mark.gpu


def _pipeline(images, **params):
    e = eng.Engine(0, threads=0, cycle_length=1000, **params)
    e.load_images(images)
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING, e.state()
    return e


def test_c2_square_to_disc_1024():
    images = scenes.square_to_disc(1024)
    e = _pipeline(images, seed=1, motion=eng.SPLINE, fading=eng.COSINE)
    assert e.chain_count() == 1 and e.table_device_ptr(0)[1] == 1024 * 1024
    c0, st0 = e.cost(), e.swap_stats()
    e.swap_rounds(2048)
    c1, st1 = e.cost(), e.swap_stats()
    assert int(st1[0] - st0[0]) == 2048 * 1024 * 1024 // 2
    assert c0 - c1 == float(int(st1[2] - st0[2])) and c1 < 0.05 * c0
    # t = 0 is the key time of the full square (every atom sits on its own pixel, no duplicates in column 0)
    f0 = e.render([0.0])[0]
    assert np.array_equal(f0, e.fetch_image(0))
    ts = [f / 64.0 for f in (0, 1, 17, 31, 32, 33, 63)]
    batch = e.render(ts)
    for i, t in enumerate(ts):
        assert np.array_equal(batch[i], e.render([t])[0])
    # the disc's key time: the disc pixels are all covered
    f1 = eng.unpack_rgba(e.render([0.5])[0])
    assert ((f1[..., 3] > 0) >= (images[1][..., 3] > 0)).all()


def test_c3_fluid_perlin_feather_1024():
    images = scenes.square_to_disc(1024)
    e = _pipeline(images, seed=1, motion=eng.SPLINE, fading=eng.PERLIN, feather=2, fluid=10)
    e.swap_rounds(512, want_stats=False)
    n0, n1 = 1024 * 1024, int((images[1][..., 3] != 0).sum())
    for t in (0.0, 0.05):
        img = eng.unpack_rgba(e.render([t])[0])
        rec = e.fluid_get_particles(1024 * 1024)
        tl = 2.0 * t
        assert int((rec[:, 7] != 0).sum()) == int(round((1.0 - tl) * n0 + tl * n1))
        assert (img[..., 3] > 0).mean() > 0.75
        act = rec[:, 7] != 0
        assert np.isfinite(rec[act][:, :4]).all()
        assert (rec[act, 0] >= 1.0).all() and (rec[act, 0] <= 1024 + 18).all()


def test_c4_blob_heavy_512():
    from scipy import ndimage
    images = scenes.rect_blobs(512, 2500, frames=2, seed=11, min_side=2, max_side=20)
    e = eng.Engine(0, seed=1, motion=eng.LINEAR, fading=eng.COSINE, density=2, blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3,
                   threads=0, cycle_length=1000)
    e.load_images(images)
    e.blobify()
    for i in range(2):
        labels, stats, meta = e.export_blobs(i)
        ref_labels, nref = ndimage.label(images[i][..., 3] != 0)
        assert int((meta[:, 1] > 0).sum()) == nref
        # same partition: one engine label per scipy component and vice versa
        pairs = np.unique(np.stack([labels[labels >= 0], ref_labels[labels >= 0]]), axis=1)
        assert pairs.shape[1] == nref
    e.match_init()
    e0 = e.match_energy()
    e.match_rounds(2000)
    assert e.match_energy() < 0.6 * e0
    e.init_chains()
    assert e.chain_count() >= 2500
    c0, st0 = e.cost(), e.swap_stats()
    e.swap_rounds(400)                       # all chains per round (k_swap_multi)
    c1, st1 = e.cost(), e.swap_stats()
    assert c0 - c1 == float(int(st1[2] - st0[2])) and c1 < c0
    ts = [f / 128.0 for f in (0, 5, 64, 100)]
    batch = e.render(ts)
    for i, t in enumerate(ts):
        assert np.array_equal(batch[i], e.render([t])[0])
    assert (eng.unpack_rgba(batch[1])[..., 3] > 0).mean() > 0.3


@pytest.mark.parametrize("n", [1024, 4096])
def test_c5_eight_cyclic_key_frames(n):
    images = scenes.rotating_shapes(n, 8)
    e = _pipeline(images, seed=1, motion=eng.SPLINE, fading=eng.COSINE)
    A = e.table_device_ptr(0)[1]
    sizes = [int((im[..., 3] != 0).sum()) for im in images]
    assert A == max(sizes)
    c0, st0 = e.cost(), e.swap_stats()
    e.swap_rounds(256)
    c1, st1 = e.cost(), e.swap_stats()
    assert c0 - c1 == float(int(st1[2] - st0[2])) and c1 < 0.7 * c0
    k = int(np.argmax(sizes))                # the key frame without duplicate atoms
    fk = e.render([k / 8.0])[0]
    assert np.array_equal(fk, e.fetch_image(k))
    ts = [0.01, 0.124, 0.126, 0.5, 0.99]     # incl. both sides of a key-frame boundary (batches never span one)
    batch = e.render(ts)
    for i, t in enumerate(ts):
        assert np.array_equal(batch[i], e.render([t])[0])
