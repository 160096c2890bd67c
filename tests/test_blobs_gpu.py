"""K2/K3/K4 parity: blob partition (bit-exact as a partition in the deterministic regime), blob
statistics, matching energy, chain construction (SURVEY.md rows a-B1, a-B3, a-C)."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from helpers import build_ref

pytestmark = pytest.mark.gpu


def _canonical(labels):
    """partition -> image of the smallest row-major index of each blob (-1 where absent)."""
    out = np.full(labels.shape, -1, dtype=np.int64)
    flat = labels.reshape(-1)
    idx = np.arange(flat.size)
    ok = flat >= 0
    if ok.any():
        mins = np.full(flat.max() + 1, flat.size, dtype=np.int64)
        np.minimum.at(mins, flat[ok], idx[ok])
        out.reshape(-1)[ok] = mins[flat[ok]]
    return out


@pytest.mark.parametrize("scene", ["rects", "cloud", "ellipses"])
def test_partition_and_stats(reflib, scene):
    if scene == "rects":
        images = scenes.rect_blobs(96, 60, frames=2, seed=3, min_side=2, max_side=10)
    elif scene == "cloud":
        images = scenes.random_cloud(64, 2, fill=0.62, seed=5, margin=2)
    else:
        images = scenes.ellipses(48, 3, seed=2)
    m = build_ref(reflib, images, seed=1, target=reflib.STATE_BLOB_MATCHING)
    e = eng.Engine(0, seed=1)
    e.load_images(images)
    e.blobify()
    for i, key in enumerate(m.frame_keys()):
        ref_labels = m.blob_labels(key)
        labels, stats, meta = e.export_blobs(i)
        nreal = int((meta[:, 1] > 0).sum())
        ref_blobs = [b for b in m.blobs(key) if len(b["surface"])]
        assert nreal == len(ref_blobs)
        assert np.array_equal(_canonical(ref_labels), _canonical(labels)), "partition differs"
        # statistics, matched through the canonical label
        can = _canonical(labels)
        for rb in ref_blobs:
            p = int(rb["surface"][0])
            x, y = p % 65536, p // 65536
            b = labels[y, x]
            assert int(meta[b, 1]) == len(rb["surface"])
            assert np.allclose(stats[b], rb["stats"], rtol=0, atol=1e-9)
        # fetch / stored colours (K7)
        assert np.array_equal(e.fetch_image(i), m.fetch_image(key))
        st, pres = m.stored_image(key)
        assert np.array_equal(e.stored_image(i)[pres], st[pres])


def test_matching_energy_not_worse_than_reference(reflib):
    images = scenes.rect_blobs(128, 120, frames=2, seed=9, min_side=2, max_side=9)
    params = dict(blob_rgba_weight=2, blob_size_weight=1, blob_xy_weight=3, seed=4)
    nb = 120
    m = build_ref(reflib, images, target=reflib.STATE_BLOB_MATCHING, **params)
    m.iterate(1000 * nb)
    m.sync()
    e_ref = m.best_blob_energy()
    e = eng.Engine(0, **params)
    e.load_images(images)
    e.blobify()
    e.match_init()
    e0 = e.match_energy()
    e.match_rounds(4000)
    e1 = e.match_energy()
    assert e1 <= e0
    assert e1 <= 1.01 * e_ref, (e1, e_ref)


def test_chain_build_matches_reference_without_duplicates(reflib):
    # equal-size single blobs, density 1 -> no duplicates -> the table is deterministic
    images = scenes.ellipses(40, 2, seed=2)
    images[1] = np.roll(images[0], 3, axis=1)       # same pixel count in both frames
    m = build_ref(reflib, images, seed=1)
    ref = m.chains()
    e = eng.Engine(0, seed=1, threads=0, cycle_length=0)
    e.load_images(images)
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    got = e.chains()
    assert len(got) == len(ref) == 1
    assert np.array_equal(got[0]["words"], ref[0]["words"])
    assert got[0]["max_surface"] == ref[0]["max_surface"]


def test_chain_build_with_duplicates_and_volatile(reflib):
    images = scenes.ellipses(40, 3, seed=3)
    images[2][...] = 0                                # empty key frame -> volatile blob
    m = build_ref(reflib, images, seed=1, density=2)
    ref = m.chains()[0]
    e = eng.Engine(0, seed=1, density=2, threads=0, cycle_length=0)
    e.load_images(images)
    e.step(8)
    got = e.chains()[0]
    assert got["words"].shape == ref["words"].shape
    assert got["max_surface"] == ref["max_surface"]
    for j in range(ref["height"]):
        r, g = ref["words"][j], got["words"][j]
        # originals (HAS_FLUID set) are identical and in the same order
        ro, go = r[(r >> 48) & 2 != 0], g[(g >> 48) & 2 != 0]
        assert np.array_equal(ro, go)
        # duplicates: same multiset of source pixels (x, y), flags without HAS_FLUID
        rd, gd = r[(r >> 48) & 2 == 0], g[(g >> 48) & 2 == 0]
        assert len(rd) == len(gd)
        assert np.array_equal(np.unique((rd >> 48) & 0xff), np.unique((gd >> 48) & 0xff))
        cr = np.unique(rd & 0xffffffff, return_counts=True)
        cg = np.unique(gd & 0xffffffff, return_counts=True)
        assert np.array_equal(cr[0], cg[0])
        assert np.abs(cr[1].astype(int) - cg[1].astype(int)).max() <= 1
