"""K6 parity: rendered frames vs the reference run on the same inputs, fed the reference's own
correspondence / trajectory tables (SURVEY.md row a-R).  Bar: <= 1 LSB per 8-bit channel, and the
differing pixels (exact .5 ties of the double sums) stay a small fraction."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from helpers import build_ref, diff_stats, engine_from_ref

pytestmark = pytest.mark.gpu

TIMES = [0.0, 0.1, 0.25, 1.0 / 3.0, 0.5, 0.61, 0.75, 0.999]

CASES = [
    # name, scene, params
    ("linear_linear_k2", lambda: scenes.ellipses(48, 2, seed=7), dict(motion=eng.LINEAR, fading=eng.LINEAR)),
    ("spline_cosine_k2_d2", lambda: scenes.ellipses(48, 2, seed=8), dict(motion=eng.SPLINE, fading=eng.COSINE, density=2)),
    ("spline_cosine_k4", lambda: scenes.ellipses(48, 4, seed=9), dict(motion=eng.SPLINE, fading=eng.COSINE)),
    ("linear_perlin_k3", lambda: scenes.ellipses(40, 3, seed=10, alpha_noise=True), dict(motion=eng.LINEAR, fading=eng.PERLIN)),
    ("spline_perlin_feather", lambda: scenes.ellipses(40, 3, seed=11), dict(motion=eng.SPLINE, fading=eng.PERLIN, feather=3)),
    ("none_none_rgb", lambda: scenes.ellipses(32, 2, seed=12), dict(motion=eng.NONE, fading=eng.NONE, blob_delimiter=eng.RGB)),
    ("cloud_bg", lambda: scenes.random_cloud(40, 3, seed=3), dict(motion=eng.LINEAR, fading=eng.COSINE, keep_background=1, feather=2, density=2)),
    ("cloud_bg_perlin", lambda: scenes.random_cloud(32, 2, seed=4), dict(motion=eng.SPLINE, fading=eng.PERLIN, keep_background=1)),
]


@pytest.mark.parametrize("name,scene,params", CASES, ids=[c[0] for c in CASES])
def test_render_matches_reference(reflib, name, scene, params):
    images = scene()
    m = build_ref(reflib, images, seed=1, **params)
    m.set(cycle_length=500)
    m.sync()
    m.iterate(40)          # some matching so the table is not the trivial initial one
    m.sync()
    e = engine_from_ref(m, images, seed=1, **params)
    assert abs(e.cost() - m.true_cost()) == 0
    total = ndiff = 0
    for t in TIMES:
        ref = m.render(t)
        got = e.render([t])[0]
        n, mx = diff_stats(ref, got)
        assert mx <= 1, "%s t=%g: channel diff %d" % (name, t, mx)
        ndiff += n
        total += ref.size
    assert ndiff <= 0.01 * total, "%s: %d of %d pixels differ" % (name, ndiff, total)
    # a single chain without feather goes through the tiled path (shared-memory tiles), everything else through the general A-buffer path
    paths = e.render_path_frames()
    if params.get("feather", 0) == 0 and e.chain_count() == 1:
        assert paths["tiled"] == len(TIMES) and paths["general"] == 0, paths
    else:
        assert paths["tiled"] == 0, paths


def test_render_multi_blob_order_and_show_blobs(reflib):
    images = scenes.rect_blobs(64, 12, frames=2, seed=5, min_side=4, max_side=14)
    for show in (eng.TEXTURE, eng.AVERAGE, eng.DISTINCT):
        params = dict(motion=eng.LINEAR, fading=eng.LINEAR, show_blobs=show, feather=1)
        m = build_ref(reflib, images, seed=2, match_steps=200, **params)
        e = engine_from_ref(m, images, seed=2, **params)
        for t in (0.0, 0.3, 0.5, 0.8):
            n, mx = diff_stats(m.render(t), e.render([t])[0])
            assert mx <= 1
            assert n <= 0.01 * 64 * 64


def test_batch_equals_single(reflib):
    images = scenes.ellipses(32, 2, seed=21)
    params = dict(motion=eng.SPLINE, fading=eng.COSINE)
    m = build_ref(reflib, images, seed=1, **params)
    e = engine_from_ref(m, images, seed=1, **params)
    ts = [m.get_time(f, 16) for f in range(16)]
    batch = e.render(ts)
    for i, t in enumerate(ts):
        assert np.array_equal(batch[i], e.render([t])[0])


def test_tiled_path_equals_general_path(reflib, monkeypatch):
    """The two render paths (tiled / general A-buffer) are both exact: identical frames, 2x2 and 3x3 tiles, partial edge tiles."""
    for size, k, params in ((64, 2, dict(motion=eng.SPLINE, fading=eng.COSINE)), (80, 3, dict(motion=eng.LINEAR, fading=eng.PERLIN, density=2)),
                            (72, 2, dict(motion=eng.SPLINE, fading=eng.LINEAR, keep_background=1))):
        images = scenes.ellipses(size, k, seed=31 + size)
        m = build_ref(reflib, images, seed=3, **params)
        m.set(cycle_length=2000)
        m.sync()
        m.iterate(30)
        m.sync()
        ts = [m.get_time(f, 12) for f in range(12)]
        monkeypatch.setenv("AMX_RENDER_TILED", "1")
        e1 = engine_from_ref(m, images, seed=3, **params)
        a = e1.render(ts)
        assert e1.render_path_frames() == dict(tiled=12, general=0)
        monkeypatch.setenv("AMX_RENDER_TILED", "0")
        e0 = engine_from_ref(m, images, seed=3, **params)
        b = e0.render(ts)
        assert e0.render_path_frames() == dict(tiled=0, general=12)
        assert np.array_equal(a, b)
        for i in (0, 5, 11):
            n, mx = diff_stats(m.render(ts[i]), a[i])
            assert mx <= 1 and n <= 0.01 * a[i].size


def test_tiled_path_overflow_falls_back(reflib):
    """More atoms than a tile's bins take (density 8 = 8 atoms per pixel): the frames are rendered again by the general path."""
    images = scenes.ellipses(96, 2, seed=17)             # the middle tile is covered completely: 8192 records > 7168
    params = dict(motion=eng.LINEAR, fading=eng.LINEAR, density=8)
    m = build_ref(reflib, images, seed=1, **params)
    e = engine_from_ref(m, images, seed=1, **params)
    for t in (0.0, 0.4, 0.9):
        n, mx = diff_stats(m.render(t), e.render([t])[0])
        assert mx <= 1 and n <= 0.01 * 96 * 96
    paths = e.render_path_frames()
    assert paths["general"] >= 3 and paths["tiled"] <= 2, paths        # one attempt per key-frame interval (two here), then general


def test_dense_tiles_stay_on_the_tiled_path(reflib, monkeypatch):
    """Three atoms per pixel that pile up to 5800 records in one tile in mid-morph -- more than the ordering kernel's shared
    memory takes -- stay on the tiled path: the accumulating kernel streams the records from the bins.  Same frames as the general
    path and the reference; the ordering kernel (AMX_RENDER_ACC=0) has to fall back here."""
    images = scenes.ellipses(96, 2, seed=17)
    params = dict(motion=eng.LINEAR, fading=eng.LINEAR, density=3)
    m = build_ref(reflib, images, seed=1, **params)
    ts = [0.0, 0.4, 0.9]
    e = engine_from_ref(m, images, seed=1, **params)
    a = e.render(ts)
    assert e.render_path_frames() == dict(tiled=3, general=0), e.render_path_frames()
    assert e.render_tiled_stats()["max_tile"] > 4096
    monkeypatch.setenv("AMX_RENDER_TILED", "0")
    e0 = engine_from_ref(m, images, seed=1, **params)
    assert np.array_equal(a, e0.render(ts))
    assert e0.render_path_frames() == dict(tiled=0, general=3)
    monkeypatch.setenv("AMX_RENDER_TILED", "1")
    monkeypatch.setenv("AMX_RENDER_ACC", "0")
    e1 = engine_from_ref(m, images, seed=1, **params)
    assert np.array_equal(a, e1.render(ts))
    assert e1.render_path_frames()["general"] >= 3
    for k, t in enumerate(ts):
        n, mx = diff_stats(m.render(t), a[k])
        assert mx <= 1 and n <= 0.01 * 96 * 96


def test_pixel_with_hundreds_of_atoms_takes_the_general_path_for_that_call(reflib):
    """300 atoms on one pixel: the 32-bit sums of the accumulating tile kernel could wrap, so it asks for the frames of that call
    to be rendered again by the general path (64-bit sums) -- without switching the tiled path off for the table."""
    n = 16
    images = []
    for k in range(2):
        im = np.zeros((n, n, 4), dtype=np.uint8)
        im[4 + k:8 + k, 5 + k:9 + k, :3] = np.arange(48, dtype=np.uint8).reshape(4, 4, 3) * 5 + 3 * k
        im[4 + k:8 + k, 5 + k:9 + k, 3] = 255
        images.append(im)
    params = dict(motion=eng.LINEAR, fading=eng.LINEAR, density=300)
    m = build_ref(reflib, images, seed=1, **params)
    e = engine_from_ref(m, images, seed=1, **params)
    for t in (0.0, 0.5):
        n_diff, mx = diff_stats(m.render(t), e.render([t])[0])
        assert mx <= 1 and n_diff <= 4, (t, n_diff, mx)
    st, paths = e.render_tiled_stats(), e.render_path_frames()
    assert st["fallbacks"] >= 1 and not st["blocked"], st
    assert paths["general"] >= 1 and paths["tiled"] >= 2, paths


def test_tiled_diagnostics_and_kernel_times():
    """amx_render_tiled_stats / amx_kernel_times: bin occupancy of a rendered batch and per-kernel device times."""
    images = scenes.square_to_disc(128) if hasattr(scenes, "square_to_disc") else scenes.ellipses(128, 2, seed=5)
    e = eng.Engine(0, seed=1, motion=eng.SPLINE, fading=eng.COSINE, threads=0, cycle_length=1000)
    e.load_images(images)
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    e.kernel_times(True)
    frames = e.render([f / 16.0 for f in range(16)])
    kt = e.kernel_times(False)
    assert frames.shape[0] == 16
    assert kt[0]["frames"] == 16 and kt[1]["frames"] == 16 and kt[0]["launches"] == kt[1]["launches"] >= 2
    assert kt[0]["ms"] > 0 and kt[1]["ms"] > 0
    st = e.render_tiled_stats()
    assert st["fallbacks"] == 0 and not st["blocked"]
    assert 0 < st["max_bin"][0] <= 7168 and st["max_tile"] >= st["max_bin"][0]
    assert e.render_path_frames() == dict(tiled=16, general=0)


def test_tiled_path_with_several_chains_equals_general_path(reflib, monkeypatch):
    """AMX_RENDER_TILED=2 forces the tiled path for multi-chain morphs (off by default: slower there): identical frames,
    including pixels shared by several blobs, show_blobs and the background."""
    images = scenes.rect_blobs(96, 30, frames=2, seed=8, min_side=4, max_side=16)
    for params in (dict(motion=eng.LINEAR, fading=eng.LINEAR, density=2), dict(motion=eng.SPLINE, fading=eng.COSINE, show_blobs=eng.AVERAGE, keep_background=1)):
        m = build_ref(reflib, images, seed=2, match_steps=200, **params)
        ts = [m.get_time(f, 10) for f in range(10)]
        monkeypatch.setenv("AMX_RENDER_TILED", "2")
        e2 = engine_from_ref(m, images, seed=2, **params)
        assert e2.chain_count() > 1
        a = e2.render(ts)
        assert e2.render_path_frames() == dict(tiled=10, general=0)
        monkeypatch.setenv("AMX_RENDER_TILED", "1")
        e1 = engine_from_ref(m, images, seed=2, **params)
        b = e1.render(ts)
        assert e1.render_path_frames() == dict(tiled=0, general=10)
        assert np.array_equal(a, b)
        n, mx = diff_stats(m.render(ts[4]), a[4])
        assert mx <= 1 and n <= 0.01 * a[4].size


def test_full_size_frame_tiled_path_matches_reference(reflib):
    """BASELINE config 2 at its real size: one 1024 x 1024 frame of the (partly matched) square -> disc morph through the
    TILED path against am::morph::get_pixels of the reference fed the same chain table (about 10-15 s of reference time).
    Bar: <= 1 LSB per channel, differing pixels (exact .5 ties) a small fraction; cost recomputed identically."""
    size = 1024
    images = scenes.square_to_disc(size)
    params = dict(seed=1, motion=eng.SPLINE, fading=eng.COSINE)
    e = eng.Engine(0, threads=0, cycle_length=100000, **params)
    e.load_images(images)
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    e.swap_rounds(1024, want_stats=False)                 # a partly matched table: atoms cross paths, tiles are unevenly filled
    chains = e.chains()
    assert chains[0]["words"].shape == (2, size * size)
    m = reflib.RefMorph(**params)
    for k, im in enumerate(images):
        m.add_image(k, im)
    m.set_resolution(size, size)
    for i in range(2):
        ys, xs = np.nonzero(images[i][..., 3] != 0)
        labels, stats, meta = e.export_blobs(i)
        m.import_blobs(i, [dict(group=int(meta[0, 0]), stats=stats[0], surface=(ys.astype(np.uint64) * np.uint64(65536) + xs.astype(np.uint64)))])
    for c in chains:
        m.import_chain(c["key"], c["words"], c["max_surface"])
    m.finish_import()
    assert m.true_cost() == e.cost()
    t = 0.37
    got = e.render([t])[0]
    paths, tiled = e.render_path_frames(), e.render_tiled_stats()
    assert paths["tiled"] == 1 and paths["general"] == 0 and tiled["fallbacks"] == 0, (paths, tiled)
    ref = m.render(t)
    n, mx = diff_stats(ref, got)
    assert mx <= 1, "channel diff %d" % mx
    assert n <= 0.005 * size * size, "%d pixels differ" % n
    print("1024^2 frame: %d of %d pixels differ by 1 LSB (max %d)" % (n, size * size, mx))


def test_general_path_piles_and_many_blobs(reflib):
    """Several chains without feather: the general path's fused gather with its three follow-up kernels.  Big rectangles matched to
    small ones pile dozens of atoms onto a pixel near the small key frame (density 2 doubles them): positions behind more than
    MAXK records are summed by a warp each (k_resolve, heavy role), positions that three or more blobs reach are resolved from integer
    sums per blob or replayed in order (k_resolve, list role), and the homes' overflow records lie in contiguous pool ranges."""
    rng = np.random.default_rng(3)
    size = 72
    big = np.zeros((size, size, 4), dtype=np.uint8)
    small = np.zeros((size, size, 4), dtype=np.uint8)
    for k, (x, y) in enumerate(((4, 4), (38, 6), (6, 40), (40, 40))):
        big[y:y + 26, x:x + 26, :3] = rng.integers(0, 256, size=(26, 26, 3), dtype=np.uint8)
        big[y:y + 26, x:x + 26, 3] = 255 - 40 * (k & 1)
        sx, sy = 30 + 5 * (k & 1), 30 + 5 * (k >> 1)                   # the small ones lie close together: their morphs cross
        small[sy:sy + 3, sx:sx + 3, :3] = rng.integers(0, 256, size=(3, 3, 3), dtype=np.uint8)
        small[sy:sy + 3, sx:sx + 3, 3] = 255
    images = [big, small]
    params = dict(motion=eng.LINEAR, fading=eng.COSINE, density=2)
    m = build_ref(reflib, images, seed=4, match_steps=100, **params)
    e = engine_from_ref(m, images, seed=4, **params)
    assert e.chain_count() >= 4
    ts = [0.0, 0.2, 0.35, 0.45, 0.49, 0.499, 0.5, 0.6, 0.9]
    got = e.render(ts)
    total = ndiff = 0
    for i, t in enumerate(ts):
        n, mx = diff_stats(m.render(t), got[i])
        assert mx <= 1, "t=%g: channel diff %d" % (t, mx)
        ndiff += n
        total += got[i].size
    assert ndiff <= 0.01 * total
    st = e.render_stats()
    assert st["overflow"] > 2000 and st["generic"] > 100, st         # piles behind single homes, and replayed / listed positions
    assert e.render_path_frames() == dict(tiled=0, general=len(ts))
    for i in (3, 5):
        assert np.array_equal(got[i], e.render([ts[i]])[0])


def test_bin_overflow_blocks_only_its_interval(monkeypatch):
    """Three key frames, one chain: big -> big -> small.  The atoms of a covered 96 x 96 canvas (density 3) pile onto a 20 x 20
    square in the second and third interval -- more than a tile's bins hold --, while the first interval never has more than three
    atoms per pixel.  After the overflow only the piling intervals leave the tiled path; the frames equal the general path's."""
    rng = np.random.default_rng(9)
    n = 96
    images = []
    for k in range(3):
        im = np.zeros((n, n, 4), dtype=np.uint8)
        lo, hi = (0, n) if k < 2 else (38, 58)
        im[lo:hi, lo:hi, :3] = rng.integers(0, 256, size=(hi - lo, hi - lo, 3), dtype=np.uint8)
        im[lo:hi, lo:hi, 3] = 255
        images.append(im)
    params = dict(seed=2, motion=eng.LINEAR, fading=eng.LINEAR, density=3, threads=0, cycle_length=1000)
    ts = [0.05, 0.2, 0.3, 0.4, 0.6, 0.65, 0.7, 0.9]                   # intervals 0 0 0 1 1 1 2 2

    def run(tiled):
        monkeypatch.setenv("AMX_RENDER_TILED", tiled)
        e = eng.Engine(0, **params)
        e.load_images(images)
        e.step(8)
        assert e.state() == eng.STATE_ATOM_MORPHING and e.chain_count() == 1
        e.swap_rounds(64, want_stats=False)
        return e
    e1 = run("1")
    a = e1.render(ts)
    st = e1.render_tiled_stats()
    assert st["fallbacks"] == 1 and st["blocked"], st
    p1 = e1.render_path_frames()
    a2 = e1.render(ts)                                                # no second fallback: the blocked intervals go straight to the general path
    p2 = e1.render_path_frames()
    assert e1.render_tiled_stats()["fallbacks"] == 1
    assert p2["tiled"] - p1["tiled"] == 3 and p2["general"] - p1["general"] == 5, (p1, p2)
    assert np.array_equal(a, a2)
    # same table, general path only
    e0 = run("0")
    e0.import_chains(e1.chains())
    b = e0.render(ts)
    assert e0.render_path_frames()["tiled"] == 0
    assert np.array_equal(a, b)
