"""K1 parity: the parallel pair-swap matcher against the reference's serial stochastic matcher
(SURVEY.md row a-M).  The serial RNG stream cannot be reproduced in parallel, so the criterion is
the transport cost (recomputed from the table, reference metric point_distance)."""
import numpy as np
import pytest

from atomorph_b200 import engine as eng
from atomorph_b200 import scenes
from helpers import build_ref, engine_from_ref

pytestmark = pytest.mark.gpu


def _rows_are_permutation(before, after):
    for j in range(before.shape[0]):
        if not np.array_equal(np.sort(before[j]), np.sort(after[j])):
            return False
    return True


@pytest.mark.parametrize("frames", [2, 4])
def test_cost_reaches_reference(reflib, frames):
    images = scenes.ellipses(40, frames, seed=30 + frames)
    m = build_ref(reflib, images, seed=1)
    e = engine_from_ref(m, images, seed=1)
    before = e.chains()[0]["words"].copy()
    W = before.shape[1]
    c0 = e.cost()
    assert c0 == m.true_cost()
    # reference: 2000 proposals per atom
    m.set(cycle_length=W)
    m.sync()
    m.iterate(2000)
    m.sync()
    c_ref = m.true_cost()
    # engine: same number of proposals
    st0 = e.swap_stats()
    rounds = 0
    while int(e.swap_stats()[0] - st0[0]) < 2000 * W:
        e.swap_rounds(200)
        rounds += 200
        assert rounds < 200000
    st = e.swap_stats()
    c_gpu = e.cost()
    after = e.chains()[0]["words"]
    assert _rows_are_permutation(before, after), "a swap round lost or duplicated a key point"
    # the gain bookkeeping is exact integer arithmetic
    assert c0 - c_gpu == float(int(st[2] - st0[2]))
    assert c_gpu <= 1.01 * c_ref, (c_gpu, c_ref)


def test_step_counts_proposals(reflib):
    images = scenes.ellipses(32, 2, seed=40)
    m = build_ref(reflib, images, seed=1)
    e = engine_from_ref(m, images, seed=1, threads=0, cycle_length=5000)
    assert e.state() == eng.STATE_ATOM_MORPHING
    st0 = e.swap_stats()
    e.step(10)
    st = e.swap_stats()
    assert int(st[0] - st0[0]) >= 10 * 5000
    assert int(st[0] - st0[0]) <= 10 * 5000 + 2 * e.chains()[0]["width"]


def test_multi_chain_sweep(reflib):
    images = scenes.rect_blobs(64, 10, frames=2, seed=6, min_side=5, max_side=14)
    m = build_ref(reflib, images, seed=3, match_steps=100)
    e = engine_from_ref(m, images, seed=3)
    befores = [c["words"].copy() for c in e.chains()]
    c0 = e.cost()
    assert c0 == m.true_cost()
    e.swap_rounds(500)
    c1 = e.cost()
    assert c1 <= c0
    for b, a in zip(befores, e.chains()):
        assert _rows_are_permutation(b, a["words"])


def test_sharded_rounds_and_pack_unpack(reflib):
    """Atom-range sharding on ONE GPU: the slices of 2 (and 4) virtual ranks are refined one after the other on the
    same table (what the ranks do concurrently on their own copies), then packed, concatenated like an
    all-gather and unpacked into a second engine."""
    import torch
    from atomorph_b200 import dist as amd
    images = scenes.ellipses(40, 2, seed=50)
    m = build_ref(reflib, images, seed=1)
    e = engine_from_ref(m, images, seed=1)
    e2 = engine_from_ref(m, images, seed=1)
    before = e.chains()[0]["words"].copy()
    W = before.shape[1]
    c0 = e.cost()
    for world in (2, 4):
        for epoch in range(6):
            mask = amd.select_mask(W, world, epoch, seed=1)
            sends = []
            for r in range(world):
                idx, val = amd.owned_atoms(W, mask, r)
                st0 = e.swap_stats()
                e.swap_rounds_sharded(50, mask, val, chain=0, column=1)
                assert int(e.swap_stats()[0] - st0[0]) > 0
                buf = torch.empty(len(idx), dtype=torch.int64, device="cuda")
                torch.cuda.synchronize()
                n = e.pack_owned(1, mask, val, buf.data_ptr())
                e.sync()
                assert n == len(idx)
                col = e.chains()[0]["words"][1]
                ok = idx < W
                assert np.array_equal(buf.cpu().numpy().astype(np.uint64)[ok], col[idx[ok].astype(np.int64)])
                sends.append(buf)
            gathered = torch.cat(sends)
            torch.cuda.synchronize()
            e2.unpack_owned(1, mask, world, gathered.data_ptr())
            e2.sync()
            assert np.array_equal(e2.chains()[0]["words"][1], e.chains()[0]["words"][1])
    after = e.chains()[0]["words"]
    assert _rows_are_permutation(before, after)
    assert e.cost() < c0


def _big_engine(size=128, seed=1, **kw):
    """square -> disc at `size`^2 through the whole device pipeline (size^2 atoms, one chain, h = 2)."""
    e = eng.Engine(0, seed=seed, motion=eng.LINEAR, fading=eng.LINEAR, threads=0, cycle_length=1000, **kw)
    e.load_images(scenes.square_to_disc(size))
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    return e


def test_tiled_path_is_exact_and_converges_like_global_rounds(monkeypatch):
    """The shared-memory tiled epochs (chains >= 8192 atoms) against the one-round-per-launch kernel on the same table and
    the same number of proposals: rows stay permutations, the gain bookkeeping is exact, the cost curve is the same."""
    e = _big_engine(128)
    start = e.chains()
    before = start[0]["words"].copy()
    W = before.shape[1]
    assert W == 128 * 128
    c0 = e.cost()
    st0 = e.swap_stats()
    e.swap_rounds(2048, column=1)            # tiled: 32 epochs x 64 rounds
    st = e.swap_stats()
    c_tiled = e.cost()
    after = e.chains()[0]["words"]
    assert _rows_are_permutation(before, after), "a tiled epoch lost or duplicated a key point"
    assert np.array_equal(before[0], after[0]), "column 0 must not move"
    assert int(st[0] - st0[0]) == 2048 * W // 2
    assert c0 - c_tiled == float(int(st[2] - st0[2]))
    monkeypatch.setenv("AMX_SWAP_GLOBAL", "1")
    g = _big_engine(128)
    g.import_chains([dict(key=start[0]["key"], words=before, max_surface=start[0]["max_surface"])])
    assert g.cost() == c0
    g.swap_rounds(2048, column=1)
    c_global = g.cost()
    assert c_tiled <= 1.02 * c_global, (c_tiled, c_global)
    assert c_global <= 1.02 * c_tiled, (c_tiled, c_global)


def test_tiled_epochs_shard_across_ranks():
    """Two virtual ranks on one GPU: each runs its half of the tiles of an epoch on its own copy of the table, the owned
    slots are packed, concatenated like an all-gather and unpacked: both copies end up identical and improved."""
    import torch
    from atomorph_b200 import dist as amd
    ea, eb = _big_engine(128), _big_engine(128)
    start = ea.chains()
    eb.import_chains([dict(key=start[0]["key"], words=start[0]["words"], max_surface=start[0]["max_surface"])])
    c0 = ea.cost()
    assert eb.cost() == c0
    W = start[0]["width"]
    for epoch in range(8):
        bufs = []
        for r, e in enumerate((ea, eb)):
            e.swap_tiled_epoch(epoch, 32, 1, rank=r, nranks=2)
            buf = torch.empty(W // 2, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            assert e.pack_tiled(epoch, 1, r, 2, buf.data_ptr()) == W // 2
            e.sync()
            # the numpy mirror of the bijection (dist.tile_slots) names the same atoms
            slots = amd.tile_slots(W, 1, 0, epoch)[r * (W // 2):(r + 1) * (W // 2)].astype(np.int64)
            assert np.array_equal(buf.cpu().numpy().astype(np.uint64), e.chains()[0]["words"][1][slots])
            bufs.append(buf)
        gathered = torch.cat(bufs)
        torch.cuda.synchronize()
        for e in (ea, eb):
            e.unpack_tiled(epoch, 1, gathered.data_ptr())
            e.sync()
        wa, wb = ea.chains()[0]["words"], eb.chains()[0]["words"]
        assert np.array_equal(wa, wb)
    assert np.array_equal(np.sort(wa[1]), np.sort(start[0]["words"][1]))
    assert ea.cost() < 0.5 * c0


def test_locality_epochs_converge_faster_and_keep_the_permutation():
    """Locality-biased proposals (default off): every other epoch pairs spatial neighbours.  Same
    number of proposals, same start: the cost must end lower than with uniform partners only, the gain bookkeeping stays
    exact and every column stays a permutation of its key points."""
    images = scenes.square_to_disc(256)
    params = dict(seed=1, motion=eng.LINEAR, fading=eng.LINEAR, threads=0, cycle_length=1000)

    def fresh():
        e = eng.Engine(0, **params)
        e.load_images(images)
        e.step(8)
        assert e.state() == eng.STATE_ATOM_MORPHING
        return e

    ROUNDS = 64 * 48
    a = fresh()
    before = a.chains()[0]["words"].copy()
    c0 = a.cost()
    sa0 = a.swap_stats()
    a.swap_rounds(ROUNDS, column=1)
    n_uniform = int(a.swap_stats()[0] - sa0[0])
    c_uniform = a.cost()

    b = fresh()
    assert b.cost() == c0
    b.set_swap_locality(2)
    st0 = b.swap_stats()
    b.swap_rounds(ROUNDS, column=1)
    st = b.swap_stats()
    c_mixed = b.cost()
    after = b.chains()[0]["words"]
    assert _rows_are_permutation(before, after)
    assert c0 - c_mixed == float(int(st[2] - st0[2]))
    assert int(st[0] - st0[0]) == n_uniform                          # the same number of proposals
    assert c_mixed < c_uniform < c0, (c0, c_uniform, c_mixed)
    print("cost: start %.4g, uniform %.4g, with locality epochs %.4g (%.1f %% lower)" % (c0, c_uniform, c_mixed, 100.0 * (1.0 - c_mixed / c_uniform)))

    # the explicit entry point: one locality epoch on its own lowers the cost as well
    c = fresh()
    c.swap_local_epoch(0, 64, column=1)
    assert c.cost() < c0
