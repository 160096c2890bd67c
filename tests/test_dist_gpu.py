"""The sharded matcher of amx_dist.cu on hardware.

Single-GPU tests run everywhere a GPU is present: amx_swap_part_step / amx_swap_columns_step without a communicator are
the one-rank case of the very same code.  The world-2 tests need two GPUs (real NCCL, real peer mapping) and are skipped
otherwise; they issue their steps back to back WITHOUT any host synchronisation in between -- the ordering of
kernel -> exchange -> kernel has to come from the library's own stream (reference semantics to keep:
thread.cpp:1014-1038 -- after any number of proposals a column is the same multiset of key points and the cost never rises).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _engine(size, frames=2, device=0, seed=1):
    from atomorph_b200 import engine as eng
    from atomorph_b200 import scenes
    e = eng.Engine(device, seed=seed, motion=eng.LINEAR, fading=eng.LINEAR, threads=0, cycle_length=1000)
    e.load_images(scenes.square_to_disc(size) if frames == 2 else scenes.ellipses(size, frames, seed=3))
    e.step(8)
    assert e.state() == eng.STATE_ATOM_MORPHING
    return e


def test_part_step_single_rank_matches_the_numpy_mirror():
    from atomorph_b200 import dist as amd
    e = _engine(128)
    before = e.chains()[0]["words"].copy()
    W = before.shape[1]
    c0 = e.cost()
    h0 = [e.column_hash(j) for j in range(2)]
    st0 = e.swap_stats()
    e.swap_part_step(step=5, sub_epochs=1, rounds=64, column=1)
    st1 = e.swap_stats()
    after = e.chains()[0]["words"]
    assert np.array_equal(before[0], after[0])
    assert np.array_equal(np.sort(before[1]), np.sort(after[1]))
    h1 = [e.column_hash(j) for j in range(2)]
    assert h0[0] == h1[0] and h0[1][1] == h1[1][1] and h0[1][0] != h1[1][0]
    assert e.cost() < c0 and c0 - e.cost() == float(int(st1[2] - st0[2]))
    assert int(st1[0] - st0[0]) == 64 * W // 2
    # a key point only ever moves inside its tile: the mirror lists the atoms in tile order (tiles are aligned runs of
    # 256 / 512 / 1024 of that order, so a run of 1024 is closed under the step's swaps)
    atoms = amd.part_tile_atoms(W, 1, 0, 5, 0, 0, 1).astype(np.int64)
    for t in range(0, W, 1024):
        run = atoms[t:t + 1024]
        assert np.array_equal(np.sort(before[1][run]), np.sort(after[1][run])), "a key point left its tile"
    # several sub-epochs re-tile: still a permutation, cost keeps falling
    c1 = e.cost()
    e.swap_part_step(step=6, sub_epochs=4, rounds=32, column=1)
    assert e.cost() < c1
    assert e.column_hash(1)[1] == h0[1][1]


def test_columns_step_single_rank_refines_only_the_phase():
    e = _engine(96, frames=4)
    assert e.swap_phase_count() == 2
    before = e.chains()[0]["words"].copy()
    c0 = e.cost()
    e.swap_columns_step(phase=1, step=0, epochs=2, rounds=64)
    mid = e.chains()[0]["words"].copy()
    assert np.array_equal(before[0], mid[0]) and np.array_equal(before[2], mid[2])
    for j in (1, 3):
        assert not np.array_equal(before[j], mid[j])
        assert np.array_equal(np.sort(before[j]), np.sort(mid[j]))
    c1 = e.cost()
    assert c1 < c0
    e.swap_columns_step(phase=0, step=0, epochs=2, rounds=64)
    after = e.chains()[0]["words"]
    assert np.array_equal(mid[1], after[1]) and np.array_equal(mid[3], after[3])
    for j in (0, 2):
        assert np.array_equal(np.sort(before[j]), np.sort(after[j]))
    assert e.cost() < c1


def test_set_stream_null_is_the_legacy_stream_and_private_comes_back():
    import ctypes as C
    e = _engine(48)
    own = e.get_stream()
    assert own != 0
    e.set_stream(0)
    assert e.get_stream() == 0
    c0 = e.cost()
    e.swap_rounds(8, column=1)
    assert e.cost() <= c0
    e.set_stream(C.c_void_p(-1).value)           # AMX_STREAM_PRIVATE
    assert e.get_stream() != 0
    e.swap_rounds(8, column=1)


# ---------------------------------------------------------------------------------------------- two real GPUs
def _two_gpu_worker(rank, world, size, frames, q_id, q_out):
    try:
        import torch  # noqa: F401  (initialises CUDA the way the bench does; the exchange itself does not use torch)
        from atomorph_b200 import engine as eng
        e = _engine(size, frames=frames, device=rank)
        if rank == 0:
            ident = e.comm_unique_id()
            for _ in range(world - 1):
                q_id.put(ident)
        else:
            ident = q_id.get(timeout=120)
        e.comm_init(ident, rank, world)
        e.table_broadcast(0)
        out = {}
        h = frames
        for mode in ("nccl", "p2p"):
            if mode == "p2p":
                out["p2p_available"] = e.comm_enable_p2p()
                if not out["p2p_available"]:
                    break
            assert e.comm_info()["p2p"] == (mode == "p2p")
            multiset0 = [e.column_hash(j)[1] for j in range(h)]
            c0 = e.cost()
            st0 = e.swap_stats()
            costs = []
            # back to back, no host synchronisation between the steps
            if h == 2:
                for step in range(12):
                    e.swap_part_step(step + (100 if mode == "p2p" else 0), sub_epochs=world, rounds=64, column=1)
            else:
                for step in range(6):
                    for phase in range(e.swap_phase_count()):
                        e.swap_columns_step(phase, step + (100 if mode == "p2p" else 0), epochs=2, rounds=64)
            e.comm_check()
            c1 = e.cost()
            costs.append(c1)
            hashes = [e.column_hash(j) for j in range(h)]
            out[mode] = dict(c0=c0, c1=c1, multiset_kept=[hh[1] for hh in hashes] == multiset0, pos=[hh[0] for hh in hashes],
                             proposals=int(e.swap_stats()[0] - st0[0]))
        e.comm_destroy()
        q_out.put((rank, out))
    except Exception as ex:  # surface the failure in the parent instead of a silent timeout
        import traceback
        q_out.put((rank, {"error": "%r\n%s" % (ex, traceback.format_exc())}))


def _run_two_gpus(size, frames):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, size, frames, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, out = q_out.get(timeout=600)
        res[rank] = out
    for p in procs:
        p.join(timeout=60)
    for r in (0, 1):
        assert "error" not in res[r], res[r].get("error")
    return res


@pytest.mark.parametrize("size,frames", [(256, 2), (128, 4)])
def test_sharded_matcher_two_gpus_nccl_and_p2p(size, frames):
    res = _run_two_gpus(size, frames)
    modes = ["nccl"] + (["p2p"] if res[0].get("p2p_available") else [])
    for mode in modes:
        a, b = res[0][mode], res[1][mode]
        assert a["multiset_kept"] and b["multiset_kept"], "a key point was lost or duplicated (%s)" % mode
        assert a["pos"] == b["pos"], "the two replicas of the table differ after the %s exchange" % mode
        assert a["c0"] == b["c0"] and a["c1"] == b["c1"]
        assert a["c1"] < a["c0"], (mode, a["c0"], a["c1"])
        assert a["proposals"] > 0 and b["proposals"] > 0
    print("two-GPU matcher:", {m: (res[0][m]["c0"], res[0][m]["c1"]) for m in modes}, "p2p available:", res[0].get("p2p_available"))
