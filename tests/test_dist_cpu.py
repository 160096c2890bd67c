"""Host-side logic of the multi-GPU partitioning (atomorph_b200/dist.py) on CPU: index arithmetic, and a
world_size-2 gloo run of the atom-range sharded epoch (numpy stand-in for the device rounds)."""
import os
import socket

import numpy as np
import pytest

from atomorph_b200 import dist as amd


def test_frame_ranges_cover_exactly():
    for total in (1, 7, 64, 512):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b = amd.frame_range(total, r, world)
                seen += list(range(a, b))
            assert seen == list(range(total))
    t = np.concatenate([amd.frame_times(64, r, 4) for r in range(4)])
    assert np.array_equal(t, np.arange(64) / 64.0)


def test_owned_columns_cover_every_column_and_never_touch_neighbours():
    for h in (2, 3, 4, 5, 7, 8):
        for world in (1, 2, 4, 8):
            covered = []
            for phase in range(amd.phase_count(h)):
                cols = sorted(sum((amd.owned_columns(h, phase, r, world) for r in range(world)), []))
                assert cols == sorted(amd.phase_columns(h, phase))
                covered += cols
                # no two columns refined together are cyclic neighbours (they are read-only for each other's proposals)
                for a in cols:
                    for b in cols:
                        assert a == b or ((a - b) % h not in (1, h - 1))
            assert sorted(covered) == list(range(h)), "a key-frame column is never refined"
    assert amd.phase_count(8) == 2 and amd.phase_count(5) == 3


@pytest.mark.parametrize("width", [8192, 10000, 1 << 15])
def test_parts_of_a_step_partition_the_chain(width):
    """amx_swap_part_step: the outer bijection's contiguous slot ranges are disjoint pseudo-random parts; a sub-epoch's
    inner bijection only permutes a part."""
    k = int(width - 1).bit_length()
    for step in range(3):
        slots = amd.part_slots(width, seed=1, chain=0, step=step)
        assert np.array_equal(np.sort(slots), np.arange(1 << k, dtype=np.uint64))
        for world in (1, 2, 4, 8):
            n = (1 << k) // world
            for r in range(world):
                for sub in range(2):
                    atoms = amd.part_tile_atoms(width, 1, 0, step, sub, r, world)
                    assert np.array_equal(np.sort(atoms), np.sort(slots[r * n:(r + 1) * n]))
                a0, a1 = amd.part_tile_atoms(width, 1, 0, step, 0, r, world), amd.part_tile_atoms(width, 1, 0, step, 1, r, world)
                assert not np.array_equal(a0, a1), "sub-epochs must re-tile the part"
        live = np.sort(slots[: (1 << k) // 2][slots[: (1 << k) // 2] < width].astype(np.int64))
        assert live[-1] - live[0] > width // 2                      # a part is spread over the whole chain
    assert not np.array_equal(amd.part_slots(width, 1, 0, 0), amd.part_slots(width, 1, 0, 1))


@pytest.mark.parametrize("width", [1 << 10, 1000, 37])
def test_owned_atoms_partition_the_chain(width):
    for world in (2, 4, 8):
        if world >= width:
            continue
        for epoch in range(5):
            mask = amd.select_mask(width, world, epoch, seed=3)
            assert bin(mask).count("1") == world.bit_length() - 1
            allidx = np.concatenate([amd.owned_atoms(width, mask, r)[0] for r in range(world)])
            k = int(width - 1).bit_length()
            assert np.array_equal(np.sort(allidx), np.arange(1 << k, dtype=np.uint64))
            for r in range(world):
                idx, val = amd.owned_atoms(width, mask, r)
                assert np.all((idx & np.uint64(mask)) == np.uint64(val))


@pytest.mark.parametrize("width", [8192, 10000, 1 << 15])
def test_tile_slots_are_a_bijection(width):
    k = int(width - 1).bit_length()
    for epoch in range(4):
        v = amd.tile_slots(width, seed=1, chain=0, epoch=epoch)
        assert np.array_equal(np.sort(v), np.arange(1 << k, dtype=np.uint64))
        # tiles are pseudo-random subsets: the atoms of one tile are spread over the whole chain
        t0 = np.sort(v[: 1 << amd.TILE_BITS].astype(np.int64))
        assert t0[-1] - t0[0] > (1 << k) // 2
    assert not np.array_equal(amd.tile_slots(width, 1, 0, 0), amd.tile_slots(width, 1, 0, 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, width, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    column = rng.integers(0, 2 ** 48, width, dtype=np.uint64)       # identical on both ranks
    other = rng.integers(0, 2 ** 20, width, dtype=np.uint64)
    original = column.copy()
    for epoch in range(3):
        mask = amd.select_mask(width, world, epoch, seed=9)
        idx, val = amd.owned_atoms(width, mask, rank)
        ok = idx < width
        # stand-in for the device rounds: sort the owned slots by a key (any permutation inside the slice)
        own = idx[ok].astype(np.int64)
        vals = column[own]
        order = np.argsort(other[own] ^ np.uint64(epoch), kind="stable")
        column[own] = vals[order]
        send = np.zeros(len(idx), dtype=np.int64)
        send[ok] = column[own].astype(np.int64)
        recv = [torch.zeros(len(idx), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(recv, torch.from_numpy(send))
        column = amd.scatter_gathered(column, [r.numpy().astype(np.uint64) for r in recv], width, mask, world)
    # tile-range sharding (long chains): rank r owns the slots [r * 2^k / world, (r + 1) * 2^k / world) of the epoch's bijection
    if width >= 64:
        k = int(width - 1).bit_length()
        n = (1 << k) // world
        for epoch in range(3):
            slots = amd.tile_slots(width, seed=9, chain=0, epoch=epoch)
            mine = slots[rank * n:(rank + 1) * n]
            ok = mine < width
            own = mine[ok].astype(np.int64)
            vals = column[own]
            column[own] = vals[np.argsort(other[own] ^ np.uint64(epoch + 7), kind="stable")]
            send = np.zeros(n, dtype=np.int64)
            send[ok] = column[own].astype(np.int64)
            recv = [torch.zeros(n, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(recv, torch.from_numpy(send))
            allv = np.concatenate([r.numpy().astype(np.uint64) for r in recv])
            okall = slots < width
            column[slots[okall].astype(np.int64)] = allv[okall]
    # parts of a step (amx_swap_part_step): rank r refines its slot range through two re-tiled sub-epochs, then all-gather
    if width >= 64:
        k = int(width - 1).bit_length()
        n = (1 << k) // world
        for step in range(3):
            slots = amd.part_slots(width, seed=9, chain=0, step=step)
            for sub in range(2):
                atoms = amd.part_tile_atoms(width, 9, 0, step, sub, rank, world)
                own = atoms[atoms < width].astype(np.int64)
                vals = column[own]
                column[own] = vals[np.argsort(other[own] ^ np.uint64(step * 2 + sub + 11), kind="stable")]
            mine = slots[rank * n:(rank + 1) * n]
            ok = mine < width
            send = np.zeros(n, dtype=np.int64)
            send[ok] = column[mine[ok].astype(np.int64)].astype(np.int64)
            recv = [torch.zeros(n, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(recv, torch.from_numpy(send))
            allv = np.concatenate([r.numpy().astype(np.uint64) for r in recv])
            okall = slots < width
            column[slots[okall].astype(np.int64)] = allv[okall]
    q.put((rank, column, original))
    dist.destroy_process_group()


@pytest.mark.parametrize("width", [256, 200])
def test_sharded_epoch_world2_gloo(width):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, width, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, col, orig = q.get(timeout=120)
        res[rank] = (col, orig)
    for p in procs:
        p.join(timeout=60)
    assert np.array_equal(res[0][0], res[1][0]), "ranks disagree after the all-gather"
    assert np.array_equal(np.sort(res[0][0]), np.sort(res[0][1])), "a key point was lost or duplicated"
    assert not np.array_equal(res[0][0], res[0][1])
