"""Multi-GPU partitioning of the morph path: one process per GPU, launched with
torchrun.  Every collective of the path is issued by the LIBRARY (atomorph_b200/csrc/amx_dist.cu) on the
engine's own stream; `torch.distributed` is only used to hand the NCCL unique id to the ranks, for barriers
and for reducing timings (gloo in the CPU tests).

* rendering      : output frames are independent -> contiguous frame range per rank, no collective
* matching, h>=3 : key-frame columns of one PHASE (even / odd / last column of an odd cycle) are independent
                   given frozen neighbours -> rank g refines every G-th column of the phase, then the refined
                   columns go to every replica (amx_swap_columns_step)
* matching, h=2  : a single free column -> the ATOMS are split: a per-step bijection of the atom index gives
                   rank r a pseudo-random 1/G of the atoms, which it refines for several re-tiled epochs in
                   shared memory; then the parts are exchanged (amx_swap_part_step)

The exchange is either written through peer-mapped replicas from inside the swap kernel (P2P over NVLink,
amx_comm_enable_p2p) or pack -> ncclAllGather -> unpack / ncclBroadcast, both on the engine's stream.

The index arithmetic (which atoms a rank owns, how the gathered buffers map back) is mirrored here in
numpy so that the world_size-2 gloo tests can check it without a GPU.
"""
import numpy as np


# ------------------------------------------------------------------ frame ranges
def frame_range(total_frames, rank, world):
    """Contiguous block of output frames of `rank` (sizes differ by at most one)."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def frame_times(total_frames, rank, world, finite=False, key_frames=2):
    """Times of the frames of `rank`, as morph::get_time maps them (reference morph.cpp:1522-1538)."""
    a, b = frame_range(total_frames, rank, world)
    out = []
    for f in range(a, b):
        if not finite:
            out.append(f / float(total_frames))
        elif total_frames == 1:
            out.append((1.0 - 1.0 / key_frames) / 2.0)
        else:
            out.append(f / float(total_frames - 1) * (1.0 - 1.0 / key_frames))
    return np.array(out, dtype=np.float64)


# ------------------------------------------------------------------ column ownership (h >= 3)
def phase_count(height):
    """Half-sweeps of a cycle of `height` key-frame columns (amx_swap_phase_count)."""
    return 0 if height < 2 else (3 if height % 2 else 2)


def phase_columns(height, phase):
    """Columns refined together in `phase`: 0 = even, 1 = odd, 2 = the last column of an odd cycle (it neighbours
    column 0, so it cannot join the even ones)."""
    if phase == 2:
        return [height - 1] if height % 2 and height >= 3 else []
    return [j for j in range(phase, height, 2) if not (phase == 0 and height % 2 and height >= 3 and j == height - 1)]


def owned_columns(height, phase, rank, world):
    """Columns of `phase` that `rank` refines (every world-th one, amx_swap_columns_step)."""
    return [j for i, j in enumerate(phase_columns(height, phase)) if i % world == rank]


# ------------------------------------------------------------------ atom-range ownership (h = 2)
def deposit_bits(u, mask):
    """Software pdep: spread the low bits of `u` over the set bits of `mask` (numpy, vectorised)."""
    u = np.asarray(u, dtype=np.uint64)
    out = np.zeros_like(u)
    bit_index = 0
    m = int(mask)
    while m:
        bit = m & (-m)
        out |= ((u >> np.uint64(bit_index)) & np.uint64(1)) * np.uint64(bit)
        bit_index += 1
        m &= m - 1
    return out


def select_mask(width, world, epoch, seed=0):
    """log2(world) index bits chosen for `epoch` (deterministic on every rank)."""
    k = max(1, int(width - 1).bit_length())
    s = int(world).bit_length() - 1
    if (1 << s) != world or s >= k:
        raise ValueError("world size must be a power of two smaller than the chain width")
    rng = np.random.default_rng([int(seed) & 0xffffffff, int(epoch)])
    bits = rng.choice(k, size=s, replace=False)
    mask = 0
    for b in bits:
        mask |= 1 << int(b)
    return mask


def owned_atoms(width, sel_mask, rank):
    """Atom indices (ascending slot order u = 0,1,...) of `rank` for the epoch's `sel_mask`."""
    k = max(1, int(width - 1).bit_length())
    free_mask = ((1 << k) - 1) & ~sel_mask
    n = 1 << bin(free_mask).count("1")
    sel_val = int(deposit_bits(np.array([rank]), sel_mask)[0])
    idx = deposit_bits(np.arange(n, dtype=np.uint64), free_mask) | np.uint64(sel_val)
    return idx, sel_val


def scatter_gathered(column, gathered, width, sel_mask, world):
    """numpy mirror of amx_unpack_owned: gathered[r] holds the owned slots of rank r."""
    for r in range(world):
        idx, _ = owned_atoms(width, sel_mask, r)
        ok = idx < width
        column[idx[ok].astype(np.int64)] = gathered[r][ok]
    return column


# ------------------------------------------------------------------ tile-range ownership (long chains, amx_swap.cu k_swap_tiled)
TILE_BITS = 10
_M64 = (1 << 64) - 1


def _mix64(z):
    z = (z + 0x9e3779b97f4a7c15) & _M64
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & _M64
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & _M64
    return z ^ (z >> 31)


def _rng64(seed, stream, counter):
    return _mix64(_mix64((seed ^ ((stream * 0xd1342543de82ef95) & _M64)) & _M64) ^ ((counter * 0x2545f4914f6cdd1d) & _M64))


def tile_slots(width, seed, chain, epoch):
    """numpy mirror of the epoch's bijection (amx_swap.cu: make_tilemap / tile_atom): atom index of every slot u in
    [0, 2^k).  Slots whose atom is >= width are padding.  Rank r of n owns slots [r * 2^k / n, (r + 1) * 2^k / n)."""
    k = max(1, int(width - 1).bit_length())
    mask = (1 << k) - 1 if k < 32 else 0xffffffff
    r1, r2 = _rng64(seed, 0x7111 + chain, epoch), _rng64(seed, 0x7222 + chain, epoch)
    a1 = ((r1 & 0xffffffff) | 1) & mask
    a2 = (((r1 >> 32) & 0xffffffff) | 1) & mask
    c = (r2 & 0xffffffff) & mask
    s1, s2 = max(1, k // 2), max(1, (k + 1) // 2)
    u = np.arange(1 << k, dtype=np.uint64)
    v = (u * np.uint64(a1)) & np.uint64(mask)
    v ^= v >> np.uint64(s1)
    v = (v * np.uint64(a2) + np.uint64(c)) & np.uint64(mask)
    v ^= v >> np.uint64(s2)
    return v


def tiled_supported(width):
    return width >= 4 * (1 << TILE_BITS)


# ------------------------------------------------------------------ parts of a step (amx_dist.cu: engine_swap_part_step)
def _affine_xorshift(u, k, r1, r2):
    mask = (1 << k) - 1 if k < 32 else 0xffffffff
    a1 = ((r1 & 0xffffffff) | 1) & mask
    a2 = (((r1 >> 32) & 0xffffffff) | 1) & mask
    c = (r2 & 0xffffffff) & mask
    s1, s2 = max(1, k // 2), max(1, (k + 1) // 2)
    v = (u * np.uint64(a1)) & np.uint64(mask)
    v ^= v >> np.uint64(s1)
    v = (v * np.uint64(a2) + np.uint64(c)) & np.uint64(mask)
    v ^= v >> np.uint64(s2)
    return v


def part_slots(width, seed, chain, step):
    """Atom index of every slot u in [0, 2^k) of a step's OUTER bijection.  Rank r of n owns the contiguous slot
    range [r 2^k / n, (r + 1) 2^k / n): a pseudo-random 1/n of the atoms (slots whose atom is >= width are padding)."""
    k = max(1, int(width - 1).bit_length())
    st = 0x100 + chain
    u = np.arange(1 << k, dtype=np.uint64)
    return _affine_xorshift(u, k, _rng64(seed, 0x7111 + st, step), _rng64(seed, 0x7222 + st, step))


def part_tile_atoms(width, seed, chain, step, sub, rank, world):
    """Atoms of rank `rank` in sub-epoch `sub` of `step`, in tile order (consecutive runs of 2^tb are one tile): the
    INNER bijection re-tiles the rank's slot range, the outer one names the atoms."""
    k = max(1, int(width - 1).bit_length())
    s = int(world).bit_length() - 1
    kk = k - s
    st = 0x100 + chain
    lo = np.arange(1 << kk, dtype=np.uint64)
    sub_id = step * 4096 + sub
    lo = _affine_xorshift(lo, kk, _rng64(seed, 0x7333 + st, sub_id), _rng64(seed, 0x7444 + st, sub_id))
    u = (np.uint64(rank) << np.uint64(kk)) | lo
    return part_slots(width, seed, chain, step)[u.astype(np.int64)]


# ------------------------------------------------------------------ device orchestration
def init_comm(engine, rank, world, device=None):
    """Join the library's NCCL communicator: rank 0 creates the unique id, torch.distributed carries it."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return
    ident = engine.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(ident.copy())
    if device is not None and dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=0)
    engine.comm_init(t.cpu().numpy(), rank, world)


class ShardedMatcher:
    """The sharded pair-swap matcher: a thin caller of amx_swap_part_step (h = 2) / amx_swap_columns_step (h >= 3).
    Pack, collective and unpack (or the P2P write-through and its flag barrier) all run inside the library on the
    engine's stream -- nothing here touches a stream."""

    def __init__(self, engine, rank, world, device=None, seed=0, p2p=True):
        self.e, self.rank, self.world = engine, rank, world
        self.step = 0
        self.p2p = False
        if world > 1:
            if engine.comm_info()["nranks"] != world:
                init_comm(engine, rank, world, device)
            self.p2p = bool(p2p) and engine.comm_enable_p2p()

    def run_step(self, sub_epochs=None, rounds=64, column=1, chain=0):
        """h = 2: one step on `column` -- every rank refines its 1/world of the atoms for `sub_epochs` (default: world)
        re-tiled epochs of `rounds` rounds, then the parts are exchanged.  Weak-scaling unit: a rank proposes
        sub_epochs * rounds * (W / world) / 2 pairs per step, the same at every world size when sub_epochs = world."""
        sub = self.world if sub_epochs is None else int(sub_epochs)
        self.e.swap_part_step(self.step, sub, rounds, column=column, chain=chain)
        self.step += 1

    def run_sweep(self, epochs=1, rounds=64, chain=-1):
        """h >= 3: one sweep = every phase once (each column refined for epochs * rounds rounds by its owner)."""
        for phase in range(self.e.swap_phase_count()):
            self.e.swap_columns_step(phase, self.step, epochs, rounds, chain=chain)
        self.step += 1

    # kept for callers of the round-1 interface
    def run_epoch(self, rounds, column=1):
        left = int(rounds)
        while left > 0:
            r = min(left, 64)
            self.run_step(sub_epochs=1, rounds=r, column=column)
            left -= r


def broadcast_table(engine, rank, world, device=None):
    """Replicate rank 0's chain table on every rank (one ncclBroadcast inside the library)."""
    if world == 1:
        return
    if engine.comm_info()["nranks"] != world:
        init_comm(engine, rank, world, device)
    engine.table_broadcast(0)
