"""Multi-GPU partitioning of the morph path: one process per GPU, launched with
torchrun; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.

* rendering      : output frames are independent -> contiguous frame range per rank, no collective
* matching, h>=4 : key-frame columns of equal parity are independent given frozen neighbours ->
                   rank g owns columns {j : (j // 2) % G == g} of the current parity; after each
                   half-sweep the updated columns are all-gathered
* matching, h=2  : a single free column -> ATOM-range sharding: an epoch selects log2(G) index bits,
                   rank r owns the atoms whose selected bits equal r and draws pairing masks that are
                   zero on those bits; after the epoch the owned slices are exchanged with ONE
                   all-gather of the per-atom trajectory-table column (the only collective)

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


# ------------------------------------------------------------------ column ownership (h >= 4)
def owned_columns(height, parity, rank, world):
    """Columns of `parity` (0 = even, 1 = odd) that `rank` refines in this half-sweep."""
    cols = [j for j in range(height) if j % 2 == parity]
    if height % 2 == 1 and parity == 0:
        cols = cols[:-1]          # odd cycle: the last even column neighbours column 0 -> refine it with the odd ones
    return [j for i, j in enumerate(cols) if i % world == rank]


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


# ------------------------------------------------------------------ device orchestration
class ShardedMatcher:
    """Atom-range sharded pair-swap rounds for a single-chain, h = 2 morph (BASELINE config 2)."""

    def __init__(self, engine, rank, world, device=None, seed=0):
        import torch
        self.torch = torch
        self.e, self.rank, self.world, self.seed = engine, rank, world, seed
        self.width = engine.table_device_ptr(0)[1]
        self.epoch = 0
        k = max(1, int(self.width - 1).bit_length())
        s = world.bit_length() - 1
        self.slots = 1 << (k - s)
        self.device = device
        self.send = torch.empty(self.slots, dtype=torch.int64, device=device)
        self.recv = torch.empty(self.slots * world, dtype=torch.int64, device=device)
        ntiles = (1 << k) >> TILE_BITS
        self.tiled = tiled_supported(self.width) and ntiles % world == 0

    def run_epoch(self, rounds, column=1):
        """`rounds` sharded rounds on `column`, then one all-gather of that column."""
        import torch.distributed as dist
        if self.world == 1:
            self.e.swap_rounds(rounds, chain=0, column=column, want_stats=False)
            return
        if self.tiled:
            # long chain: every rank refines its share of the epoch's shared-memory tiles, then the owned slots travel
            ep = self.epoch
            self.epoch += 1
            self.e.swap_tiled_epoch(ep, rounds, column, rank=self.rank, nranks=self.world)
            n = self.e.pack_tiled(ep, column, self.rank, self.world, self.send.data_ptr())
            assert n == self.slots
            dist.all_gather_into_tensor(self.recv, self.send)      # NCCL over NVLink: W*8 B per epoch
            self.e.unpack_tiled(ep, column, self.recv.data_ptr())
            return
        mask = select_mask(self.width, self.world, self.epoch, self.seed)
        self.epoch += 1
        sel_val = int(deposit_bits(np.array([self.rank]), mask)[0])
        self.e.swap_rounds_sharded(rounds, mask, sel_val, chain=0, column=column)
        self.e.pack_owned(column, mask, sel_val, self.send.data_ptr())
        dist.all_gather_into_tensor(self.recv, self.send)          # NCCL over NVLink: W*8 B per epoch
        self.e.unpack_owned(column, mask, self.world, self.recv.data_ptr())


def broadcast_table(engine, rank, world, device):
    """Replicate rank 0's chain table on every rank (one-time, before rendering)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return
    chains = engine.chains()
    for c in chains:
        t = torch.from_numpy(c["words"].astype(np.int64)).to(device)
        dist.broadcast(t, src=0)
        c["words"] = t.cpu().numpy().astype(np.uint64)
    if rank != 0:
        engine.import_chains(chains)
