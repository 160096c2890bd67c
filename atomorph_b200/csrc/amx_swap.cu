/*
 * amx_swap.cu -- K1: atom correspondence refinement as a parallel pair-swap optimal-transport
 * local search (SURVEY.md row a-M).
 *
 * Reference: morph_asynch (thread.cpp:990-1041): uniform x1 != x2 at one key-frame column y;
 * swap the two key points of column y iff the summed squared travel to the neighbouring
 * columns does not increase (c1 >= c2, equal cost accepted).  The reference serialises the
 * swaps behind one global mutex.
 *
 * Here one ROUND evaluates a perfect matching of the atoms of a chain: atom l is paired with
 * l XOR m for a mask m drawn per round from the counter-based RNG.  Pairs of a round are
 * disjoint, columns y-1 / y+1 are read-only during a round, so every proposal sees exactly the
 * state a sequential execution in any order would see -- no locks, no races.  Over rounds every
 * pair {a,b} is proposed with equal probability (m = a^b), like the reference's uniform draw.
 * All loads are coalesced: lanes read l (consecutive) and l^m (a lane permutation of an aligned
 * block).  Cost arithmetic is exact 64-bit integer (the reference sums u64 squares in double,
 * exact below 2^53).
 *
 * Two kernels do this.  k_swap_single / k_swap_multi run ONE round per launch straight on the table in
 * L2 / HBM (32-48 B of global loads per proposal).  k_swap_tiled is the fast path for a long chain: an
 * EPOCH scatters the atoms of the chain into tiles of 2^TB atoms through a per-epoch random bijection of
 * the atom index (so that any two atoms share a tile with the same probability), each CTA loads its tile
 * ONCE into shared memory (key points unpacked to 24-bit fixed point), runs R rounds of disjoint
 * XOR-pairings inside the tile from shared memory, and writes the column back.  Global traffic per
 * proposal drops by R (16-24 B per atom per epoch instead of 32-48 B per proposal) and a proposal costs
 * about 45 instructions, so the kernel is bound by the issue rate, not by memory.  Over epochs every pair
 * of atoms is proposed with equal probability, like the reference's uniform draw.
 *
 * Algorithmic bytes per proposal: 32 B (h = 2) or 48 B (h >= 3) read,
 * +16 B written per accepted swap.
 */
#include <algorithm>
#include <cub/cub.cuh>
#include "amx_engine.h"
#include "amx_swap.h"

namespace amx {

struct SwapStats { unsigned long long proposals, accepted, gain; };

// proposals / accepted / gain: warp shuffle, then shared memory, then ONE set of atomics per block
// (same-address atomics serialise in L2 at about one per clock: per-warp atomics used to dominate a round)
__device__ __forceinline__ void warp_add_stats(unsigned long long *stats, unsigned prop, unsigned acc, unsigned long long gain) {
    __shared__ unsigned s_prop[32], s_acc[32];
    __shared__ unsigned long long s_gain[32];
    for (int o = 16; o > 0; o >>= 1) {
        prop += __shfl_down_sync(0xffffffffu, prop, o);
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        gain += __shfl_down_sync(0xffffffffu, gain, o);
    }
    unsigned warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_prop[warp] = prop; s_acc[warp] = acc; s_gain[warp] = gain; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned nw = (blockDim.x + 31) >> 5;
        unsigned p = 0, a = 0;
        unsigned long long g = 0;
        for (unsigned i = 0; i < nw; ++i) { p += s_prop[i]; a += s_acc[i]; g += s_gain[i]; }
        if (p) atomicAdd(stats + 0, (unsigned long long) p);
        if (a) { atomicAdd(stats + 1, (unsigned long long) a); atomicAdd(stats + 2, g); }
    }
}

// evaluate the swap of rows i, j at column `col`; h2: prev == next column
__device__ __forceinline__ bool propose(pword *__restrict__ col, const pword *__restrict__ prev, const pword *__restrict__ next,
                                        bool h2, size_t i, size_t j, unsigned long long *gain) {
    pword a = col[i], b = col[j];
    pword an = next[i], bn = next[j];
    unsigned long long c1, c2;
    if (h2) {
        c1 = 2ull * (point_distance(a, an) + point_distance(b, bn));
        c2 = 2ull * (point_distance(b, an) + point_distance(a, bn));
    } else {
        pword ap = prev[i], bp = prev[j];
        c1 = point_distance(ap, a) + point_distance(a, an) + point_distance(bp, b) + point_distance(b, bn);
        c2 = point_distance(ap, b) + point_distance(b, an) + point_distance(bp, a) + point_distance(a, bn);
    }
    if (c1 >= c2) {           // thread.cpp:1022 -- equal cost is accepted
        col[i] = b;
        col[j] = a;
        *gain = c1 - c2;
        return true;
    }
    return false;
}

__device__ __forceinline__ uint64_t insert_zero_bit(uint64_t t, unsigned b) {
    uint64_t lo = t & ((1ull << b) - 1ull);
    return ((t >> b) << (b + 1)) | lo;
}

// one chain [off, off+w): thread t handles the pair (l, l^m), l = t with a zero inserted at the top bit of m
__global__ void __launch_bounds__(256)
k_swap_single(pword *__restrict__ col, const pword *__restrict__ prev, const pword *__restrict__ next, int h2, uint64_t off,
              uint64_t w, uint64_t m, unsigned topbit, uint64_t npairs, unsigned long long *__restrict__ stats) {
    unsigned prop = 0, acc = 0;
    unsigned long long gain = 0;
    for (uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; t < npairs; t += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t l = insert_zero_bit(t, topbit);
        uint64_t r = l ^ m;
        if (r < w) {
            unsigned long long g1 = 0;
            prop += 1;
            if (propose(col, prev, next, h2 != 0, off + l, off + r, &g1)) { acc += 1; gain += g1; }
        }
    }
    warp_add_stats(stats, prop, acc, gain);
}

// all chains at once: thread a = atom; the lower row of each pair does the work
__global__ void __launch_bounds__(256)
k_swap_multi(pword *__restrict__ col, const pword *__restrict__ prev, const pword *__restrict__ next, int h2,
             const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off, uint64_t A, uint64_t seed,
             uint64_t round, unsigned long long *__restrict__ stats) {
    uint64_t a = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned prop = 0, acc = 0;
    unsigned long long gain = 0;
    if (a < A) {
        uint32_t c = chain_of[a];
        uint64_t off = chain_off[c], w = chain_off[c + 1] - off;
        if (w >= 2) {
            unsigned k = 64 - __clzll(w - 1);                 // 2^k >= w
            uint64_t m = 1ull + rng64(seed, 0x5157u + c, round) % ((1ull << k) - 1ull);
            uint64_t l = a - off, r = l ^ m;
            if (r > l && r < w) {
                prop = 1;
                acc = propose(col, prev, next, h2 != 0, a, off + r, &gain) ? 1u : 0u;
            }
        }
    }
    warp_add_stats(stats, prop, acc, gain);
}

// ---- shared-memory tiled rounds -----------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t tile_atom(const TileMap &tm, uint32_t u) {
    if (tm.perm) return tm.perm[(u + tm.shift) & tm.mask];
    if (tm.imask) {
        uint32_t lo = ((u & tm.imask) * tm.ia1) & tm.imask;
        lo ^= lo >> tm.is1;
        lo = (lo * tm.ia2 + tm.ic) & tm.imask;
        lo ^= lo >> tm.is2;
        u = (u & ~tm.imask) | lo;
    }
    uint32_t v = (u * tm.a1) & tm.mask;
    v ^= v >> tm.s1;
    v = (v * tm.a2 + tm.c) & tm.mask;
    v ^= v >> tm.s2;
    return v;
}
TileMap make_tilemap(uint64_t seed, uint64_t stream, uint64_t epoch, unsigned k) {
    TileMap tm;
    uint64_t r1 = rng64(seed, 0x7111u + stream, epoch), r2 = rng64(seed, 0x7222u + stream, epoch);
    tm.mask = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
    tm.a1 = ((uint32_t) r1 | 1u) & tm.mask;
    tm.a2 = ((uint32_t) (r1 >> 32) | 1u) & tm.mask;
    tm.c = (uint32_t) r2 & tm.mask;
    tm.s1 = std::max(1u, k / 2);
    tm.s2 = std::max(1u, (k + 1) / 2);
    tm.ia1 = tm.ia2 = 1u; tm.ic = 0u; tm.is1 = tm.is2 = 1u; tm.imask = 0u;
    tm.perm = nullptr; tm.shift = 0u;
    return tm;
}
// inner bijection of the low `ik` bits for sub-epoch `sub` of a step
void tilemap_set_inner(TileMap &tm, uint64_t seed, uint64_t stream, uint64_t sub, unsigned ik) {
    uint64_t r1 = rng64(seed, 0x7333u + stream, sub), r2 = rng64(seed, 0x7444u + stream, sub);
    tm.imask = ik >= 32 ? 0xffffffffu : ((1u << ik) - 1u);
    tm.ia1 = ((uint32_t) r1 | 1u) & tm.imask;
    tm.ia2 = ((uint32_t) (r1 >> 32) | 1u) & tm.imask;
    tm.ic = (uint32_t) r2 & tm.imask;
    tm.is1 = std::max(1u, ik / 2);
    tm.is2 = std::max(1u, (ik + 1) / 2);
}

// key point in shared memory: x in 1/256 px (24 bits) | flags << 24 in the low word, y in 1/256 px in the high word
__device__ __forceinline__ uint2 kp_unpack(pword w) {
    return make_uint2((uint32_t) pw_x256(w) | (pw_flags(w) << 24), (uint32_t) pw_y256(w));
}
__device__ __forceinline__ pword kp_pack(uint2 p) {
    uint32_t x = p.x & 0xffffffu, f = p.x >> 24;
    return pw_make(x >> 8, p.y >> 8, x & 255u, p.y & 255u, f);
}
__device__ __forceinline__ unsigned long long kp_dist(uint2 a, uint2 b) {
    // 24-bit coordinates: the differences fit 32 bits, each square is ONE 32 x 32 -> 64 bit multiply-add
    const int dx = (int) (a.x & 0xffffffu) - (int) (b.x & 0xffffffu);
    const int dy = (int) a.y - (int) b.y;
    return (unsigned long long) ((long long) dx * (long long) dx) + (unsigned long long) ((long long) dy * (long long) dy);
}

// One CTA = one tile of 2^TB atoms, NT threads.  `rounds` rounds of 2^(TB-1) disjoint proposals each, all inside shared
// memory.  Slots whose atom index falls behind the chain (w not a power of two) are marked invalid and never proposed.
template <bool H2, int TB, int NT>
__global__ void __launch_bounds__(NT)
k_swap_tiled(pword *__restrict__ col, const pword *__restrict__ prev, const pword *__restrict__ next, uint64_t off, uint32_t w,
             TileMap tm, uint32_t tile0, uint32_t rounds, uint64_t seed, uint64_t round_base, unsigned long long *__restrict__ stats,
             PeerCols peers) {
    constexpr uint32_t TA = 1u << TB;
    extern __shared__ uint2 sm[];
    uint2 *s_col = sm, *s_next = sm + TA, *s_prev = sm + 2 * TA;     // s_prev only when !H2
    const uint32_t tile = tile0 + blockIdx.x;
    const uint32_t ubase = tile << TB;
    // load: atom index from the bijection (recomputed at write-back), invalid slots get flag byte 0xff
    for (uint32_t j = threadIdx.x; j < TA; j += NT) {
        uint32_t a = tile_atom(tm, ubase + j);
        if (a < w) {
            s_col[j] = kp_unpack(col[off + a]);
            s_next[j] = kp_unpack(next[off + a]);
            if (!H2) s_prev[j] = kp_unpack(prev[off + a]);
        } else {
            s_col[j] = make_uint2(0xff000000u, 0u);
        }
    }
    // pairing masks of the epoch's rounds inside this tile: uniform over the non-zero TB-bit values
    __shared__ uint32_t s_mask[TILE_MAX_ROUNDS];
    for (uint32_t r = threadIdx.x; r < rounds; r += NT)
        s_mask[r] = 1u + (uint32_t) (rng64(seed, 0x5157u + tile, round_base + r) % (TA - 1u));
    __syncthreads();
    unsigned prop = 0, acc = 0;
    unsigned long long gain = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t m = s_mask[r];
        const unsigned topbit = 31u - __clz(m);
        const uint32_t lowmask = (1u << topbit) - 1u;
#pragma unroll
        for (uint32_t q = 0; q < TA / 2 / NT; ++q) {
            const uint32_t t = q * NT + threadIdx.x;
            const uint32_t i = ((t & ~lowmask) << 1) | (t & lowmask);        // t with a zero inserted at m's top bit
            const uint32_t j = i ^ m;
            const uint2 a = s_col[i], b = s_col[j];
            if ((a.x >> 24) == 0xffu || (b.x >> 24) == 0xffu) continue;       // padding slot
            const uint2 an = s_next[i], bn = s_next[j];
            unsigned long long c1, c2;
            if (H2) {
                c1 = kp_dist(a, an) + kp_dist(b, bn);
                c2 = kp_dist(b, an) + kp_dist(a, bn);
            } else {
                const uint2 ap = s_prev[i], bp = s_prev[j];
                c1 = kp_dist(ap, a) + kp_dist(a, an) + kp_dist(bp, b) + kp_dist(b, bn);
                c2 = kp_dist(ap, b) + kp_dist(b, an) + kp_dist(bp, a) + kp_dist(a, bn);
            }
            ++prop;
            if (c1 >= c2) {                // thread.cpp:1022 -- equal cost is accepted
                s_col[i] = b; s_col[j] = a;
                ++acc;
                gain += H2 ? 2ull * (c1 - c2) : (c1 - c2);
            }
        }
        __syncthreads();
    }
    for (uint32_t j = threadIdx.x; j < TA; j += NT) {
        uint32_t a = tile_atom(tm, ubase + j);
        if (a < w) {
            const pword word = kp_pack(s_col[j]);
            col[off + a] = word;
            // NVLink stores, fire and forget: contiguous per tile into the peers' staging buffers, or straight to the atom's place
            const size_t at = peers.staged ? (size_t) ubase + j : (size_t) off + a;
            for (uint32_t p = 0; p < peers.n; ++p) peers.dst[p][at] = word;
        }
    }
    warp_add_stats(stats, prop, acc, gain);
}

// launch one epoch's kernel for tiles [t0, t0 + ntl) of 2^tb atoms
void launch_swap_tiled(Engine *E, bool h2, int tb, pword *col, const pword *prev, const pword *next, uint64_t off, uint32_t w, const TileMap &tm,
                       uint32_t t0, uint32_t ntl, uint32_t rounds, uint64_t round_base, const PeerCols &peers) {
    unsigned long long *st = (unsigned long long *) E->d_swapstats;
    const size_t smem = (size_t) (h2 ? 2 : 3) * ((size_t) 1 << tb) * sizeof(uint2);
#define AMX_SWT(H, TB, NT) do { \
        cudaFuncSetAttribute(k_swap_tiled<H, TB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        k_swap_tiled<H, TB, NT><<<ntl, NT, smem, E->stream>>>(col, prev, next, off, w, tm, t0, rounds, E->p.seed, round_base, st, peers); } while (0)
    if (tb == 10) { if (h2) AMX_SWT(true, 10, 256); else AMX_SWT(false, 10, 256); }
    else if (tb == 9) { if (h2) AMX_SWT(true, 9, 256); else AMX_SWT(false, 9, 256); }
    else { if (h2) AMX_SWT(true, 8, 128); else AMX_SWT(false, 8, 128); }
#undef AMX_SWT
    E->launches++;
}

// owned slots of an epoch <-> contiguous buffer (multi-GPU: rank r runs tiles [r * T / N, (r + 1) * T / N))
__global__ void __launch_bounds__(256)
k_pack_tiled(const pword *__restrict__ col, uint64_t off, uint32_t w, TileMap tm, uint32_t u0, uint32_t n, pword *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = tile_atom(tm, u0 + i);
    out[i] = a < w ? col[off + a] : 0ull;
}
__global__ void __launch_bounds__(256)
k_unpack_tiled(pword *__restrict__ col, uint64_t off, uint32_t w, TileMap tm, uint32_t u0, uint32_t n, uint32_t skip0, uint32_t skip1,
               const pword *__restrict__ in) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i += u0;                                   // slots [u0, u0 + n) of the buffer (indexed by slot), without [skip0, skip1)
    if (i >= skip0 && i < skip1) return;       // (the caller's own part: already in place)
    uint32_t a = tile_atom(tm, i);
    if (a < w) col[off + a] = in[i];
}

void launch_pack_tiled(Engine *E, const pword *col, uint64_t off, uint32_t w, const TileMap &tm, uint32_t u0, uint32_t n, pword *out) {
    if (n == 0) return;
    k_pack_tiled<<<div_up(n, 256), 256, 0, E->stream>>>(col, off, w, tm, u0, n, out);
    E->launches++;
}
void launch_unpack_tiled(Engine *E, pword *col, uint64_t off, uint32_t w, const TileMap &tm, uint32_t u0, uint32_t n, uint32_t skip0, uint32_t skip1,
                         const pword *in) {
    if (n == 0) return;
    k_unpack_tiled<<<div_up(n, 256), 256, 0, E->stream>>>(col, off, w, tm, u0, n, skip0, skip1, in);
    E->launches++;
}

// ---- multi-GPU atom-range sharding (h = 2: a single free column) ----------------------
// An EPOCH fixes `sel_mask` (log2 N index bits): rank r owns the atoms whose selected bits equal r and only
// draws pairing masks with zeros on those bits, so every pair stays inside the rank's slice.  Between epochs
// the owned slices are exchanged with one all-gather of the column (pack -> NCCL -> unpack).
__device__ __forceinline__ uint64_t deposit_bits(uint64_t u, uint64_t free_mask) {     // software pdep
    uint64_t out = 0;
    while (free_mask) {
        uint64_t bit = free_mask & (~free_mask + 1);
        if (u & 1ull) out |= bit;
        u >>= 1;
        free_mask &= free_mask - 1;
    }
    return out;
}

// thread t enumerates the owned pairs: free-bit index with a zero inserted at the (free) position of m's top bit
__global__ void __launch_bounds__(256)
k_swap_sharded(pword *__restrict__ col, const pword *__restrict__ prev, const pword *__restrict__ next, int h2, uint64_t off, uint64_t w,
               uint64_t m, unsigned top_free_pos, uint64_t free_mask, uint64_t sel_val, uint64_t npairs, unsigned long long *__restrict__ stats) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned prop = 0, acc = 0;
    unsigned long long gain = 0;
    if (t < npairs) {
        uint64_t l = deposit_bits(insert_zero_bit(t, top_free_pos), free_mask) | sel_val;
        uint64_t r = l ^ m;
        if (l < w && r < w) {
            prop = 1;
            acc = propose(col, prev, next, h2 != 0, off + l, off + r, &gain) ? 1u : 0u;
        }
    }
    warp_add_stats(stats, prop, acc, gain);
}

// owned atoms of a column <-> contiguous buffer (slot u of rank r = atom deposit(u, free_mask) | r's bits)
__global__ void __launch_bounds__(256)
k_pack_owned(const pword *__restrict__ col, uint64_t off, uint64_t w, uint64_t free_mask, uint64_t sel_val, uint64_t n, pword *__restrict__ out) {
    uint64_t u = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    uint64_t l = deposit_bits(u, free_mask) | sel_val;
    out[u] = l < w ? col[off + l] : 0ull;
}
__global__ void __launch_bounds__(256)
k_unpack_owned(pword *__restrict__ col, uint64_t off, uint64_t w, uint64_t free_mask, uint64_t sel_mask, uint64_t n, uint32_t nranks,
               const pword *__restrict__ in) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * nranks) return;
    uint64_t r = i / n, u = i % n;
    uint64_t l = deposit_bits(u, free_mask) | deposit_bits(r, sel_mask);
    if (l < w) col[off + l] = in[i];
}

// cost partial sums: sum_j d(p[x][j], p[x][j+1 mod h]) over atoms of chains with width > 1 (thread.cpp:1109-1125)
__global__ void __launch_bounds__(256)
k_cost(const pword *__restrict__ table, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off, uint64_t A,
       uint32_t h, unsigned long long *__restrict__ partials) {
    __shared__ unsigned long long sh[8];
    unsigned long long s = 0;
    for (uint64_t a = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; a < A; a += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t c = chain_of[a];
        if (chain_off[c + 1] - chain_off[c] <= 1) continue;
        pword first = table[a], cur = first;
        for (uint32_t j = 1; j < h; ++j) {
            pword nx = table[(size_t) j * A + a];
            s += point_distance(cur, nx);
            cur = nx;
        }
        s += point_distance(cur, first);     // j = h-1 -> 0 (for h == 1 this is d(p,p) = 0)
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        partials[blockIdx.x] = t;
    }
}

unsigned ceil_log2(uint64_t w) { unsigned k = 0; while ((1ull << k) < w) ++k; return k; }

// true when the chain is long enough for the shared-memory tiled path
bool tiled_ok(Engine *E, uint32_t chain) {
    uint64_t w = E->chain_off[chain + 1] - E->chain_off[chain];
    return w >= 4ull * TILE_ATOMS && w <= 0x80000000ull;
}

// One EPOCH of the tiled path on column y of `chain`: `rounds` rounds (TILE_MAX_ROUNDS per launch) inside every tile of the
// rank's share [rank * T / nranks, (rank + 1) * T / nranks) of the T tiles.
int engine_swap_tiled_epoch(Engine *E, uint32_t chain, uint32_t y, uint64_t epoch, uint32_t rounds, uint32_t rank, uint32_t nranks) {
    if (chain >= E->nchains || y >= E->h || rounds == 0 || nranks == 0 || rank >= nranks || !tiled_ok(E, chain))
        return AMX_ERR_ARG;
    const uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    const unsigned k = ceil_log2(w);
    const uint32_t ntiles = 1u << (k - TILE_BITS);
    const uint32_t t0 = (uint32_t) ((uint64_t) ntiles * rank / nranks), t1 = (uint32_t) ((uint64_t) ntiles * (rank + 1) / nranks);
    if (t1 <= t0) return AMX_OK;
    const TileMap tm = make_tilemap(E->p.seed, chain, epoch, k);
    const uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
    pword *col = E->table + (size_t) y * E->A;
    const pword *prev = E->table + (size_t) yp * E->A, *next = E->table + (size_t) yn * E->A;
    const bool h2 = E->h == 2;
    PeerCols nopeers; nopeers.n = 0; nopeers.staged = 0;
    // more than TILE_MAX_ROUNDS rounds: further launches on the SAME tiles (same bijection) with fresh pairing masks
    for (uint32_t done = 0; done < rounds; done += TILE_MAX_ROUNDS) {
        const uint32_t r = std::min<uint32_t>(rounds - done, TILE_MAX_ROUNDS);
        const uint64_t round_base = (epoch << 20) + done;
        launch_swap_tiled(E, h2, TILE_BITS, col, prev, next, off, (uint32_t) w, tm, t0, t1 - t0, r, round_base, nopeers);
    }
    E->render_ready = false;
    return E->check("tiled swap epoch") ? AMX_ERR_CUDA : AMX_OK;
}

// ---- locality epochs ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton_spread(uint32_t v) {        // 16 bits -> every other bit
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
__global__ void __launch_bounds__(256)
k_local_keys(const pword *__restrict__ col, uint64_t off, uint32_t w, uint32_t n2k, uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n2k) return;
    if (u < w) { const pword p = col[off + u]; key[u] = morton_spread(pw_x(p)) | (morton_spread(pw_y(p)) << 1); val[u] = u; }
    else { key[u] = 0xffffffffu; val[u] = 0xffffffffu; }          // padding slot: behind every atom, never proposed
}

// One locality epoch on column y of `chain`: the atoms are sorted by the Morton code of their CURRENT column-y position,
// consecutive runs of 1024 (shifted by a per-epoch random offset) form the tiles, and `rounds` rounds of random pairings
// run inside every tile exactly as in a uniform epoch.  Proposals between spatial neighbours are the ones that still pay
// late in the refinement, when uniform partners are accepted with probability ~ 1/proposals-per-atom.
int engine_swap_local_epoch(Engine *E, uint32_t chain, uint32_t y, uint64_t epoch, uint32_t rounds) {
    if (chain >= E->nchains || y >= E->h || rounds == 0 || !tiled_ok(E, chain)) return AMX_ERR_ARG;
    const uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    const unsigned k = ceil_log2(w);
    const uint32_t n2k = 1u << k, ntiles = 1u << (k - TILE_BITS);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t *) nullptr, (uint32_t *) nullptr, (uint32_t *) nullptr, (uint32_t *) nullptr, (int) n2k, 0, 32, E->stream);
    if (E->loc_cap < n2k || E->loc_tmp_bytes < tmp_bytes) {
        dev_free(E->loc_buf); dev_free(E->loc_tmp); E->loc_buf = nullptr; E->loc_tmp = nullptr; E->loc_cap = 0; E->loc_tmp_bytes = 0;
        if (!dev_alloc(E, (void **) &E->loc_buf, (size_t) n2k * 16, "locality keys") || !dev_alloc(E, &E->loc_tmp, tmp_bytes, "locality sort")) return AMX_ERR_NOMEM;
        E->loc_cap = n2k; E->loc_tmp_bytes = tmp_bytes;
    }
    uint32_t *key = E->loc_buf, *key2 = key + n2k, *val = key2 + n2k, *perm = val + n2k;
    pword *col = E->table + (size_t) y * E->A;
    k_local_keys<<<div_up(n2k, 256), 256, 0, E->stream>>>(col, off, (uint32_t) w, n2k, key, val);
    cub::DeviceRadixSort::SortPairs(E->loc_tmp, tmp_bytes, key, key2, val, perm, (int) n2k, 0, 32, E->stream);
    E->launches += 2;
    TileMap tm = make_tilemap(E->p.seed, chain, epoch, k);
    tm.perm = perm;
    tm.shift = (uint32_t) rng64(E->p.seed, 0x10ca1u + chain, epoch) & tm.mask;
    const uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
    const pword *prev = E->table + (size_t) yp * E->A, *next = E->table + (size_t) yn * E->A;
    const bool h2 = E->h == 2;
    PeerCols nopeers; nopeers.n = 0; nopeers.staged = 0;
    // the positions move while the tiles are refined: the order is rebuilt per epoch, further launches reuse it
    for (uint32_t done = 0; done < rounds; done += TILE_MAX_ROUNDS) {
        const uint32_t r = std::min<uint32_t>(rounds - done, TILE_MAX_ROUNDS);
        const uint64_t round_base = (epoch << 20) + done + (1ull << 19);
        launch_swap_tiled(E, h2, TILE_BITS, col, prev, next, off, (uint32_t) w, tm, 0u, ntiles, r, round_base, nopeers);
    }
    E->render_ready = false;
    return E->check("locality swap epoch") ? AMX_ERR_CUDA : AMX_OK;
}

int engine_swap_rounds(Engine *E, int32_t chain, int32_t column, uint64_t rounds) {
    if (E->nchains == 0 || E->h < 2) return AMX_OK;
    if (chain >= (int32_t) E->nchains || column >= (int32_t) E->h) return AMX_ERR_ARG;
    bool h2 = E->h == 2;
    if ((chain >= 0 || E->nchains == 1) && tiled_ok(E, chain >= 0 ? (uint32_t) chain : 0u) && !E->swap_global_only) {
        // long single chain: epochs of up to TILE_MAX_ROUNDS rounds in shared memory (one column per epoch)
        const uint32_t c = chain >= 0 ? (uint32_t) chain : 0u;
        while (rounds > 0) {
            const uint32_t r = (uint32_t) std::min<uint64_t>(rounds, TILE_MAX_ROUNDS);
            const uint64_t epoch = E->rng_round++;
            const uint32_t y = column >= 0 ? (uint32_t) column : (uint32_t) (rng64(E->p.seed, 0xc01u, epoch) % E->h);
            // every `swap_locality`-th epoch pairs spatial neighbours instead of uniform partners (default: never)
            const bool local = E->swap_locality > 0 && (epoch % E->swap_locality) == E->swap_locality - 1;
            int rcode = local ? engine_swap_local_epoch(E, c, y, epoch, r) : engine_swap_tiled_epoch(E, c, y, epoch, r, 0, 1);
            if (rcode != AMX_OK) return rcode;
            rounds -= r;
        }
        return AMX_OK;
    }
    for (uint64_t r = 0; r < rounds; ++r) {
        uint64_t round = E->rng_round++;
        uint32_t y = column >= 0 ? (uint32_t) column : (uint32_t) (rng64(E->p.seed, 0xc01u, round) % E->h);
        uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
        pword *col = E->table + (size_t) y * E->A;
        const pword *prev = E->table + (size_t) yp * E->A, *next = E->table + (size_t) yn * E->A;
        if (chain >= 0 || E->nchains == 1) {
            uint32_t c = chain >= 0 ? (uint32_t) chain : 0u;
            uint64_t off = E->chain_off[c], w = E->chain_off[c + 1] - off;
            if (w < 2) continue;
            unsigned k = 0;
            while ((1ull << k) < w) ++k;
            uint64_t m = 1ull + rng64(E->p.seed, 0x5157u + c, round) % ((1ull << k) - 1ull);
            unsigned topbit = 63 - __builtin_clzll(m);
            uint64_t npairs = 1ull << (k - 1);
            unsigned nb = (unsigned) std::min<uint64_t>(div_up(npairs, 256), 148 * 8);     // persistent: 8 CTAs of 256 threads per SM
            k_swap_single<<<nb, 256, 0, E->stream>>>(col, prev, next, h2, off, w, m, topbit, npairs, (unsigned long long *) E->d_swapstats);
        } else {
            k_swap_multi<<<div_up(E->A, 256), 256, 0, E->stream>>>(col, prev, next, h2, E->chain_of, E->d_chain_off, E->A, E->p.seed, round,
                                                                 (unsigned long long *) E->d_swapstats);
        }
        E->launches++;
    }
    E->render_ready = false;
    return E->check("swap rounds") ? AMX_ERR_CUDA : AMX_OK;
}

static bool shard_geometry(Engine *E, uint32_t chain, uint64_t sel_mask, unsigned *k, uint64_t *free_mask) {
    uint64_t w = E->chain_off[chain + 1] - E->chain_off[chain];
    if (w < 2) return false;
    *k = 0;
    while ((1ull << *k) < w) ++*k;
    uint64_t all = (1ull << *k) - 1ull;
    if (sel_mask & ~all) return false;
    *free_mask = all & ~sel_mask;
    return *free_mask != 0;
}

int engine_swap_rounds_sharded(Engine *E, uint32_t chain, int32_t column, uint64_t rounds, uint64_t sel_mask, uint64_t sel_val) {
    if (E->nchains == 0 || E->h < 2 || chain >= E->nchains || column >= (int32_t) E->h) return AMX_ERR_ARG;
    unsigned k; uint64_t free_mask;
    if (!shard_geometry(E, chain, sel_mask, &k, &free_mask) || (sel_val & ~sel_mask)) return AMX_ERR_ARG;
    unsigned nfree = __builtin_popcountll(free_mask);
    uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    bool h2 = E->h == 2;
    for (uint64_t r = 0; r < rounds; ++r) {
        uint64_t round = E->rng_round++;
        uint32_t y = column >= 0 ? (uint32_t) column : (uint32_t) (rng64(E->p.seed, 0xc01u, round) % E->h);
        uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
        // mask: uniform over the non-zero values of the free bits
        uint64_t mfree = 1ull + rng64(E->p.seed ^ sel_val, 0x5157u + chain, round) % ((1ull << nfree) - 1ull);
        unsigned top_free_pos = 63 - __builtin_clzll(mfree);
        uint64_t m = 0, fm = free_mask, bits = mfree;
        while (fm) { uint64_t bit = fm & (~fm + 1); if (bits & 1ull) m |= bit; bits >>= 1; fm &= fm - 1; }
        uint64_t npairs = 1ull << (nfree - 1);
        k_swap_sharded<<<div_up(npairs, 256), 256, 0, E->stream>>>(E->table + (size_t) y * E->A, E->table + (size_t) yp * E->A, E->table + (size_t) yn * E->A,
                                                                  h2, off, w, m, top_free_pos, free_mask, sel_val, npairs, (unsigned long long *) E->d_swapstats);
        E->launches++;
    }
    E->render_ready = false;
    return E->check("sharded swap rounds") ? AMX_ERR_CUDA : AMX_OK;
}

int engine_cost(Engine *E, double *cost) {
    *cost = 0.0;
    if (E->nchains == 0 || E->A == 0) return AMX_OK;
    uint32_t nb = (uint32_t) std::min<uint64_t>(div_up(E->A, 256), 148 * 8);
    if (E->n_partials < nb) {
        dev_free(E->d_partials);
        if (!dev_alloc(E, (void **) &E->d_partials, nb * 8, "partials")) return AMX_ERR_NOMEM;
        E->n_partials = nb;
    }
    k_cost<<<nb, 256, 0, E->stream>>>(E->table, E->chain_of, E->d_chain_off, E->A, E->h, (unsigned long long *) E->d_partials);
    E->launches++;
    std::vector<uint64_t> host(nb);
    if (E->fail(cudaMemcpyAsync(host.data(), E->d_partials, nb * 8, cudaMemcpyDeviceToHost, E->stream), "cost D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "cost"))
        return AMX_ERR_CUDA;
    unsigned __int128 total = 0;
    for (uint64_t v : host) total += v;
    *cost = (double) total;
    return AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_swap_rounds(amx_ctx *ctx, int32_t chain, int32_t column, uint64_t rounds, uint64_t stats3[3]) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    int rc = engine_swap_rounds(E, chain, column, rounds);
    if (rc != AMX_OK) return rc;
    if (stats3) return amx_swap_stats(ctx, stats3);
    return AMX_OK;
}

int amx_swap_stats(amx_ctx *ctx, uint64_t stats3[3]) {
    if (!ctx || !stats3) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (E->fail(cudaMemcpyAsync(E->swapstats, E->d_swapstats, 24, cudaMemcpyDeviceToHost, E->stream), "stats D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "stats"))
        return AMX_ERR_CUDA;
    for (int i = 0; i < 3; ++i) stats3[i] = E->swapstats[i];
    return AMX_OK;
}

int amx_swap_rounds_sharded(amx_ctx *ctx, uint32_t chain, int32_t column, uint64_t rounds, uint64_t sel_mask, uint64_t sel_val) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_swap_rounds_sharded(&ctx->e, chain, column, rounds, sel_mask, sel_val);
}

int amx_swap_tiled_epoch(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rounds, uint32_t rank, uint32_t nranks) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_swap_tiled_epoch(&ctx->e, chain, column, epoch, rounds, rank, nranks);
}

int amx_swap_local_epoch(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rounds) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_swap_local_epoch(&ctx->e, chain, column, epoch, rounds);
}
int amx_set_swap_locality(amx_ctx *ctx, uint32_t every) {
    if (!ctx) return AMX_ERR_ARG;
    ctx->e.swap_locality = every;
    return AMX_OK;
}

// the slots u of the epoch's bijection that rank `rank` owns, as a contiguous buffer (count = its tiles * 2^TILE_BITS)
int amx_pack_tiled(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rank, uint32_t nranks, void *d_out, uint64_t *count) {
    if (!ctx || !d_out || chain >= ctx->e.nchains || column >= ctx->e.h || nranks == 0 || rank >= nranks || !tiled_ok(&ctx->e, chain)) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    const uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    const unsigned k = ceil_log2(w);
    const uint32_t ntiles = 1u << (k - TILE_BITS);
    const uint32_t t0 = (uint32_t) ((uint64_t) ntiles * rank / nranks), t1 = (uint32_t) ((uint64_t) ntiles * (rank + 1) / nranks);
    const uint32_t n = (t1 - t0) << TILE_BITS;
    if (count) *count = n;
    if (n == 0) return AMX_OK;
    k_pack_tiled<<<div_up(n, 256), 256, 0, E->stream>>>(E->table + (size_t) column * E->A, off, (uint32_t) w, make_tilemap(E->p.seed, chain, epoch, k), t0 << TILE_BITS, n, (pword *) d_out);
    E->launches++;
    return E->check("pack tiled") ? AMX_ERR_CUDA : AMX_OK;
}

// scatter the all-gathered slots (all ranks, slot order) back into the column
int amx_unpack_tiled(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, const void *d_in) {
    if (!ctx || !d_in || chain >= ctx->e.nchains || column >= ctx->e.h || !tiled_ok(&ctx->e, chain)) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    const uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    const unsigned k = ceil_log2(w);
    const uint32_t n = 1u << k;
    k_unpack_tiled<<<div_up(n, 256), 256, 0, E->stream>>>(E->table + (size_t) column * E->A, off, (uint32_t) w, make_tilemap(E->p.seed, chain, epoch, k), 0u, n, 0u, 0u, (const pword *) d_in);
    E->launches++;
    E->render_ready = false;
    return E->check("unpack tiled") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_pack_owned(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t sel_mask, uint64_t sel_val, void *d_out, uint64_t *count) {
    if (!ctx || !d_out || chain >= ctx->e.nchains || column >= ctx->e.h) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    unsigned k; uint64_t free_mask;
    if (!shard_geometry(E, chain, sel_mask, &k, &free_mask)) return AMX_ERR_ARG;
    uint64_t n = 1ull << __builtin_popcountll(free_mask), off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    k_pack_owned<<<div_up(n, 256), 256, 0, E->stream>>>(E->table + (size_t) column * E->A, off, w, free_mask, sel_val, n, (pword *) d_out);
    E->launches++;
    if (count) *count = n;
    return E->check("pack owned") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_unpack_owned(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t sel_mask, uint32_t nranks, const void *d_in) {
    if (!ctx || !d_in || chain >= ctx->e.nchains || column >= ctx->e.h) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    unsigned k; uint64_t free_mask;
    if (!shard_geometry(E, chain, sel_mask, &k, &free_mask) || nranks != (1u << __builtin_popcountll(sel_mask))) return AMX_ERR_ARG;
    uint64_t n = 1ull << __builtin_popcountll(free_mask), off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    k_unpack_owned<<<div_up(n * nranks, 256), 256, 0, E->stream>>>(E->table + (size_t) column * E->A, off, w, free_mask, sel_mask, n, nranks, (const pword *) d_in);
    E->launches++;
    E->render_ready = false;
    return E->check("unpack owned") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_cost(amx_ctx *ctx, double *cost) {
    if (!ctx || !cost) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_cost(&ctx->e, cost);
}

}
