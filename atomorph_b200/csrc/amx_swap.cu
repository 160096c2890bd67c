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
 * Algorithmic bytes per proposal: 32 B (h = 2) or 48 B (h >= 3) read,
 * +16 B written per accepted swap.
 */
#include "amx_engine.h"

namespace amx {

struct SwapStats { unsigned long long proposals, accepted, gain; };

__device__ __forceinline__ void warp_add_stats(unsigned long long *stats, unsigned prop, unsigned acc, unsigned long long gain) {
    for (int o = 16; o > 0; o >>= 1) {
        prop += __shfl_down_sync(0xffffffffu, prop, o);
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        gain += __shfl_down_sync(0xffffffffu, gain, o);
    }
    if ((threadIdx.x & 31) == 0 && prop) {
        atomicAdd(stats + 0, (unsigned long long) prop);
        if (acc) { atomicAdd(stats + 1, (unsigned long long) acc); atomicAdd(stats + 2, gain); }
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
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned prop = 0, acc = 0;
    unsigned long long gain = 0;
    if (t < npairs) {
        uint64_t l = insert_zero_bit(t, topbit);
        uint64_t r = l ^ m;
        if (r < w) {
            prop = 1;
            acc = propose(col, prev, next, h2 != 0, off + l, off + r, &gain) ? 1u : 0u;
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

int engine_swap_rounds(Engine *E, int32_t chain, int32_t column, uint64_t rounds) {
    if (E->nchains == 0 || E->h < 2) return AMX_OK;
    if (chain >= (int32_t) E->nchains || column >= (int32_t) E->h) return AMX_ERR_ARG;
    bool h2 = E->h == 2;
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
            k_swap_single<<<div_up(npairs, 256), 256, 0, E->stream>>>(col, prev, next, h2, off, w, m, topbit, npairs, (unsigned long long *) E->d_swapstats);
        } else {
            k_swap_multi<<<div_up(E->A, 256), 256, 0, E->stream>>>(col, prev, next, h2, E->chain_of, E->d_chain_off, E->A, E->p.seed, round,
                                                                 (unsigned long long *) E->d_swapstats);
        }
        E->launches++;
    }
    E->render_ready = false;
    return E->check("swap rounds") ? AMX_ERR_CUDA : AMX_OK;
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

int amx_cost(amx_ctx *ctx, double *cost) {
    if (!ctx || !cost) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_cost(&ctx->e, cost);
}

}
