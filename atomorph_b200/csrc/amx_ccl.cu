/*
 * amx_ccl.cu -- K2: blob segmentation and blob statistics (SURVEY.md row a-B1).
 *
 * Reference: thread::blobify_frame (thread.cpp:225-412) -- stochastic agglomerative region
 * merging, one merge per step.  With the library defaults (blob_threshold = 1.0,
 * blob_max_size = SIZE_MAX, blob_min_size = 1, blob_number = 1) every pair of 4-adjacent
 * pixels ends in the same blob, i.e. the final partition is exactly the set of 4-connected
 * components of the presence mask (measured).  That deterministic regime is
 * what is bit-exact here; the other regimes (blob_number > 1, blob_threshold < 1, size limits)
 * apply the reference's merge rules in parallel rounds (agglomerate_frame below) and are
 * compared statistically only -- the reference's own result depends on its RNG order.
 *
 * Algorithm: lock-free union-find (label = smaller root wins through atomicMin), first per
 * 32x32-pixel tile in shared memory on tile-local indices, then across the tile borders in
 * global memory, then path compression; the root of a component is its smallest canvas index
 * (= smallest xy2pos, the canonical label of SURVEY.md M2).  Statistics are exact integer
 * sums (count, x, y, r, g, b, a) reduced per warp with __match_any_sync before one atomic per
 * distinct root, then divided in double -- the reference reaches the same means through a
 * chain of pairwise weighted averages (thread.cpp:348-357).
 *
 * Blob vector order: ascending canonical label (the reference's order is an RNG shuffle).
 * Algorithmic bytes: ~16 B per canvas position per key frame.
 */
#include <algorithm>
#include <cub/cub.cuh>
#include "amx_engine.h"

namespace amx {

__device__ __forceinline__ uint32_t uf_find(const uint32_t *lab, uint32_t i) {
    uint32_t p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(uint32_t *lab, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }     // a > b: hang a under b
        uint32_t old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;                                           // somebody else re-rooted a meanwhile
    }
}

__global__ void __launch_bounds__(256)
k_ccl_init(const uint8_t *__restrict__ present, uint32_t *__restrict__ lab, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lab[i] = present[i] ? (uint32_t) i : 0xffffffffu;
}

__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t *__restrict__ present, const uint32_t *__restrict__ stored, uint32_t *__restrict__ lab, uint32_t cw, uint32_t ch,
            int use_threshold, double threshold) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) cw * ch || !present[i]) return;
    uint32_t x = (uint32_t) (i % cw), y = (uint32_t) (i / cw);
    if (x > 0 && present[i - 1] && (!use_threshold || color_distance(stored[i], stored[i - 1]) <= threshold)) uf_union(lab, (uint32_t) i, (uint32_t) i - 1);
    if (y > 0 && present[i - cw] && (!use_threshold || color_distance(stored[i], stored[i - cw]) <= threshold)) uf_union(lab, (uint32_t) i, (uint32_t) (i - cw));
}

// Tiled labelling: one CTA per 32x32-pixel tile runs the same union-find on TILE-LOCAL indices in shared memory (the
// contention of a large component stays on chip), flattens it and writes, for every pixel, the canvas index of the
// smallest pixel of its tile-local component -- the tile-local row-major order equals the canvas order inside a tile, so
// that is the component's canonical label.  Only the pixels in the first column / row of a tile then need a global union
// with their left / upper neighbour (k_ccl_border): 1/16 of the global atomics of the one-pass version.
#define CCL_T 32u
__device__ __forceinline__ uint32_t suf_find(const uint32_t *lab, uint32_t i) {
    uint32_t p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}
__device__ __forceinline__ void suf_union(uint32_t *lab, uint32_t a, uint32_t b) {
    for (;;) {
        a = suf_find(lab, a);
        b = suf_find(lab, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        uint32_t old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
k_ccl_tile(const uint8_t *__restrict__ present, const uint32_t *__restrict__ stored, uint32_t *__restrict__ lab, uint32_t cw, uint32_t ch,
           int use_threshold, double threshold, int merge) {
    __shared__ uint32_t s_lab[CCL_T * CCL_T];
    __shared__ uint32_t s_col[CCL_T * CCL_T];
    __shared__ uint8_t s_pre[CCL_T * CCL_T];
    const uint32_t x0 = blockIdx.x * CCL_T, y0 = blockIdx.y * CCL_T;
    for (uint32_t l = threadIdx.x; l < CCL_T * CCL_T; l += 256u) {
        const uint32_t x = x0 + (l & 31u), y = y0 + (l >> 5);
        const bool in = x < cw && y < ch;
        const size_t g = (size_t) y * cw + x;
        const bool pr = in && present[g];
        s_pre[l] = pr ? 1 : 0;
        s_col[l] = pr ? stored[g] : 0u;
        s_lab[l] = l;
    }
    __syncthreads();
    if (merge) {
        for (uint32_t l = threadIdx.x; l < CCL_T * CCL_T; l += 256u) {
            if (!s_pre[l]) continue;
            if ((l & 31u) && s_pre[l - 1u] && (!use_threshold || color_distance(s_col[l], s_col[l - 1u]) <= threshold)) suf_union(s_lab, l, l - 1u);
            if ((l >> 5) && s_pre[l - CCL_T] && (!use_threshold || color_distance(s_col[l], s_col[l - CCL_T]) <= threshold)) suf_union(s_lab, l, l - CCL_T);
        }
        __syncthreads();
    }
    for (uint32_t l = threadIdx.x; l < CCL_T * CCL_T; l += 256u) {
        const uint32_t x = x0 + (l & 31u), y = y0 + (l >> 5);
        if (x >= cw || y >= ch) continue;
        const size_t g = (size_t) y * cw + x;
        if (!s_pre[l]) { lab[g] = 0xffffffffu; continue; }
        const uint32_t r = suf_find(s_lab, l);
        lab[g] = (uint32_t) ((size_t) (y0 + (r >> 5)) * cw + x0 + (r & 31u));
    }
}

// global unions across tile borders: pixels in the first column / row of a tile with their left / upper neighbour
__global__ void __launch_bounds__(256)
k_ccl_border(const uint8_t *__restrict__ present, const uint32_t *__restrict__ stored, uint32_t *__restrict__ lab, uint32_t cw, uint32_t ch,
             int use_threshold, double threshold) {
    // thread -> (tile row or column line, position along it): first the vertical border lines (x % 32 == 0, x > 0), then the horizontal ones
    const uint32_t nvx = (cw - 1u) / CCL_T, nhy = (ch - 1u) / CCL_T;               // number of interior border lines
    const size_t nv = (size_t) nvx * ch, nh = (size_t) nhy * cw;
    const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nv) {
        const uint32_t x = (uint32_t) (t / ch + 1u) * CCL_T, y = (uint32_t) (t % ch);
        const size_t g = (size_t) y * cw + x;
        if (present[g] && present[g - 1] && (!use_threshold || color_distance(stored[g], stored[g - 1]) <= threshold)) uf_union(lab, (uint32_t) g, (uint32_t) g - 1u);
    } else if (t < nv + nh) {
        const size_t u = t - nv;
        const uint32_t y = (uint32_t) (u / cw + 1u) * CCL_T, x = (uint32_t) (u % cw);
        const size_t g = (size_t) y * cw + x;
        if (present[g] && present[g - cw] && (!use_threshold || color_distance(stored[g], stored[g - cw]) <= threshold)) uf_union(lab, (uint32_t) g, (uint32_t) (g - cw));
    }
}

// ---------------------------------------------------------------------------------------- agglomerative regime
// blob_number > 1, blob_threshold < 1, blob_max_size or blob_min_size set: the reference grows blobs one merge at a time
// in an RNG order (thread.cpp:288-404: the blob at the back of a shuffled vector merges with the neighbour whose MEAN
// colour is nearest, if within the threshold; undersized blobs ignore the threshold; blobs at blob_max_size stop) until
// blob_number blobs are left.  Its result depends on the RNG order, so it cannot be reproduced bit for bit;
// what is kept here is the rule set, applied in parallel rounds: every blob proposes its nearest-colour neighbour, MUTUAL
// proposals merge (disjoint pairs, so a round is conflict-free), means are size-weighted averages of the two means
// exactly as thread.cpp:348-357, distances are taken between the means rounded to 8 bits (thread.cpp:324-326), and the
// last round applies only as many merges (smallest distance first) as are needed to land on blob_number exactly.
struct Agg {
    uint32_t *lab;          // parent per canvas position (root = its own index)
    double   *mean;         // [4][canvas] mean colour of the blob rooted at a position (0..255 scale)
    uint32_t *size;         // [canvas]
    uint32_t *col;          // [canvas] mean colour rounded to 8 bits per channel
    unsigned long long *best;   // [canvas] (squared colour distance << 46 | tie-break hash << 32 | neighbour root), ~0 = none
    size_t    n;
};

__global__ void __launch_bounds__(256)
k_agg_init(const uint8_t *__restrict__ present, const uint32_t *__restrict__ stored, Agg g) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    g.best[i] = ~0ull;
    if (!present[i]) { g.lab[i] = 0xffffffffu; g.size[i] = 0u; return; }
    const uint32_t c = stored[i];
    g.lab[i] = (uint32_t) i; g.size[i] = 1u; g.col[i] = c;
    g.mean[i] = (double) c_r(c); g.mean[g.n + i] = (double) c_g(c); g.mean[2 * g.n + i] = (double) c_b(c); g.mean[3 * g.n + i] = (double) c_a(c);
}

// every pair of 4-adjacent pixels in different blobs: both blobs consider each other
__global__ void __launch_bounds__(256)
k_agg_best(Agg g, uint32_t cw, uint32_t ch, double threshold, uint32_t max_size, uint32_t min_size, uint32_t round) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const uint32_t a = g.lab[i];
    if (a == 0xffffffffu) return;
    const uint32_t x = (uint32_t) (i % cw), y = (uint32_t) (i / cw);
    const uint32_t sa = g.size[a], ca = g.col[a];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k == 0 ? x + 1 >= cw : y + 1 >= ch) continue;
        const uint32_t b = g.lab[k == 0 ? i + 1 : i + cw];
        if (b == 0xffffffffu || b == a) continue;
        const uint32_t sb = g.size[b], cb = g.col[b];
        if (sa >= max_size || sb >= max_size) continue;                              // thread.cpp:304: a blob at the cap stops growing
        const int rd = (int) c_r(ca) - (int) c_r(cb), gd = (int) c_g(ca) - (int) c_g(cb), bd = (int) c_b(ca) - (int) c_b(cb), ad = (int) c_a(ca) - (int) c_a(cb);
        const uint32_t sq = (uint32_t) (rd * rd + gd * gd + bd * bd + ad * ad);
        const bool within = sqrt((double) sq) / 510.0 <= threshold;                  // color_distance, color.h:17-24
        if (!within && sa >= min_size && sb >= min_size) continue;                   // thread.cpp:382-387: undersized blobs ignore the threshold
        // key: squared distance (18 bits) | per-round hash of the neighbour (14 bits: equal distances -- flat regions -- are
        // broken in a different random order every round, so about a third of the blobs find a mutual partner) | neighbour
        const unsigned long long d = (unsigned long long) sq << 46;
        atomicMin(&g.best[a], d | ((unsigned long long) (mix64(((unsigned long long) round << 32) | b) >> 50) << 32) | b);
        atomicMin(&g.best[b], d | ((unsigned long long) (mix64(((unsigned long long) round << 32) | a) >> 50) << 32) | a);
    }
}

// mutual proposals -> pair list (the smaller root of a pair reports it)
__global__ void __launch_bounds__(256)
k_agg_pairs(Agg g, unsigned long long *__restrict__ pairs, uint32_t *__restrict__ count) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n || g.lab[i] != (uint32_t) i) return;
    const unsigned long long ba = g.best[i];
    if (ba == ~0ull) return;
    const uint32_t b = (uint32_t) ba;
    if (b > (uint32_t) i && (uint32_t) g.best[b] == (uint32_t) i) {
        const uint32_t k = atomicAdd(count, 1u);
        pairs[k] = (ba & 0xffffffff00000000ull) | (uint32_t) i;                      // (distance, smaller root): the partner is best[i]
    }
}

// merge the first `m` pairs of the (sorted) list: the larger root hangs under the smaller one
__global__ void __launch_bounds__(256)
k_agg_merge(Agg g, const unsigned long long *__restrict__ pairs, uint32_t m) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t a = (uint32_t) pairs[k], b = (uint32_t) g.best[a];
    const double s1 = (double) g.size[a], s2 = (double) g.size[b];
    const double weight = s2 / (s1 + s2);                                            // thread.cpp:348-357
    double ch4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { ch4[c] = (1.0 - weight) * g.mean[c * g.n + a] + weight * g.mean[c * g.n + b]; g.mean[c * g.n + a] = ch4[c]; }
    g.size[a] = g.size[a] + g.size[b];
    g.col[a] = c_make(to_u8(round(ch4[0])), to_u8(round(ch4[1])), to_u8(round(ch4[2])), to_u8(round(ch4[3])));
    g.lab[b] = a;
}

// point every pixel at its root again and clear the proposals
__global__ void __launch_bounds__(256)
k_agg_flatten(Agg g) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    g.best[i] = ~0ull;
    if (g.lab[i] == 0xffffffffu) return;
    g.lab[i] = uf_find(g.lab, (uint32_t) i);
}

__global__ void __launch_bounds__(256)
k_ccl_compress(uint32_t *__restrict__ lab, size_t n, uint32_t *__restrict__ root_flag) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t l = lab[i];
    if (l == 0xffffffffu) { root_flag[i] = 0; return; }
    uint32_t r = uf_find(lab, (uint32_t) i);
    lab[i] = r;
    root_flag[i] = (r == (uint32_t) i) ? 1u : 0u;
}

// label[i] = rank of the root (blob index); accumulate exact integer statistics per blob
__global__ void __launch_bounds__(256)
k_ccl_stats(const uint32_t *__restrict__ lab, const uint32_t *__restrict__ root_rank, const uint32_t *__restrict__ stored,
            int32_t *__restrict__ label, uint32_t cw, size_t n, unsigned long long *__restrict__ sums /* [nblobs][8] */) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r = 0xffffffffu;
    if (i < n) r = lab[i];
    int32_t b = -1;
    if (r != 0xffffffffu) b = (int32_t) root_rank[r];
    if (i < n) label[i] = b;
    // a block whose pixels all belong to ONE blob (large blobs: most blocks) adds its seven sums through shared memory:
    // seven global atomics per block instead of seven per warp on the same addresses
    __shared__ int s_first;
    __shared__ unsigned long long s_sum[7];
    if (threadIdx.x == 0) s_first = -2;
    if (threadIdx.x < 7) s_sum[threadIdx.x] = 0ull;
    __syncthreads();
    if (b >= 0) s_first = b;                                     // any writer wins: only used when all agree
    __syncthreads();
    const int first = s_first;
    const bool uniform = __syncthreads_and(b < 0 || b == first) != 0;
    unsigned active = __ballot_sync(0xffffffffu, b >= 0);
    unsigned lane = threadIdx.x & 31;
    if (b >= 0) {
        unsigned peers = __match_any_sync(active, b);
        unsigned leader = __ffs(peers) - 1;
        uint32_t c = stored[i];
        unsigned long long v[7] = {1ull, (unsigned long long) (i % cw), (unsigned long long) (i / cw), c_r(c), c_g(c), c_b(c), c_a(c)};
        // reduce within the peer group: one REDUX per sum when the whole warp is one blob (the usual case; every value is
        // < 2^16, so 32 of them fit 32 bits), a loop over the set bits otherwise
        const bool whole = peers == 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            unsigned long long s = 0;
            if (whole) s = __reduce_add_sync(0xffffffffu, (unsigned) v[k]);
            else {
                unsigned m = peers;
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    s += __shfl_sync(peers, v[k], src);
                }
            }
            if (lane == leader) {
                if (uniform) atomicAdd(&s_sum[k], s);
                else atomicAdd(&sums[(size_t) b * 8 + k], s);
            }
        }
    }
    __syncthreads();
    if (uniform && first >= 0 && threadIdx.x < 7 && s_sum[threadIdx.x]) atomicAdd(&sums[(size_t) first * 8 + threadIdx.x], s_sum[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
k_iota_keys(const int32_t *__restrict__ label, unsigned long long *__restrict__ keys, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t b = label[i];
    keys[i] = b < 0 ? 0xffffffffffffffffull : (((unsigned long long) (uint32_t) b << 32) | (unsigned long long) i);
}
__global__ void __launch_bounds__(256)
k_keys_to_pix(const unsigned long long *__restrict__ keys, uint32_t *__restrict__ pix, size_t npix) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) pix[i] = (uint32_t) (keys[i] & 0xffffffffull);
}

// per-blob pixel lists (ascending position inside a blob) from the label image: radix sort of (blob, index) keys
int engine_build_blob_pixels(Engine *E, uint32_t index) {
    FrameDev &f = E->frames[index];
    size_t n = E->canvas();
    dev_free(f.blob_pix); f.blob_pix = nullptr;
    f.blob_pix_off.assign(f.blobs.size() + 1, 0);
    for (size_t b = 0; b < f.blobs.size(); ++b) f.blob_pix_off[b + 1] = f.blob_pix_off[b] + f.blobs[b].size;
    uint64_t npix = f.blob_pix_off.back();
    if (npix == 0) return AMX_OK;
    unsigned long long *k_in = nullptr, *k_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    if (!dev_alloc(E, (void **) &k_in, n * 8, "sort keys") || !dev_alloc(E, (void **) &k_out, n * 8, "sort keys")) { dev_free(k_in); return AMX_ERR_NOMEM; }
    k_iota_keys<<<div_up(n, 256), 256, 0, E->stream>>>(f.label, k_in, n);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k_in, k_out, (int) n, 0, 64, E->stream);
    if (!dev_alloc(E, &tmp, tmp_bytes, "sort tmp")) { dev_free(k_in); dev_free(k_out); return AMX_ERR_NOMEM; }
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, k_in, k_out, (int) n, 0, 64, E->stream);
    int rc = AMX_OK;
    if (!dev_alloc(E, (void **) &f.blob_pix, npix * 4, "blob_pix")) rc = AMX_ERR_NOMEM;
    else k_keys_to_pix<<<div_up(npix, 256), 256, 0, E->stream>>>(k_out, f.blob_pix, npix);
    E->launches += 3;
    if (E->fail(cudaStreamSynchronize(E->stream), "blob pixels") || E->check("blob pixels")) rc = AMX_ERR_CUDA;
    dev_free(k_in); dev_free(k_out); dev_free(tmp);
    return rc;
}

// labels of the agglomerative regime into `lab` (roots = smallest canvas index of a blob); AMX_OK or an error
static int agglomerate_frame(Engine *E, FrameDev &f, uint32_t *lab) {
    const size_t n = E->canvas();
    Agg g;
    g.lab = lab; g.n = n; g.mean = nullptr; g.size = nullptr; g.col = nullptr; g.best = nullptr;
    unsigned long long *pairs = nullptr, *pairs2 = nullptr;
    uint32_t *d_count = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (unsigned long long *) nullptr, (unsigned long long *) nullptr, (int) (n / 2 + 1), 0, 64, E->stream);
    int rc = AMX_OK;
    if (!dev_alloc(E, (void **) &g.mean, n * 4 * sizeof(double), "agg means") || !dev_alloc(E, (void **) &g.size, n * 4, "agg sizes") ||
        !dev_alloc(E, (void **) &g.col, n * 4, "agg colours") || !dev_alloc(E, (void **) &g.best, n * 8, "agg proposals") ||
        !dev_alloc(E, (void **) &pairs, (n / 2 + 1) * 8, "agg pairs") || !dev_alloc(E, (void **) &pairs2, (n / 2 + 1) * 8, "agg pairs") ||
        !dev_alloc(E, (void **) &d_count, 4, "agg count") || !dev_alloc(E, &tmp, tmp_bytes, "agg sort"))
        rc = AMX_ERR_NOMEM;
    if (rc == AMX_OK) {
        const uint32_t max_size = (uint32_t) std::min<uint64_t>(E->p.blob_max_size, 0xffffffffull);
        const uint32_t min_size = (uint32_t) std::min<uint64_t>(E->p.blob_min_size, 0xffffffffull);
        const uint64_t target = std::max<uint64_t>(E->p.blob_number, 1);
        k_agg_init<<<div_up(n, 256), 256, 0, E->stream>>>(f.present, f.stored, g);
        E->launches++;
        uint64_t blobs = f.pixel_count;
        for (int round = 0; round < 4096 && blobs > target; ++round) {
            cudaMemsetAsync(d_count, 0, 4, E->stream);
            k_agg_best<<<div_up(n, 256), 256, 0, E->stream>>>(g, E->cw, E->ch, E->p.blob_threshold, max_size, min_size, (uint32_t) round + E->p.seed * 7919u);
            k_agg_pairs<<<div_up(n, 256), 256, 0, E->stream>>>(g, pairs, d_count);
            uint32_t npairs = 0;
            cudaMemcpyAsync(&npairs, d_count, 4, cudaMemcpyDeviceToHost, E->stream);
            if (E->fail(cudaStreamSynchronize(E->stream), "agglomerate")) { rc = AMX_ERR_CUDA; break; }
            E->launches += 2;
            if (npairs == 0) break;                                          // nothing can expand any more (thread.cpp:402)
            uint32_t m = npairs;
            const unsigned long long *list = pairs;
            if (blobs - npairs < target) {
                // the last round: only the blobs - target nearest pairs
                m = (uint32_t) (blobs - target);
                cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, pairs, pairs2, (int) npairs, 0, 64, E->stream);
                list = pairs2;
            }
            k_agg_merge<<<div_up(m, 256), 256, 0, E->stream>>>(g, list, m);
            k_agg_flatten<<<div_up(n, 256), 256, 0, E->stream>>>(g);
            E->launches += 2;
            blobs -= m;
        }
        if (rc == AMX_OK && E->check("agglomerate")) rc = AMX_ERR_CUDA;
    }
    dev_free(g.mean); dev_free(g.size); dev_free(g.col); dev_free(g.best); dev_free(pairs); dev_free(pairs2); dev_free(d_count); dev_free(tmp);
    return rc;
}

static int blobify_frame(Engine *E, uint32_t index) {
    FrameDev &f = E->frames[index];
    size_t n = E->canvas();
    f.blobs.clear();
    dev_free(f.blob_pix); f.blob_pix = nullptr; f.blob_pix_off.clear();
    if (f.pixel_count == 0) { cudaMemsetAsync(f.label, 0xff, n * 4, E->stream); return AMX_OK; }
    uint32_t *lab = nullptr, *flag = nullptr, *rank = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    unsigned long long *sums = nullptr;
    int rc = AMX_OK;
    if (!dev_alloc(E, (void **) &lab, n * 4, "ccl lab") || !dev_alloc(E, (void **) &flag, n * 4, "ccl flag") || !dev_alloc(E, (void **) &rank, n * 4, "ccl rank")) rc = AMX_ERR_NOMEM;
    if (rc == AMX_OK) {
        bool use_thr = E->p.blob_threshold < 1.0;
        // non-default regimes (several blobs wanted, a colour threshold, size limits): parallel agglomeration
        const bool agglomerate = E->p.blob_number > 1 || use_thr || E->p.blob_max_size < 0xffffffffull || E->p.blob_min_size > 1;
        if (agglomerate && E->p.blob_max_size > 1) rc = agglomerate_frame(E, f, lab);
        // tile-local labelling in shared memory, then unions across the tile borders only
        const bool merge = E->p.blob_max_size > 1 && !agglomerate;
        if (!agglomerate || E->p.blob_max_size <= 1) {
        const dim3 tgrid(div_up(E->cw, CCL_T), div_up(E->ch, CCL_T));
        k_ccl_tile<<<tgrid, 256, 0, E->stream>>>(f.present, f.stored, lab, E->cw, E->ch, use_thr ? 1 : 0, E->p.blob_threshold, merge ? 1 : 0);
        if (merge) {
            const size_t nborder = (size_t) ((E->cw - 1u) / CCL_T) * E->ch + (size_t) ((E->ch - 1u) / CCL_T) * E->cw;
            if (nborder) k_ccl_border<<<div_up(nborder, 256), 256, 0, E->stream>>>(f.present, f.stored, lab, E->cw, E->ch, use_thr ? 1 : 0, E->p.blob_threshold);
        }
        }
        k_ccl_compress<<<div_up(n, 256), 256, 0, E->stream>>>(lab, n, flag);
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag, rank, (int) n, E->stream);
        if (!dev_alloc(E, &tmp, tmp_bytes, "scan tmp")) rc = AMX_ERR_NOMEM;
    }
    uint32_t nblobs = 0;
    if (rc == AMX_OK) {
        cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, rank, (int) n, E->stream);
        uint32_t last_rank = 0, last_flag = 0;
        cudaMemcpyAsync(&last_rank, rank + n - 1, 4, cudaMemcpyDeviceToHost, E->stream);
        cudaMemcpyAsync(&last_flag, flag + n - 1, 4, cudaMemcpyDeviceToHost, E->stream);
        if (E->fail(cudaStreamSynchronize(E->stream), "ccl")) rc = AMX_ERR_CUDA;
        nblobs = last_rank + last_flag;
        E->launches += 4;
    }
    if (rc == AMX_OK && !dev_alloc(E, (void **) &sums, (size_t) std::max(nblobs, 1u) * 64, "ccl sums")) rc = AMX_ERR_NOMEM;
    if (rc == AMX_OK) {
        cudaMemsetAsync(sums, 0, (size_t) std::max(nblobs, 1u) * 64, E->stream);
        k_ccl_stats<<<div_up(n, 256), 256, 0, E->stream>>>(lab, rank, f.stored, f.label, E->cw, n, sums);
        E->launches++;
        std::vector<unsigned long long> hs((size_t) nblobs * 8);
        if (nblobs) cudaMemcpyAsync(hs.data(), sums, hs.size() * 8, cudaMemcpyDeviceToHost, E->stream);
        if (E->fail(cudaStreamSynchronize(E->stream), "ccl stats") || E->check("ccl stats")) rc = AMX_ERR_CUDA;
        else {
            f.blobs.resize(nblobs);
            for (uint32_t b = 0; b < nblobs; ++b) {
                const unsigned long long *s = &hs[(size_t) b * 8];
                BlobHost &bl = f.blobs[b];
                double cnt = (double) s[0];
                bl.size = s[0];
                bl.group = b;
                bl.stats[0] = (double) s[1] / cnt;
                bl.stats[1] = (double) s[2] / cnt;
                for (int k = 0; k < 4; ++k) bl.stats[2 + k] = ((double) s[3 + k] / cnt) / 255.0;
            }
        }
    }
    dev_free(lab); dev_free(flag); dev_free(rank); dev_free(tmp); dev_free(sums);
    if (rc == AMX_OK) rc = engine_build_blob_pixels(E, index);
    return rc;
}

// ---------------------------------------------------------------------------------------- unification (row a-B2)
// thread::unify_frame (thread.cpp:426-585): blobs that never reached blob_min_size ("dust") are clustered among themselves:
// a dust blob looks at blob_box_samples other dust blobs (the next ones in the reference's shuffled vector, i.e. random
// ones), takes the nearest whose rounded centroid lies within blob_box_grip in both axes and absorbs it (size-weighted
// means; the merged blob need not be connected -- "a pixel dust cloud"), until nothing merges any more or blob_number
// blobs are left.  Only runs when there are more blobs than blob_number and blob_box_samples > 0.  The sampling order is
// RNG-dependent in the reference; here the samples come from the counter RNG (statistical parity).
__global__ void __launch_bounds__(256)
k_relabel(int32_t *__restrict__ label, const int32_t *__restrict__ remap, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t b = label[i];
    if (b >= 0) label[i] = remap[b];
}

static int unify_frame(Engine *E, uint32_t index) {
    FrameDev &f = E->frames[index];
    const uint64_t min_size = E->p.blob_min_size, samples = E->p.blob_box_samples, target = std::max<uint64_t>(E->p.blob_number, 1);
    const size_t nb = f.blobs.size();
    if (min_size <= 1 || samples == 0 || nb <= target) return AMX_OK;
    std::vector<uint32_t> dust;
    for (size_t b = 0; b < nb; ++b) if (f.blobs[b].size > 0 && f.blobs[b].size < min_size) dust.push_back((uint32_t) b);
    if (dust.size() < 2) return AMX_OK;
    std::vector<int32_t> into(nb, -1);                  // absorbed blob -> absorbing blob
    std::vector<BlobHost> bl = f.blobs;
    uint64_t count = nb, draw = 0;
    const uint64_t grip = E->p.blob_box_grip;
    bool merged_any = false;
    for (int pass = 0; pass < 64 && count > target; ++pass) {
        bool merged = false;
        for (size_t di = 0; di < dust.size() && count > target; ++di) {
            const uint32_t u = dust[di];
            if (into[u] >= 0) continue;
            const long ux = std::lround(bl[u].stats[0]), uy = std::lround(bl[u].stats[1]);
            int64_t best = -1;
            uint64_t best_score = 0;
            for (uint64_t k = 0; k < samples; ++k) {
                const uint32_t v = dust[rng64(E->p.seed, 0xd057u + index, draw++) % dust.size()];
                if (v == u || into[v] >= 0) continue;
                const long vx = std::lround(bl[v].stats[0]), vy = std::lround(bl[v].stats[1]);
                const uint64_t dx = (uint64_t) std::labs(ux - vx), dy = (uint64_t) std::labs(uy - vy);
                if (dx > grip || dy > grip) continue;
                const uint64_t score = dx * dx + dy * dy;
                if (best < 0 || score < best_score) { best = v; best_score = score; }
            }
            if (best < 0) continue;
            const double s1 = (double) bl[u].size, s2 = (double) bl[best].size, weight = s2 / (s1 + s2);     // thread.cpp:488-505
            for (int k = 0; k < 6; ++k) bl[u].stats[k] = (1.0 - weight) * bl[u].stats[k] + weight * bl[best].stats[k];
            bl[u].size += bl[best].size;
            into[best] = (int32_t) u;
            --count;
            merged = merged_any = true;
        }
        if (!merged) break;
    }
    if (!merged_any) return AMX_OK;
    // compact the survivors (blob vector order kept) and relabel the pixels
    std::vector<int32_t> remap(nb, -1);
    std::vector<BlobHost> out;
    for (size_t b = 0; b < nb; ++b) if (into[b] < 0) { remap[b] = (int32_t) out.size(); BlobHost x = bl[b]; x.group = out.size(); out.push_back(x); }
    for (size_t b = 0; b < nb; ++b) if (into[b] >= 0) { int32_t r = into[b]; while (into[r] >= 0) r = into[r]; remap[b] = remap[r]; }
    int32_t *d_remap = nullptr;
    if (!dev_alloc(E, (void **) &d_remap, nb * 4, "blob remap")) return AMX_ERR_NOMEM;
    cudaMemcpyAsync(d_remap, remap.data(), nb * 4, cudaMemcpyHostToDevice, E->stream);
    k_relabel<<<div_up(E->canvas(), 256), 256, 0, E->stream>>>(f.label, d_remap, E->canvas());
    E->launches++;
    bool bad = E->fail(cudaStreamSynchronize(E->stream), "unify") || E->check("unify");
    dev_free(d_remap);
    if (bad) return AMX_ERR_CUDA;
    f.blobs = out;
    return engine_build_blob_pixels(E, index);
}

int engine_unify(Engine *E) {
    for (uint32_t i = 0; i < E->frames.size(); ++i) {
        int rc = unify_frame(E, i);
        if (rc != AMX_OK) return rc;
    }
    return AMX_OK;
}

int engine_blobify(Engine *E) {
    for (uint32_t i = 0; i < E->frames.size(); ++i) {
        int rc = blobify_frame(E, i);
        if (rc != AMX_OK) return rc;
    }
    return AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_blobify(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    int rc = engine_blobify(&ctx->e);
    if (rc == AMX_OK && ctx->e.state == ST_BLOB_DETECTION) ctx->e.state = ST_BLOB_UNIFICATION;
    return rc;
}

int amx_blob_count(amx_ctx *ctx, uint32_t index, uint32_t *count) {
    if (!ctx || !count || index >= ctx->e.frames.size()) return AMX_ERR_ARG;
    *count = (uint32_t) ctx->e.frames[index].blobs.size();
    return AMX_OK;
}

int amx_export_blobs(amx_ctx *ctx, uint32_t index, int32_t *labels_out, double *stats_out, uint64_t *meta_out) {
    if (!ctx || index >= ctx->e.frames.size()) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    FrameDev &f = E->frames[index];
    if (labels_out) {
        if (E->fail(cudaMemcpyAsync(labels_out, f.label, E->canvas() * 4, cudaMemcpyDeviceToHost, E->stream), "labels D2H") ||
            E->fail(cudaStreamSynchronize(E->stream), "labels"))
            return AMX_ERR_CUDA;
    }
    for (size_t b = 0; b < f.blobs.size(); ++b) {
        if (stats_out) for (int k = 0; k < 6; ++k) stats_out[6 * b + k] = f.blobs[b].stats[k];
        if (meta_out) { meta_out[2 * b] = f.blobs[b].group; meta_out[2 * b + 1] = f.blobs[b].size; }
    }
    return AMX_OK;
}

int amx_import_blobs(amx_ctx *ctx, uint32_t index, uint32_t nblobs, const int32_t *labels, const double *stats, const uint64_t *groups) {
    if (!ctx || index >= ctx->e.frames.size() || (nblobs && (!stats || !groups))) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    FrameDev &f = E->frames[index];
    size_t n = E->canvas();
    f.blobs.assign(nblobs, BlobHost());
    for (uint32_t b = 0; b < nblobs; ++b) {
        for (int k = 0; k < 6; ++k) f.blobs[b].stats[k] = stats[6 * b + k];
        f.blobs[b].group = groups[b];
        f.blobs[b].size = 0;
    }
    if (labels) {
        for (size_t i = 0; i < n; ++i) if (labels[i] >= 0 && (uint32_t) labels[i] < nblobs) f.blobs[labels[i]].size++;
        if (E->fail(cudaMemcpyAsync(f.label, labels, n * 4, cudaMemcpyHostToDevice, E->stream), "labels H2D") ||
            E->fail(cudaStreamSynchronize(E->stream), "labels"))
            return AMX_ERR_CUDA;
    } else cudaMemsetAsync(f.label, 0xff, n * 4, E->stream);
    dev_free(f.blob_pix); f.blob_pix = nullptr; f.blob_pix_off.clear();
    E->render_ready = false;
    return labels ? engine_build_blob_pixels(E, index) : AMX_OK;
}

}
