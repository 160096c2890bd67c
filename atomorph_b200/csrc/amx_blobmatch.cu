/*
 * amx_blobmatch.cu -- K3: blob matching across key frames (SURVEY.md row a-B3).
 *
 * Reference: thread::match (thread.cpp:598-738): every frame is padded with empty ("volatile")
 * blobs to the largest blob count, blob_map[slot][frame] assigns blobs to groups (= slots), and
 * one random pair of slots of one random frame is swapped per step when the summed
 * blob_distance (thread.cpp:1151-1174) to the neighbouring frames does not increase.
 *
 * Here one ROUND proposes a perfect matching of the slots of one frame (slot l with l XOR m,
 * counter-RNG mask m), all proposals of a round evaluated in parallel -- the same disjoint-pair
 * scheme as the atom matcher (amx_swap.cu).  Costs are the reference's double formula on the
 * same truncated/rounded blob features.  The search is pure descent (c1 >= c2 accepted) run as
 * 296 independent REPLICAS (one CTA each, different proposal streams, same start); the lowest
 * energy replica wins.  The reference's periodic forced uphill move ("degeneration") is not
 * reproduced; best-of-replicas plays its role of escaping poor 2-swap local optima, and the
 * best-so-far bookkeeping coincides with the current map.
 */
#include <algorithm>
#include <cmath>
#include "amx_engine.h"

namespace amx {

// feature record: 0 size 1 x(trunc u16) 2 y(trunc u16) 3 packed colour (as double bits of u32) -- 4 doubles
struct BW { double xy, rgba, size; double bbox_d; };

__device__ __forceinline__ double blob_dist(const double *a, const double *b, const BW &w) {
    double sz1 = a[0], sz2 = b[0], szs = sz1 + sz2;
    double pix = 0.0, col = 0.0, siz = 0.0;
    if (szs > 0.0) siz = fabs(sz1 - sz2) / szs;
    if (sz1 > 0.0 && sz2 > 0.0) {
        double xd = a[1] - b[1], yd = a[2] - b[2];
        pix = sqrt((xd * xd + yd * yd) / w.bbox_d);
        col = color_distance((uint32_t) a[3], (uint32_t) b[3]);
    }
    return (w.xy * pix + w.rgba * col + w.size * siz);
}

// shared by the round kernel: evaluate the swap of slots l, r of frame y on map `bmap`
__device__ __forceinline__ void match_propose(uint32_t *bmap, const double *__restrict__ feat, uint32_t W, uint32_t H, uint32_t y,
                                              uint32_t l, uint32_t r, const BW &w) {
    uint32_t yn = (y + 1) % H, yp = (y + H - 1) % H;
    const double *fy = feat + (size_t) y * W * 4, *fn = feat + (size_t) yn * W * 4, *fp = feat + (size_t) yp * W * 4;
    uint32_t b1 = bmap[(size_t) y * W + l], b2 = bmap[(size_t) y * W + r];
    const double *x1 = fy + 4 * (size_t) b1, *x2 = fy + 4 * (size_t) b2;
    if (x1[0] == 0.0 && x2[0] == 0.0) return;                      // both volatile: thread.cpp:692-694
    const double *x1p = fp + 4 * (size_t) bmap[(size_t) yp * W + l], *x2p = fp + 4 * (size_t) bmap[(size_t) yp * W + r];
    const double *x1n = fn + 4 * (size_t) bmap[(size_t) yn * W + l], *x2n = fn + 4 * (size_t) bmap[(size_t) yn * W + r];
    double x1_before = blob_dist(x1p, x1, w) + blob_dist(x1, x1n, w);
    double x2_before = blob_dist(x2p, x2, w) + blob_dist(x2, x2n, w);
    double x1_after = blob_dist(x2p, x1, w) + blob_dist(x1, x2n, w);
    double x2_after = blob_dist(x1p, x2, w) + blob_dist(x2, x1n, w);
    double c1 = x1_before + x2_before, c2 = x1_after + x2_after;
    if (c1 >= c2) {
        bmap[(size_t) y * W + l] = b2;
        bmap[(size_t) y * W + r] = b1;
    }
}

// One CTA = one REPLICA of the descent: all replicas start from the current map and follow different
// counter-RNG proposal streams for `rounds` rounds (block barrier between rounds); the host keeps the
// replica with the lowest energy.  The reference runs a single serial descent; best-of-R is the
// massively parallel way to spend the same wall time.
__global__ void __launch_bounds__(256)
k_match_replicas(const uint32_t *__restrict__ bmap0, uint32_t *__restrict__ rmaps, const double *__restrict__ feat, uint32_t W, uint32_t H,
                 unsigned kbits, uint64_t seed, uint64_t round0, uint64_t rounds, BW w, double *__restrict__ renergy) {
    __shared__ double sh[8];
    uint32_t *bmap = rmaps + (size_t) blockIdx.x * W * H;
    for (uint32_t i = threadIdx.x; i < W * H; i += blockDim.x) bmap[i] = bmap0[i];
    __syncthreads();
    for (uint64_t rd = 0; rd < rounds; ++rd) {
        uint64_t rr = rng64(seed ^ (0x9e3779b97f4a7c15ull * (blockIdx.x + 1)), 0xb10bu, round0 + rd);
        uint32_t y = (uint32_t) (rr % H);
        uint32_t m = 1u + (uint32_t) ((rr >> 20) % ((1ull << kbits) - 1ull));
        for (uint32_t l = threadIdx.x; l < W; l += blockDim.x) {
            uint32_t r = l ^ m;
            if (r > l && r < W) match_propose(bmap, feat, W, H, y, l, r, w);
        }
        __syncthreads();
    }
    double e = 0.0;
    for (uint32_t i = threadIdx.x; i < W; i += blockDim.x) {
        const double *prev = feat + ((size_t) (H - 1) * W + bmap[(size_t) (H - 1) * W + i]) * 4;
        for (uint32_t j = 0; j < H; ++j) {
            const double *cur = feat + ((size_t) j * W + bmap[(size_t) j * W + i]) * 4;
            e += blob_dist(prev, cur, w);
            prev = cur;
        }
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int k = 0; k < 8; ++k) t += sh[k]; renergy[blockIdx.x] = t; }
}

// energy = sum over slots and frames of dist(map[i][j-1], map[i][j]) (thread.cpp:1087-1107), one partial per block
__global__ void __launch_bounds__(256)
k_match_energy(const uint32_t *__restrict__ bmap, const double *__restrict__ feat, uint32_t W, uint32_t H, BW w, double *__restrict__ out) {
    __shared__ double sh[8];
    double e = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < W; i += gridDim.x * blockDim.x) {
        const double *prev = feat + ((size_t) (H - 1) * W + bmap[(size_t) (H - 1) * W + i]) * 4;
        for (uint32_t j = 0; j < H; ++j) {
            const double *cur = feat + ((size_t) j * W + bmap[(size_t) j * W + i]) * 4;
            e += blob_dist(prev, cur, w);
            prev = cur;
        }
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int k = 0; k < 8; ++k) t += sh[k]; out[blockIdx.x] = t; }
}

static BW make_weights(Engine *E) {
    BW w;
    // thread.cpp:1141-1149 (defaults 0.33/0.33/0.34 are kept when all weights are 0)
    double rg = E->p.blob_rgba_weight, sz = E->p.blob_size_weight, xy = E->p.blob_xy_weight;
    double sum = rg * rg + sz * sz + xy * xy;
    if (sum == 0.0) { w.rgba = 0.33; w.size = 0.33; w.xy = 0.34; }
    else { w.rgba = (rg * rg) / sum; w.size = (sz * sz) / sum; w.xy = 1.0 - (w.rgba + w.size); }
    int dx = (int) E->bbox[2] - (int) E->bbox[0], dy = (int) E->bbox[3] - (int) E->bbox[1];
    w.bbox_d = (E->bbox[0] > E->bbox[2] || E->bbox[1] > E->bbox[3]) ? 0.0 : (double) (uint32_t) (dx * dx + dy * dy);   // thread.cpp:1127-1139
    return w;
}

static int upload_map(Engine *E) {
    uint32_t W = E->map_w, H = E->map_h;
    std::vector<double> feat((size_t) W * H * 4);
    for (uint32_t f = 0; f < H; ++f)
        for (uint32_t b = 0; b < W; ++b) {
            const BlobHost &bl = E->frames[f].blobs[b];
            double *o = &feat[((size_t) f * W + b) * 4];
            o[0] = (double) bl.size;
            o[1] = (double) ((uint32_t) ((int32_t) bl.stats[0]) & 0xffffu);    // create_pixel(double x ...) truncates to u16, thread.cpp:1165
            o[2] = (double) ((uint32_t) ((int32_t) bl.stats[1]) & 0xffffu);
            o[3] = (double) create_color_d(bl.stats[2], bl.stats[3], bl.stats[4], bl.stats[5]);
        }
    dev_free(E->d_bfeat); dev_free(E->d_bmap); dev_free(E->d_menergy);
    if (!dev_alloc(E, (void **) &E->d_bfeat, feat.size() * 8, "blob features") || !dev_alloc(E, (void **) &E->d_bmap, (size_t) W * H * 4, "blob map") ||
        !dev_alloc(E, (void **) &E->d_menergy, 64 * 8 + 8, "match energy"))
        return AMX_ERR_NOMEM;
    cudaMemcpyAsync(E->d_bfeat, feat.data(), feat.size() * 8, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(E->d_bmap, E->blob_map.data(), (size_t) W * H * 4, cudaMemcpyHostToDevice, E->stream);
    return E->fail(cudaStreamSynchronize(E->stream), "match upload") ? AMX_ERR_CUDA : AMX_OK;
}

static void groups_from_map(Engine *E) {
    for (uint32_t f = 0; f < E->map_h; ++f)
        for (uint32_t x = 0; x < E->map_w; ++x) E->frames[f].blobs[E->blob_map[(size_t) f * E->map_w + x]].group = x;
}

int engine_match_energy(Engine *E, double *e) {
    *e = 0.0;
    if (!E->map_ready || E->map_w == 0 || E->map_h == 0) return AMX_OK;
    BW w = make_weights(E);
    uint32_t nb = std::min<uint32_t>(64, div_up(E->map_w, 256));
    k_match_energy<<<nb, 256, 0, E->stream>>>(E->d_bmap, E->d_bfeat, E->map_w, E->map_h, w, E->d_menergy);
    E->launches++;
    double host[64];
    if (E->fail(cudaMemcpyAsync(host, E->d_menergy, nb * 8, cudaMemcpyDeviceToHost, E->stream), "energy D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "energy"))
        return AMX_ERR_CUDA;
    for (uint32_t i = 0; i < nb; ++i) *e += host[i];
    E->blob_map_e = *e;
    return AMX_OK;
}

int engine_match_init(Engine *E) {
    size_t nf = E->frames.size();
    uint32_t W = 0;
    for (auto &f : E->frames) W = std::max<uint32_t>(W, (uint32_t) f.blobs.size());
    E->map_w = W; E->map_h = (uint32_t) nf;
    for (auto &f : E->frames) {
        while (f.blobs.size() < W) {          // volatile padding, thread.cpp:636-647
            BlobHost b;
            for (int k = 0; k < 6; ++k) b.stats[k] = f.means[k];
            b.size = 0; b.group = f.blobs.size();
            f.blobs.push_back(b);
        }
        f.blob_pix_off.resize(W + 1, f.blob_pix_off.empty() ? 0 : f.blob_pix_off.back());
    }
    E->blob_map.assign((size_t) W * nf, 0);
    for (size_t f = 0; f < nf; ++f)
        for (uint32_t x = 0; x < W; ++x) E->blob_map[f * W + x] = x;
    E->map_ready = true;
    groups_from_map(E);
    if (W == 0 || nf == 0) { E->blob_map_e = 0.0; return AMX_OK; }
    int rc = upload_map(E);
    if (rc != AMX_OK) return rc;
    double e;
    return engine_match_energy(E, &e);
}

int engine_match_rounds(Engine *E, uint64_t rounds) {
    if (!E->map_ready) { int rc = engine_match_init(E); if (rc != AMX_OK) return rc; }
    uint32_t W = E->map_w, H = E->map_h;
    if (H < 2 || W <= 1 || rounds == 0) return AMX_OK;
    BW w = make_weights(E);
    unsigned k = 0;
    while ((1u << k) < W) ++k;
    const uint32_t R = 296;                       // 2 replicas per SM
    uint32_t *d_rmaps = nullptr;
    double *d_re = nullptr;
    if (!dev_alloc(E, (void **) &d_rmaps, (size_t) R * W * H * 4, "replica maps") || !dev_alloc(E, (void **) &d_re, R * 8, "replica energy")) {
        dev_free(d_rmaps);
        return AMX_ERR_NOMEM;
    }
    k_match_replicas<<<R, 256, 0, E->stream>>>(E->d_bmap, d_rmaps, E->d_bfeat, W, H, k, E->p.seed, E->rng_round, rounds, w, d_re);
    E->rng_round += rounds;
    E->launches++;
    std::vector<double> re(R);
    int rc = AMX_OK;
    if (E->fail(cudaMemcpyAsync(re.data(), d_re, R * 8, cudaMemcpyDeviceToHost, E->stream), "replica energy D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "match rounds") || E->check("match rounds"))
        rc = AMX_ERR_CUDA;
    if (rc == AMX_OK) {
        uint32_t best = 0;
        for (uint32_t i = 1; i < R; ++i) if (re[i] < re[best]) best = i;
        if (E->fail(cudaMemcpyAsync(E->d_bmap, d_rmaps + (size_t) best * W * H, (size_t) W * H * 4, cudaMemcpyDeviceToDevice, E->stream), "best map") ||
            E->fail(cudaMemcpyAsync(E->blob_map.data(), E->d_bmap, (size_t) W * H * 4, cudaMemcpyDeviceToHost, E->stream), "map D2H") ||
            E->fail(cudaStreamSynchronize(E->stream), "match rounds"))
            rc = AMX_ERR_CUDA;
    }
    dev_free(d_rmaps); dev_free(d_re);
    if (rc != AMX_OK) return rc;
    groups_from_map(E);
    double e;
    return engine_match_energy(E, &e);
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_match_init(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_match_init(&ctx->e);
}
int amx_match_rounds(amx_ctx *ctx, uint64_t rounds) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_match_rounds(&ctx->e, rounds);
}
int amx_match_energy(amx_ctx *ctx, double *energy) {
    if (!ctx || !energy) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_match_energy(&ctx->e, energy);
}

}
