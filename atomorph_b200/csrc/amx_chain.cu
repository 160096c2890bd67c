/*
 * amx_chain.cu -- K4: chain / atom table construction (SURVEY.md row a-C) and table import/export.
 *
 * Reference: thread::init_morph (thread.cpp:740-891), renew_chain (atomorph.cpp:1085-1157),
 * fix_volatiles (thread.cpp:1187-1233).  Per blob group i: width = max_j |surface| * density;
 * column j = the blob's pixels in ascending position order (flags HAS_PIXEL|HAS_FLUID, fract 0);
 * an empty blob contributes one volatile point at its (interpolated) centroid (HAS_FLUID only);
 * the remaining rows repeat a SHUFFLED copy of those points with random sub-pixel offsets and
 * HAS_FLUID cleared.  The reference shuffles with std::shuffle(minstd_rand0); here the
 * duplicate order and the fract bytes come from the counter-based RNG (statistically the same
 * construction; columns without duplicates are bit-identical to the reference's).
 *
 * Device layout: one table for all chains, column-major: table[j*A + off_c + x].
 */
#include <algorithm>
#include <cmath>
#include "amx_engine.h"

namespace amx {

int engine_alloc_chains(Engine *E, uint32_t nchains, const uint64_t *keys, const uint64_t *widths, const uint64_t *max_surface, uint32_t height) {
    cudaStreamSynchronize(E->stream);
    std::vector<uint64_t> offs(nchains + 1, 0);
    for (uint32_t c = 0; c < nchains; ++c) offs[c + 1] = offs[c] + widths[c];
    // same geometry as before (a table re-import): every device buffer, the render buffers included, is kept
    const bool same = E->table && E->chain_of && E->d_chain_off && E->nchains == nchains && E->h == height && E->chain_off == offs;
    if (!same) {
        engine_dist_table_gone(E);
        dev_free(E->table); dev_free(E->chain_of); dev_free(E->d_chain_off);
        E->table = nullptr; E->chain_of = nullptr; E->d_chain_off = nullptr;
        engine_render_free(E);
    }
    E->render_ready = false;
    E->nchains = nchains; E->h = height;
    E->chain_key.assign(keys, keys + nchains);
    E->chain_max_surface.assign(max_surface, max_surface + nchains);
    E->chain_off.assign(nchains + 1, 0);
    for (uint32_t c = 0; c < nchains; ++c) E->chain_off[c + 1] = E->chain_off[c] + widths[c];
    E->A = E->chain_off[nchains];
    if (!same && (!dev_alloc(E, (void **) &E->table, (size_t) height * E->A * 8, "chain table") ||
                  !dev_alloc(E, (void **) &E->chain_of, E->A * 4, "chain_of") ||
                  !dev_alloc(E, (void **) &E->d_chain_off, (size_t) (nchains + 1) * 8, "chain_off")))
        return AMX_ERR_NOMEM;
    std::vector<uint32_t> cof(same ? 0 : E->A);
    for (uint32_t c = 0; c < nchains && !same; ++c) std::fill(cof.begin() + E->chain_off[c], cof.begin() + E->chain_off[c + 1], c);
    if ((!same && E->fail(cudaMemcpyAsync(E->chain_of, cof.data(), E->A * 4, cudaMemcpyHostToDevice, E->stream), "chain_of H2D")) ||
        E->fail(cudaMemcpyAsync(E->d_chain_off, E->chain_off.data(), (size_t) (nchains + 1) * 8, cudaMemcpyHostToDevice, E->stream), "chain_off H2D") ||
        E->fail(cudaStreamSynchronize(E->stream), "alloc chains"))
        return AMX_ERR_CUDA;
    cudaMemsetAsync(E->d_swapstats, 0, 24, E->stream);
    E->swapstats[0] = E->swapstats[1] = E->swapstats[2] = 0;
    return AMX_OK;
}

// random bijection of [0, n) by cycle-walking a 4-round Feistel network over 2^(2*hb) >= n
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t n, uint64_t seed) {
    unsigned bits = 64 - __clzll(n | 1ull);
    unsigned hb = (bits + 1) / 2;
    if (hb == 0) hb = 1;
    uint64_t mask = (1ull << hb) - 1ull;
    uint64_t x = i;
    do {
        uint64_t L = x >> hb, R = x & mask;
        for (int r = 0; r < 4; ++r) {
            uint64_t F = mix64(R ^ (seed + 0x9e37u * (uint64_t) r)) & mask;
            uint64_t nl = R;
            R = L ^ F;
            L = nl;
        }
        x = (L << hb) | R;
    } while (x >= n);
    return x;
}

// fill column j of every chain (thread.cpp:793-848).
// All chains of one key frame in ONE launch (thread = atom of the concatenated chains): a scene with thousands of blob groups
// (BASELINE config 4: 2 500 chains x 2 key frames) spent 28 ms launching k_fill_column 5 000 times for 5.6 us each.
struct FillDesc { uint32_t pix_off, npix, vx, vy; };      // per chain: its blob's pixel list in blob_pix, the volatile point
__global__ void __launch_bounds__(256)
k_fill_columns(pword *__restrict__ colbase, uint64_t A, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off,
               const FillDesc *__restrict__ desc, const uint32_t *__restrict__ blob_pix, uint32_t cw, uint64_t seed, uint64_t nf, uint64_t f) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    const uint32_t g = chain_of[i];
    const uint64_t p = i - chain_off[g], stream = (uint64_t) g * nf + f + 1;
    const FillDesc d = desc[g];
    const uint64_t npix = d.npix;
    const uint32_t *pix = blob_pix + d.pix_off;
    pword w;
    if (npix == 0) {
        if (p == 0) w = pw_make(d.vx, d.vy, 0, 0, F_HAS_FLUID);
        else {
            uint64_t r = rng64(seed, stream, p);
            w = pw_make(d.vx, d.vy, (uint32_t) (r & 255u), (uint32_t) ((r >> 8) & 255u), 0);   // duplicate of the volatile point
        }
    } else if (p < npix) {
        uint32_t ci = pix[p];
        w = pw_make(ci % cw, ci / cw, 0, 0, F_HAS_PIXEL | F_HAS_FLUID);
    } else {
        uint64_t src = feistel_perm(p % npix, npix, seed ^ (stream << 20));
        uint32_t ci = pix[src];
        uint64_t r = rng64(seed, stream, p);
        w = pw_make(ci % cw, ci / cw, (uint32_t) (r & 255u), (uint32_t) ((r >> 8) & 255u), F_HAS_PIXEL);
    }
    colbase[i] = w;
}

// thread.cpp:1187-1233 on the host mirror: volatile blobs take positions interpolated between the
// neighbouring non-empty blobs of their group along the key-frame cycle.
static void fix_volatiles(std::vector<BlobHost *> &v) {
    size_t sz = v.size();
    if (sz <= 1) return;
    size_t i = 0;
    bool started = false;
    std::vector<BlobHost *> vol;
    BlobHost *first_static = nullptr, *prev_static = nullptr;
    for (;;) {
        i = (i + 1) % sz;
        BlobHost *bl = v[i];
        bool empty = bl->size == 0;
        if (!started) {
            if (empty) { if (i == 0) break; continue; }
            started = true; first_static = bl; prev_static = bl;
            continue;
        }
        if (empty) { vol.push_back(bl); continue; }
        if (!vol.empty()) {
            size_t vsz = vol.size();
            for (size_t k = 0; k < vsz; ++k) {
                double t = (k + 1.0) / double(vsz + 1.0);
                vol[k]->stats[0] = t * bl->stats[0] + (1.0 - t) * prev_static->stats[0];
                vol[k]->stats[1] = t * bl->stats[1] + (1.0 - t) * prev_static->stats[1];
            }
            vol.clear();
        }
        prev_static = bl;
        if (prev_static == first_static) break;
    }
}

int engine_init_chains(Engine *E) {
    size_t nf = E->frames.size();
    if (nf == 0 || !E->map_ready || E->map_w == 0) { E->err = "init_chains: blob map missing"; return AMX_ERR_STATE; }
    uint32_t W = E->map_w;
    // group g -> blob of frame f
    std::vector<std::vector<int64_t>> bog(nf, std::vector<int64_t>(W, -1));
    for (size_t f = 0; f < nf; ++f)
        for (size_t b = 0; b < E->frames[f].blobs.size(); ++b) {
            uint64_t g = E->frames[f].blobs[b].group;
            if (g < W) bog[f][g] = (int64_t) b;
        }
    for (uint32_t g = 0; g < W; ++g) {
        std::vector<BlobHost *> cyc;
        for (size_t f = 0; f < nf; ++f) {
            if (bog[f][g] < 0) { E->err = "init_chains: group without blob"; return AMX_ERR_STATE; }
            cyc.push_back(&E->frames[f].blobs[bog[f][g]]);
        }
        fix_volatiles(cyc);
    }
    std::vector<uint64_t> keys(W), widths(W), maxs(W);
    for (uint32_t g = 0; g < W; ++g) {
        uint64_t mx = 0;
        for (size_t f = 0; f < nf; ++f) mx = std::max<uint64_t>(mx, E->frames[f].blobs[bog[f][g]].size);
        keys[g] = g; maxs[g] = mx; widths[g] = mx * E->p.density;
    }
    int rc = engine_alloc_chains(E, W, keys.data(), widths.data(), maxs.data(), (uint32_t) nf);
    if (rc != AMX_OK) return rc;
    uint64_t maxw = 0;
    for (uint32_t g = 0; g < W; ++g) maxw = std::max(maxw, widths[g]);
    if (maxw <= 1) { E->err = "init_chains: no chain wider than 1"; return AMX_ERR_STATE; }   // thread.cpp:882-888
    FillDesc *d_desc = nullptr;
    if (!dev_alloc(E, (void **) &d_desc, (size_t) W * sizeof(FillDesc), "chain fill descriptors")) return AMX_ERR_NOMEM;
    std::vector<FillDesc> desc(W);
    for (size_t f = 0; f < nf && rc == AMX_OK; ++f) {
        FrameDev &fr = E->frames[f];
        if (!fr.blob_pix && fr.pixel_count) { rc = engine_build_blob_pixels(E, (uint32_t) f); if (rc != AMX_OK) break; }
        for (uint32_t g = 0; g < W; ++g) {
            const BlobHost &bl = fr.blobs[bog[f][g]];
            desc[g].npix = (uint32_t) bl.size;
            desc[g].pix_off = bl.size ? (uint32_t) fr.blob_pix_off[bog[f][g]] : 0u;
            desc[g].vx = (uint32_t) ((int32_t) std::round(bl.stats[0])) & 0xffffu;
            desc[g].vy = (uint32_t) ((int32_t) std::round(bl.stats[1])) & 0xffffu;
        }
        // (pageable source: the copy has left `desc` when the call returns, so the next frame may overwrite it)
        if (E->fail(cudaMemcpyAsync(d_desc, desc.data(), (size_t) W * sizeof(FillDesc), cudaMemcpyHostToDevice, E->stream), "chain fill descriptors H2D")) { rc = AMX_ERR_CUDA; break; }
        k_fill_columns<<<(unsigned) div_up(E->A, 256), 256, 0, E->stream>>>(E->table + f * E->A, E->A, E->chain_of, E->d_chain_off, d_desc, fr.blob_pix,
                                                                            E->cw, E->p.seed, (uint64_t) nf, (uint64_t) f);
        E->launches++;
    }
    if (rc == AMX_OK && (E->fail(cudaStreamSynchronize(E->stream), "init_chains") || E->check("init_chains"))) rc = AMX_ERR_CUDA;
    else if (rc != AMX_OK) cudaStreamSynchronize(E->stream);
    dev_free(d_desc);
    return rc;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_init_chains(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_init_chains(&ctx->e);
}

int amx_chain_count(amx_ctx *ctx, uint32_t *count) {
    if (!ctx || !count) return AMX_ERR_ARG;
    *count = ctx->e.nchains;
    return AMX_OK;
}

int amx_chain_info(amx_ctx *ctx, uint32_t chain, uint64_t info4[4]) {
    if (!ctx || !info4 || chain >= ctx->e.nchains) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    info4[0] = E->chain_key[chain];
    info4[1] = E->chain_off[chain + 1] - E->chain_off[chain];
    info4[2] = E->h;
    info4[3] = E->chain_max_surface[chain];
    return AMX_OK;
}

int amx_export_chain(amx_ctx *ctx, uint32_t chain, uint64_t *words_out) {
    if (!ctx || !words_out || chain >= ctx->e.nchains) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    for (uint32_t j = 0; j < E->h; ++j)
        if (E->fail(cudaMemcpyAsync(words_out + (size_t) j * w, E->table + (size_t) j * E->A + off, w * 8, cudaMemcpyDeviceToHost, E->stream), "export chain"))
            return AMX_ERR_CUDA;
    return E->fail(cudaStreamSynchronize(E->stream), "export chain") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_import_chains(amx_ctx *ctx, uint32_t nchains, const uint64_t *keys, const uint64_t *widths, const uint64_t *max_surface,
                      uint32_t height, const uint64_t *words) {
    if (!ctx || !keys || !widths || !max_surface || (!words && nchains)) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    int rc = engine_alloc_chains(E, nchains, keys, widths, max_surface, height);
    if (rc != AMX_OK) return rc;
    const uint64_t *src = words;
    for (uint32_t c = 0; c < nchains; ++c) {
        uint64_t off = E->chain_off[c], w = widths[c];
        for (uint32_t j = 0; j < height; ++j) {
            if (w && E->fail(cudaMemcpyAsync(E->table + (size_t) j * E->A + off, src + (size_t) j * w, w * 8, cudaMemcpyHostToDevice, E->stream), "import chain"))
                return AMX_ERR_CUDA;
        }
        src += (size_t) height * w;
    }
    if (E->fail(cudaStreamSynchronize(E->stream), "import chains")) return AMX_ERR_CUDA;
    E->state = ST_ATOM_MORPHING;
    E->counter = 0;
    return AMX_OK;
}

int amx_table_device_ptr(amx_ctx *ctx, uint32_t column, void **d_ptr, uint64_t *total_atoms) {
    if (!ctx || !d_ptr || column >= ctx->e.h) return AMX_ERR_ARG;
    *d_ptr = ctx->e.table + (size_t) column * ctx->e.A;
    if (total_atoms) *total_atoms = ctx->e.A;
    return AMX_OK;
}

}
