/*
 * amx_core.cu -- context lifetime, parameters, ingest and the flat C-ABI glue (include/amx.h).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include "amx_engine.h"

namespace amx {

bool Engine::fail(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return false;
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    err = buf;
    return true;
}
bool Engine::check(const char *what) { return fail(cudaGetLastError(), what); }

bool dev_alloc(Engine *E, void **p, size_t bytes, const char *what) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { E->fail(e, what); *p = nullptr; return false; }
    return true;
}
void dev_free(void *p) { if (p) cudaFree(p); }

static void free_frames(Engine *E) {
    for (auto &f : E->frames) {
        dev_free(f.stored); dev_free(f.fetch); dev_free(f.present); dev_free(f.label); dev_free(f.blob_pix);
    }
    E->frames.clear();
}
static void free_chains(Engine *E) {
    engine_dist_table_gone(E);
    dev_free(E->table); E->table = nullptr;
    dev_free(E->chain_of); E->chain_of = nullptr;
    dev_free(E->d_chain_off); E->d_chain_off = nullptr;
    dev_free(E->d_partials); E->d_partials = nullptr; E->n_partials = 0;
    E->nchains = 0; E->h = 0; E->A = 0;
    E->chain_key.clear(); E->chain_off.clear(); E->chain_max_surface.clear();
    E->swapstats[0] = E->swapstats[1] = E->swapstats[2] = 0;
    E->render_ready = false;
}
static void free_match(Engine *E) {
    dev_free(E->d_bfeat); E->d_bfeat = nullptr;
    dev_free(E->d_bmap); E->d_bmap = nullptr;
    dev_free(E->d_menergy); E->d_menergy = nullptr;
    E->blob_map.clear(); E->map_w = E->map_h = 0; E->map_ready = false; E->blob_map_e = 0.0;
}

static int reset_all(Engine *E) {
    cudaStreamSynchronize(E->stream);
    free_frames(E);
    free_chains(E);
    free_match(E);
    engine_render_free(E);
    engine_fluid_free(E);
    E->state = ST_BLOB_DETECTION;
    E->skip_state = false;
    E->counter = 0;
    E->rng_round = 0;
    return AMX_OK;
}

} // namespace amx

using namespace amx;

extern "C" {

const char *amx_version(void) { return "1.0-b200"; }

int amx_create(amx_ctx **out, int device) {
    if (!out) return AMX_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return AMX_ERR_CUDA;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return AMX_ERR_CUDA;
    amx_ctx *c = new (std::nothrow) amx_ctx();
    if (!c) return AMX_ERR_NOMEM;
    c->e.device = device;
    cudaDeviceGetAttribute(&c->e.sm_count, cudaDevAttrMultiProcessorCount, device);
    if (c->e.sm_count < 1) c->e.sm_count = 148;
    if (cudaStreamCreateWithFlags(&c->e.stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return AMX_ERR_CUDA; }
    c->e.own_stream = true;
    if (const char *nb = getenv("AMX_RENDER_BATCH")) c->e.render_batch = (uint32_t) std::max(1, atoi(nb));   // tuning knob (frames per launch pair)
    c->e.swap_global_only = getenv("AMX_SWAP_GLOBAL") != nullptr;
    if (const char *lo = getenv("AMX_SWAP_LOCALITY")) c->e.swap_locality = (uint32_t) std::max(0, atoi(lo));   // every n-th tiled epoch pairs spatial neighbours
    if (const char *ac = getenv("AMX_RENDER_ACC")) c->e.tiled_acc = atoi(ac) != 0;
    if (const char *tl = getenv("AMX_RENDER_TILED")) { c->e.tiled_enabled = atoi(tl) != 0; c->e.tiled_multi = atoi(tl) >= 2; }   // 0: general A-buffer path only; 2: tiled path for multi-chain morphs too
    if (const char *la = getenv("AMX_LOOKAHEAD")) c->e.lookahead = atoi(la) != 0;
    cudaEventCreate(&c->e.ev0);
    cudaEventCreate(&c->e.ev1);
    if (!dev_alloc(&c->e, (void **) &c->e.d_swapstats, 3 * sizeof(uint64_t), "swapstats")) { delete c; return AMX_ERR_NOMEM; }
    cudaMemset(c->e.d_swapstats, 0, 3 * sizeof(uint64_t));
    *out = c;
    return AMX_OK;
}

void amx_destroy(amx_ctx *ctx) {
    if (!ctx) return;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    reset_all(E);
    engine_dist_free(E);
    dev_free(E->d_swapstats);
    dev_free(E->loc_buf); dev_free(E->loc_tmp);
    dev_free(E->d_out);
    dev_free(E->d_perlin);
    if (E->copy_stream) cudaStreamDestroy(E->copy_stream);
    for (int k = 0; k < 4; ++k) if (E->copy_ev[k]) cudaEventDestroy(E->copy_ev[k]);
    if (E->ev0) cudaEventDestroy(E->ev0);
    if (E->ev1) cudaEventDestroy(E->ev1);
    if (E->own_stream && E->stream) cudaStreamDestroy(E->stream);
    delete ctx;
}

const char *amx_last_error(amx_ctx *ctx) { return ctx ? ctx->e.err.c_str() : "null context"; }

int amx_set_stream(amx_ctx *ctx, void *cuda_stream) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    cudaStreamSynchronize(E->stream);
    if (cuda_stream == AMX_STREAM_PRIVATE) {
        // back to a private non-blocking stream
        if (!E->own_stream) {
            if (cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking) != cudaSuccess) return AMX_ERR_CUDA;
            E->own_stream = true;
        }
        return AMX_OK;
    }
    // NULL is a stream like any other: the legacy default stream (what torch.cuda.current_stream().cuda_stream is unless
    // the caller entered a torch.cuda.stream context).  The engine's work is then ordered with everything else on it.
    if (E->own_stream && E->stream) cudaStreamDestroy(E->stream);
    E->stream = (cudaStream_t) cuda_stream;
    E->own_stream = false;
    return AMX_OK;
}

void *amx_get_stream(amx_ctx *ctx) { return ctx ? (void *) ctx->e.stream : nullptr; }

int amx_device_sync(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return ctx->e.fail(cudaStreamSynchronize(ctx->e.stream), "sync") ? AMX_ERR_CUDA : AMX_OK;
}

static uint64_t to_u64(double v) {
    if (v >= 1.8e19) return UINT64_MAX;
    if (v < 0) return 0;
    return (uint64_t) v;
}

int amx_set_param(amx_ctx *ctx, int id, double v) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    Params &p = E->p;
    switch (id) {
        case AMX_P_BLOB_DELIMITER:   p.blob_delimiter = (unsigned) v; break;
        case AMX_P_BLOB_THRESHOLD:   p.blob_threshold = v; break;
        case AMX_P_BLOB_MAX_SIZE:    p.blob_max_size = to_u64(v); break;
        case AMX_P_BLOB_MIN_SIZE:    p.blob_min_size = to_u64(v); break;
        case AMX_P_BLOB_BOX_GRIP:    p.blob_box_grip = (uint32_t) v; break;
        case AMX_P_BLOB_BOX_SAMPLES: p.blob_box_samples = to_u64(v); break;
        case AMX_P_BLOB_NUMBER:      p.blob_number = to_u64(v); break;
        case AMX_P_BLOB_RGBA_WEIGHT: p.blob_rgba_weight = (unsigned) v & 255u; break;
        case AMX_P_BLOB_SIZE_WEIGHT: p.blob_size_weight = (unsigned) v & 255u; break;
        case AMX_P_BLOB_XY_WEIGHT:   p.blob_xy_weight = (unsigned) v & 255u; break;
        case AMX_P_DEGENERATION:     p.degeneration = to_u64(v); break;
        case AMX_P_DENSITY:          p.density = (uint32_t) v; break;
        case AMX_P_MOTION:           p.motion = (unsigned) v; break;
        case AMX_P_FADING:           if (p.fading != (unsigned) v) E->render_ready = false; p.fading = (unsigned) v; break;
        case AMX_P_THREADS:          p.threads = to_u64(v); break;
        case AMX_P_CYCLE_LENGTH:     p.cycle_length = to_u64(v); break;
        case AMX_P_FEATHER:          p.feather = to_u64(v); break;
        case AMX_P_KEEP_BACKGROUND:  p.keep_background = (v != 0.0); break;
        case AMX_P_FINITE:           p.finite = (v != 0.0); break;
        case AMX_P_SHOW_BLOBS:       p.show_blobs = (unsigned) v; break;
        case AMX_P_FLUID:            p.fluid = (unsigned) v; break;
        case AMX_P_SEED:             if (p.seed != (unsigned) v) E->render_ready = false; p.seed = (unsigned) v; break;
        default: return AMX_ERR_ARG;
    }
    return AMX_OK;
}

double amx_get_param(amx_ctx *ctx, int id) {
    if (!ctx) return 0.0;
    Params &p = ctx->e.p;
    switch (id) {
        case AMX_P_BLOB_DELIMITER:   return p.blob_delimiter;
        case AMX_P_BLOB_THRESHOLD:   return p.blob_threshold;
        case AMX_P_BLOB_MAX_SIZE:    return (double) p.blob_max_size;
        case AMX_P_BLOB_MIN_SIZE:    return (double) p.blob_min_size;
        case AMX_P_BLOB_BOX_GRIP:    return p.blob_box_grip;
        case AMX_P_BLOB_BOX_SAMPLES: return (double) p.blob_box_samples;
        case AMX_P_BLOB_NUMBER:      return (double) p.blob_number;
        case AMX_P_BLOB_RGBA_WEIGHT: return p.blob_rgba_weight;
        case AMX_P_BLOB_SIZE_WEIGHT: return p.blob_size_weight;
        case AMX_P_BLOB_XY_WEIGHT:   return p.blob_xy_weight;
        case AMX_P_DEGENERATION:     return (double) p.degeneration;
        case AMX_P_DENSITY:          return p.density;
        case AMX_P_MOTION:           return p.motion;
        case AMX_P_FADING:           return p.fading;
        case AMX_P_THREADS:          return (double) p.threads;
        case AMX_P_CYCLE_LENGTH:     return (double) p.cycle_length;
        case AMX_P_FEATHER:          return (double) p.feather;
        case AMX_P_KEEP_BACKGROUND:  return p.keep_background;
        case AMX_P_FINITE:           return p.finite;
        case AMX_P_SHOW_BLOBS:       return p.show_blobs;
        case AMX_P_FLUID:            return p.fluid;
        case AMX_P_SEED:             return p.seed;
        default: return 0.0;
    }
}

int amx_reset(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return reset_all(&ctx->e);
}

int amx_set_canvas(amx_ctx *ctx, uint32_t width, uint32_t height, uint32_t canvas_w, uint32_t canvas_h, const uint16_t bbox[4]) {
    if (!ctx || !bbox) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    if (canvas_w < width) canvas_w = width;
    if (canvas_h < height) canvas_h = height;
    if (canvas_w == 0 || canvas_h == 0 || canvas_w > 65536 || canvas_h > 65536) return AMX_ERR_ARG;
    if (!E->frames.empty() && (canvas_w != E->cw || canvas_h != E->ch)) {
        E->err = "canvas change with frames present: call amx_reset first";
        return AMX_ERR_STATE;
    }
    E->width = width; E->height = height; E->cw = canvas_w; E->ch = canvas_h;
    for (int i = 0; i < 4; ++i) E->bbox[i] = bbox[i];
    E->render_ready = false;
    return AMX_OK;
}

int amx_set_frame_count(amx_ctx *ctx, uint32_t nframes, const uint64_t *keys) {
    if (!ctx || (nframes && !keys)) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (E->cw == 0) { E->err = "amx_set_canvas first"; return AMX_ERR_STATE; }
    cudaStreamSynchronize(E->stream);
    for (auto &f : E->frames) { dev_free(f.stored); dev_free(f.fetch); dev_free(f.present); dev_free(f.label); dev_free(f.blob_pix); }
    E->frames.clear();
    E->frames.resize(nframes);
    size_t n = E->canvas();
    for (uint32_t i = 0; i < nframes; ++i) {
        FrameDev &f = E->frames[i];
        f.key = keys[i];
        if (!dev_alloc(E, (void **) &f.stored, n * 4, "frame.stored") || !dev_alloc(E, (void **) &f.fetch, n * 4, "frame.fetch") ||
            !dev_alloc(E, (void **) &f.present, n, "frame.present") || !dev_alloc(E, (void **) &f.label, n * 4, "frame.label"))
            return AMX_ERR_NOMEM;
        cudaMemsetAsync(f.stored, 0, n * 4, E->stream);
        cudaMemsetAsync(f.fetch, 0, n * 4, E->stream);
        cudaMemsetAsync(f.present, 0, n, E->stream);
        cudaMemsetAsync(f.label, 0xff, n * 4, E->stream);
    }
    E->state = ST_BLOB_DETECTION;
    E->render_ready = false;
    return AMX_OK;
}

static int upload_common(Engine *E, uint32_t index, const uint32_t *d_rgba, const double means[6]) {
    FrameDev &f = E->frames[index];
    if (means) for (int i = 0; i < 6; ++i) f.means[i] = means[i];
    int rc = engine_upload_convert(E, index, d_rgba);
    if (rc != AMX_OK) return rc;
    f.uploaded = true;
    E->render_ready = false;
    return AMX_OK;
}

int amx_upload_frame(amx_ctx *ctx, uint32_t index, const uint32_t *rgba, const uint8_t *present, const double means[6]) {
    if (!ctx || !rgba || !present) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (index >= E->frames.size()) return AMX_ERR_ARG;
    size_t n = E->canvas();
    FrameDev &f = E->frames[index];
    uint32_t *d_raw = nullptr;
    if (!dev_alloc(E, (void **) &d_raw, n * 4, "upload staging")) return AMX_ERR_NOMEM;
    if (E->fail(cudaMemcpyAsync(d_raw, rgba, n * 4, cudaMemcpyHostToDevice, E->stream), "H2D rgba") ||
        E->fail(cudaMemcpyAsync(f.present, present, n, cudaMemcpyHostToDevice, E->stream), "H2D present")) {
        dev_free(d_raw);
        return AMX_ERR_CUDA;
    }
    int rc = upload_common(E, index, d_raw, means);
    cudaStreamSynchronize(E->stream);
    dev_free(d_raw);
    return rc;
}

int amx_upload_frame_device(amx_ctx *ctx, uint32_t index, const uint32_t *d_rgba, const uint8_t *d_present, const double means[6]) {
    if (!ctx || !d_rgba || !d_present) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (index >= E->frames.size()) return AMX_ERR_ARG;
    if (E->fail(cudaMemcpyAsync(E->frames[index].present, d_present, E->canvas(), cudaMemcpyDeviceToDevice, E->stream), "D2D present"))
        return AMX_ERR_CUDA;
    return upload_common(E, index, d_rgba, means);
}

static int download_u32(Engine *E, const uint32_t *d, uint32_t *out) {
    if (E->fail(cudaMemcpyAsync(out, d, E->canvas() * 4, cudaMemcpyDeviceToHost, E->stream), "D2H")) return AMX_ERR_CUDA;
    return E->fail(cudaStreamSynchronize(E->stream), "sync") ? AMX_ERR_CUDA : AMX_OK;
}
int amx_download_fetch(amx_ctx *ctx, uint32_t index, uint32_t *out) {
    if (!ctx || !out || index >= ctx->e.frames.size()) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return download_u32(&ctx->e, ctx->e.frames[index].fetch, out);
}
int amx_download_stored(amx_ctx *ctx, uint32_t index, uint32_t *out) {
    if (!ctx || !out || index >= ctx->e.frames.size()) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return download_u32(&ctx->e, ctx->e.frames[index].stored, out);
}

uint64_t amx_launch_count(amx_ctx *ctx) { return ctx ? ctx->e.launches : 0; }

int amx_timer_start(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    return ctx->e.fail(cudaEventRecord(ctx->e.ev0, ctx->e.stream), "event") ? AMX_ERR_CUDA : AMX_OK;
}
int amx_timer_stop(amx_ctx *ctx, float *ms) {
    if (!ctx || !ms) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    if (E->fail(cudaEventRecord(E->ev1, E->stream), "event") || E->fail(cudaEventSynchronize(E->ev1), "event sync") ||
        E->fail(cudaEventElapsedTime(ms, E->ev0, E->ev1), "elapsed"))
        return AMX_ERR_CUDA;
    return AMX_OK;
}

} // extern "C"
