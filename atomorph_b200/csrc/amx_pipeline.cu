/*
 * amx_pipeline.cu -- the 4-state pipeline of the reference worker (thread::step, thread.cpp:139-171)
 * driven on the device: blob detection -> unification -> matching -> atom morphing -> done.
 *
 * Step accounting mirrors the reference so that callers that count work through iterate(n) keep
 * working (SURVEY.md section 9 note 14):
 *   - blob detection: the reference needs one step per merge; the GPU pass finishes in ONE step
 *   - unification:    one step (engine_unify: dust clustering, only with blob_min_size > 1)
 *   - matching:       first step builds the blob map, each further step is one parallel round
 *   - atom morphing:  one step = max(1, threads) * cycle_length swap proposals on ONE chain,
 *                     chains served round-robin (thread.cpp:1043-1064).  n steps on a C-chain
 *                     scene are executed as floor(n/C) all-chain sweeps plus n mod C single steps;
 *                     rounds are issued until the device proposal counter reaches the target.
 */
#include <algorithm>
#include "amx_engine.h"

namespace amx {

// expected proposals of one round on a chain of width w (pairs (l, l^m) with both ends < w, m uniform in [1, 2^k))
static double pairs_per_round(uint64_t w) {
    if (w < 2) return 0.0;
    unsigned k = 0;
    while ((1ull << k) < w) ++k;
    double full = (double) (1ull << k);
    // P(l^m < w | l < w) ~ (w-1)/(2^k-1); pairs = w * that / 2
    return 0.5 * (double) w * ((double) (w - 1) / (full - 1.0));
}

static int read_stats(Engine *E) {
    if (E->fail(cudaMemcpyAsync(E->swapstats, E->d_swapstats, 24, cudaMemcpyDeviceToHost, E->stream), "stats D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "stats"))
        return AMX_ERR_CUDA;
    return AMX_OK;
}

// `steps` morph steps on chain c (c < 0: every chain gets `steps` steps, in one sweep per round)
static int morph_steps(Engine *E, int32_t c, uint64_t steps) {
    uint64_t per_step = std::max<uint64_t>(1, E->p.threads) * E->p.cycle_length;
    if (per_step == 0 || steps == 0) return AMX_OK;
    double ppr = 0.0;
    if (c >= 0) ppr = pairs_per_round(E->chain_off[c + 1] - E->chain_off[c]);
    else for (uint32_t i = 0; i < E->nchains; ++i) ppr += pairs_per_round(E->chain_off[i + 1] - E->chain_off[i]);
    if (ppr <= 0.0) return AMX_OK;                          // chains of width <= 1 burn their steps (thread.cpp:1045)
    double target_d = (double) per_step * (double) steps * (c >= 0 ? 1.0 : (double) E->nchains);
    // for an all-chain sweep the per-chain target is per_step*steps; rounds needed is governed by the slowest chain
    double rounds_d;
    if (c >= 0) rounds_d = target_d / ppr;
    else {
        rounds_d = 0.0;
        for (uint32_t i = 0; i < E->nchains; ++i) {
            double p = pairs_per_round(E->chain_off[i + 1] - E->chain_off[i]);
            if (p > 0.0) rounds_d = std::max(rounds_d, (double) per_step * (double) steps / p);
        }
    }
    uint64_t rounds = (uint64_t) std::max(1.0, std::ceil(rounds_d));
    int rc = read_stats(E);
    if (rc != AMX_OK) return rc;
    uint64_t before = E->swapstats[0];
    rc = engine_swap_rounds(E, c, -1, rounds);
    if (rc != AMX_OK) return rc;
    if (c >= 0) {
        // top up until the counted proposals reach the reference-equivalent amount
        for (int guard = 0; guard < 64; ++guard) {
            rc = read_stats(E);
            if (rc != AMX_OK) return rc;
            double done = (double) (E->swapstats[0] - before);
            if (done >= target_d) break;
            uint64_t more = (uint64_t) std::max(1.0, std::ceil((target_d - done) / ppr));
            rc = engine_swap_rounds(E, c, -1, more);
            if (rc != AMX_OK) return rc;
        }
    }
    return AMX_OK;
}

static int pipeline_steps(Engine *E, uint64_t nsteps) {
    while (nsteps > 0) {
        switch (E->state) {
            case ST_BLOB_DETECTION: {
                int rc = engine_blobify(E);
                if (rc != AMX_OK) return rc;
                E->state = ST_BLOB_UNIFICATION; E->counter = 0;
                break;
            }
            case ST_BLOB_UNIFICATION: {
                int rc = engine_unify(E);                 // dust clustering: a no-op with the default blob_min_size = 1
                if (rc != AMX_OK) return rc;
                E->state = ST_BLOB_MATCHING; E->counter = 0;
                break;
            }
            case ST_BLOB_MATCHING: {
                bool done = E->skip_state;
                if (!done) {
                    if (!E->map_ready) { int rc = engine_match_init(E); if (rc != AMX_OK) return rc; }
                    else if (E->map_h == 0 || E->map_w <= 1) done = true;
                    else {
                        // consume as many matching steps as requested in one go
                        uint64_t rounds = E->skip_state ? 0 : nsteps;
                        int rc = engine_match_rounds(E, rounds);
                        if (rc != AMX_OK) return rc;
                        E->counter += rounds;
                        if (E->blob_map_e == 0.0) done = true;
                        else { E->skip_state = false; return AMX_OK; }
                    }
                }
                if (done) {
                    if (!E->map_ready) { int rc = engine_match_init(E); if (rc != AMX_OK) return rc; }
                    int rc = engine_init_chains(E);
                    E->state = (rc == AMX_OK) ? ST_ATOM_MORPHING : ST_DONE;
                    E->counter = 0;
                    if (rc != AMX_OK && rc != AMX_ERR_STATE) return rc;
                }
                break;
            }
            case ST_ATOM_MORPHING: {
                if (E->skip_state) { E->state = ST_DONE; E->counter = 0; break; }
                uint64_t C = E->nchains;
                if (C == 0) { E->state = ST_DONE; break; }
                // all remaining steps are morph steps: batch them
                uint64_t sweeps = C > 1 ? nsteps / C : 0;
                uint64_t singles = C > 1 ? nsteps % C : nsteps;
                if (sweeps) { int rc = morph_steps(E, -1, sweeps); if (rc != AMX_OK) return rc; }
                if (C == 1) { int rc = morph_steps(E, 0, singles); if (rc != AMX_OK) return rc; }
                else for (uint64_t s = 0; s < singles; ++s) {
                    int rc = morph_steps(E, (int32_t) ((E->counter + sweeps * C + s) % C), 1);
                    if (rc != AMX_OK) return rc;
                }
                E->counter += nsteps;
                E->skip_state = false;
                return AMX_OK;
            }
            default:
                E->skip_state = false;
                return AMX_OK;
        }
        E->skip_state = false;
        E->counter++;
        --nsteps;
    }
    return AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_step(amx_ctx *ctx, uint64_t nsteps) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return pipeline_steps(&ctx->e, nsteps);
}

int amx_next_state(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    ctx->e.skip_state = true;
    return AMX_OK;
}

unsigned amx_get_state(amx_ctx *ctx) { return ctx ? ctx->e.state : ST_DONE; }

double amx_get_energy(amx_ctx *ctx) {
    if (!ctx) return -1.0;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (E->state == ST_BLOB_MATCHING) return E->map_ready ? E->blob_map_e : 0.0;
    if (E->state == ST_ATOM_MORPHING) { double c = -1.0; engine_cost(E, &c); return c; }
    return -1.0;
}

}
