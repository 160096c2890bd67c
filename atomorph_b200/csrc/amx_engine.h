/*
 * amx_engine.h -- internal state of one device engine (one per am::morph, one per GPU).
 * Not part of the ABI.  Device memory layout (DESIGN.md "Data layout in HBM"):
 *
 *   per key frame f (canvas = cw*ch positions, row-major):
 *     stored[f]  u32   colour as the reference stores it (HSP bytes when blob_delimiter==HSP)
 *     fetch[f]   u32   colour as get_pixel returns it (RGB after the 8-bit HSP round trip), 0 if absent
 *     present[f] u8    1 where a pixel was added
 *     label[f]   i32   blob index in the frame's blob vector, -1 if absent
 *   chains: all chains concatenated along x; A = sum of widths
 *     table      u64   [h][A]  key-point words, column-major (coalesced along atoms)
 *     chain_of   u32   [A]     chain index of an atom
 *   render inputs (refreshed by amx_render_prepare), per interval y:
 *     rc1, rc2   u32   [h][A]  resolved end colours (one-sided alpha rule applied)
 *     rlag,rslope f64  [h][A]  Perlin lag / slope (only when fading == PERLIN)
 */
#ifndef AMX_ENGINE_H
#define AMX_ENGINE_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/amx.h"
#include "amx_math.h"

namespace amx {

struct BlobHost {
    double   stats[6];   // x y r g b a (doubles in 0..1 for colour, stored colour space)
    uint64_t size;       // surface size (0 = volatile)
    uint64_t group;
};

struct FrameDev {
    uint64_t  key = 0;
    uint32_t *stored = nullptr;
    uint32_t *fetch = nullptr;
    uint8_t  *present = nullptr;
    int32_t  *label = nullptr;
    double    means[6] = {0, 0, 0, 0, 0, 0};
    uint64_t  pixel_count = 0;
    bool      uploaded = false;
    std::vector<BlobHost> blobs;          // blob vector order (index b of get_pixels(b,...))
    // sorted pixel list: positions (canvas linear index) grouped by blob, ascending inside a blob
    uint32_t *blob_pix = nullptr;         // [pixel_count]
    std::vector<uint64_t> blob_pix_off;   // [nblobs+1]
};

struct Params {
    unsigned blob_delimiter = K_HSP;
    double   blob_threshold = 1.0;
    uint64_t blob_max_size = UINT64_MAX;
    uint64_t blob_min_size = 1;
    uint32_t blob_box_grip = 65535;
    uint64_t blob_box_samples = 10;
    uint64_t blob_number = 1;
    unsigned blob_rgba_weight = 1, blob_size_weight = 1, blob_xy_weight = 1;
    uint64_t degeneration = 0;
    uint32_t density = 1;
    unsigned motion = K_SPLINE;
    unsigned fading = K_PERLIN;
    uint64_t threads = 0;
    uint64_t cycle_length = 1000;
    uint64_t feather = 0;
    bool     keep_background = false;
    bool     finite = false;
    unsigned show_blobs = SHOW_TEXTURE;
    unsigned fluid = 0;
    unsigned seed = 0;
};

struct Fluid;   // amx_fluid.cu
struct Dist;    // amx_dist.cu

struct Engine {
    int          device = 0;
    int          sm_count = 148;
    cudaStream_t stream = nullptr;
    bool         own_stream = false;
    std::string  err;
    uint64_t     launches = 0;
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;    // device -> host frame copies overlap rendering (amx_render with a host buffer)
    cudaEvent_t  copy_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t     copy_ev_next = 0;

    Params p;
    uint32_t width = 0, height = 0, cw = 0, ch = 0;
    uint16_t bbox[4] = {65535, 65535, 0, 0};
    std::vector<FrameDev> frames;

    // state machine
    unsigned state = ST_BLOB_DETECTION;
    bool     skip_state = false;
    uint64_t counter = 0;
    uint64_t rng_round = 0;          // counter for the counter-based RNG streams

    // blob map (K3): map[f*map_w + x] = blob index (in frame f's vector) at slot x
    uint32_t map_w = 0, map_h = 0;
    std::vector<uint32_t> blob_map;
    double   blob_map_e = 0.0;
    bool     map_ready = false;
    // device mirror for the matching kernel
    double  *d_bfeat = nullptr;      // [map_h][map_w][8]: size, x(trunc), y(trunc), r,g,b,a (u8 as double), pad
    uint32_t *d_bmap = nullptr;      // [map_h][map_w]
    double  *d_menergy = nullptr;

    // chains
    uint32_t nchains = 0, h = 0;
    uint64_t A = 0;
    std::vector<uint64_t> chain_key, chain_off, chain_max_surface;   // off has nchains+1 entries
    pword    *table = nullptr;
    uint32_t *chain_of = nullptr;
    uint64_t *d_chain_off = nullptr;  // [nchains+1]
    uint64_t *d_swapstats = nullptr;  // proposals, accepted, gain
    uint64_t  swapstats[3] = {0, 0, 0};
    bool      swap_global_only = false;   // AMX_SWAP_GLOBAL=1: one-round-per-launch kernels only (for comparisons)
    uint32_t  swap_locality = 0;          // every n-th tiled epoch pairs spatial neighbours (amx_set_swap_locality / AMX_SWAP_LOCALITY; 0 = never)
    uint32_t *loc_buf = nullptr;          // locality epochs: Morton keys / atom order (4 x 2^k u32) and the sort workspace
    void     *loc_tmp = nullptr;
    size_t    loc_cap = 0, loc_tmp_bytes = 0;
    uint64_t *d_partials = nullptr;   // cost partial sums
    uint32_t  n_partials = 0;

    // render
    bool      render_ready = false;
    uint64_t  prepare_count = 0;           // table refreshes so far (a key of the frame look-ahead ring)
    void     *pix_ring = nullptr;          // amx_render.cu: PixRing, the look-ahead of amx_render_pixels
    bool      lookahead = true;            // amx_set_lookahead / AMX_LOOKAHEAD=0
    // render inputs per interval, atoms sorted by the tile of their mid-interval position (amx_render.cu: RIn)
    uint32_t *rc1 = nullptr, *rc2 = nullptr;
    double   *rlag = nullptr, *rslope = nullptr;
    pword    *rpts = nullptr;              // [h][rnpt][A]
    uint32_t *ratom = nullptr, *rchain = nullptr;
    uint32_t  rnpt = 0;
    std::vector<uint32_t> r_live;          // live (drawn) atoms per interval
    uint64_t  rbuf_A = 0, rbuf_cv = 0;     // geometry the render buffers were allocated for
    uint32_t  rbuf_h = 0, rbuf_nchains = 0;
    uint32_t *sort_buf = nullptr;          // radix-sort scratch of render prepare
    void     *sort_tmp = nullptr;
    size_t    sort_tmp_bytes = 0;
    int32_t  *d_blob_of_chain = nullptr;   // [h][nchains] blob vector index of chain c in frame y
    uint32_t *d_blob_avg = nullptr;        // [h][nchains] blob2pixel colour (AVERAGE) per frame/chain
    uint32_t *d_blob_distinct = nullptr;   // [nchains]   DISTINCT colour per chain (host mt19937(group))
    // per-pixel A-buffer of one render batch (amx_render.cu): counters, direct record slots, overflow lists
    uint32_t *ab_cnt = nullptr;            // two buffers (ping-pong), each [RBATCH][canvas]
    uint32_t  ab_parity = 0, ab_dirty[2] = {0, 0};
    void     *d_render_stats = nullptr;
    uint4    *ab_pair = nullptr, *ab_pair2 = nullptr;
    uint32_t *ab_cnt_base = nullptr;       // allocations behind ab_cnt / ab_pair (guard in front)
    uint4    *ab_pair_base = nullptr;
    uint32_t *ab_ovf_head = nullptr;
    uint4    *ab_ovf_rec = nullptr;        // overflow POOL: the records of a home beyond the direct slots, contiguous from ab_ovf_head[home]
    uint4    *ab_ovf_list = nullptr;       // overflow records in arrival order {colour, meta, atom, claim index} ...
    uint32_t *ab_ovf_list_home = nullptr;  // ... and their home (slot * canvas + position)
    uint32_t *ab_ovf_ctrl = nullptr;       // [2][2] {list entries, pool top}; ping-pong between scatters
    uint32_t  ab_ovf_parity = 0;
    uint2    *gl_items = nullptr;          // positions the gather leaves to k_resolve_list: {canvas index, batch slot}
    uint32_t *gl_count = nullptr;          // [2][2] {replays, heavy positions}; ping-pong: a batch's resolve kernel clears the other pair
    uint32_t  gl_cap = 0;
    uint32_t  render_batch = 8;            // frames per launch pair (clamped to RBATCH on the tiled path, GBATCH on the general one)
    // tiled path (amx_render.cu: Bins): record bins per (frame slot of a batch, 32x32-pixel tile)
    uint2    *tb_rec = nullptr;
    uint32_t *tb_atom = nullptr, *tb_chain = nullptr;
    uint32_t *tb_cnt = nullptr;            // two counter buffers (ping-pong), each [RBATCH][tiles][4]
    uint32_t *tb_flag = nullptr;           // raised by the kernels when a bin / tile overflows
    bool      tiled_acc = true;            // AMX_RENDER_ACC=0: single-chain morphs through the ordering kernel k_tile (for comparisons)
    uint32_t  tb_tiles_x = 0, tb_tiles_y = 0, tb_parity = 0, tb_dirty[2] = {0, 0};
    bool      tb_has_chain = false;
    bool      tiled_enabled = true;        // AMX_RENDER_TILED=0: general A-buffer path only (for comparisons)
    bool      tiled_multi = false;         // AMX_RENDER_TILED=2: tiled path for morphs with several chains as well
    bool      tiled_blocked = false;       // this call only: everything through the general path (re-render after a wrapped pixel)
    uint32_t  tiled_blocked_mask = 0;      // bit (interval & 31): a bin overflowed there with the current table -- general path for that
                                           // key-frame interval until the next refresh
    uint64_t  tiled_frames = 0, general_frames = 0;   // diagnostics (amx_render_path_frames)
    uint32_t  tb_demand[6] = {0, 0, 0, 0, 0, 0};       // largest bin counts per class, tile total, overflow list seen (amx_render_tiled_stats)
    uint64_t  tiled_fallbacks = 0;                    // render calls that were repeated on the general path
    uint32_t *d_bg = nullptr;              // background images of one batch (keep_background)
    size_t    d_bg_cap = 0;
    // per-(pixel, blob) entries for the feather / per-blob paths (canvas sized + overflow hash)
    int32_t  *acc_owner = nullptr;
    uint8_t  *acc_hasovf = nullptr;
    unsigned long long *ovf_key = nullptr;
    uint32_t  ovf_cap = 0;
    uint32_t *d_ovf_used = nullptr;
    uint32_t *blob_px = nullptr;           // resolved per-(pixel,layer-0) colour for feather / per-blob fetch
    uint32_t *d_out = nullptr;             // staging for host output
    uint32_t *d_pix = nullptr;             // staging for amx_render_pixels (RGBA + am::pixel records)
    size_t    d_pix_cap = 0;
    uint64_t  d_out_cap = 0;
    int32_t  *d_perlin = nullptr;          // [2][512] lag / slope permutation tables (host generated)
    unsigned  perlin_seed_loaded = 0xffffffffu;

    Fluid *fluid = nullptr;
    Dist  *dist = nullptr;                 // communicator, exchange buffers and peer mappings of the multi-GPU matcher

    bool fail(cudaError_t e, const char *what);
    bool check(const char *what);
    size_t canvas() const { return (size_t) cw * ch; }
};

// helpers implemented in amx_core.cu
bool dev_alloc(Engine *E, void **p, size_t bytes, const char *what);
void dev_free(void *p);
inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned) ((a + b - 1) / b); }

// stage entry points (each in its own .cu)
int engine_upload_convert(Engine *E, uint32_t index, const uint32_t *d_rgba_raw);           // amx_color.cu
int engine_blobify(Engine *E);                                                              // amx_ccl.cu
int engine_unify(Engine *E);                                                                // amx_ccl.cu
int engine_build_blob_pixels(Engine *E, uint32_t index);                                    // amx_ccl.cu
int engine_match_init(Engine *E);                                                           // amx_blobmatch.cu
int engine_match_rounds(Engine *E, uint64_t rounds);                                        // amx_blobmatch.cu
int engine_match_energy(Engine *E, double *e);                                              // amx_blobmatch.cu
int engine_init_chains(Engine *E);                                                          // amx_chain.cu
int engine_alloc_chains(Engine *E, uint32_t nchains, const uint64_t *keys, const uint64_t *widths,
                        const uint64_t *max_surface, uint32_t height);                      // amx_chain.cu
int engine_swap_rounds(Engine *E, int32_t chain, int32_t column, uint64_t rounds);          // amx_swap.cu
int engine_swap_tiled_epoch(Engine *E, uint32_t chain, uint32_t y, uint64_t epoch, uint32_t rounds, uint32_t rank, uint32_t nranks);
int engine_swap_local_epoch(Engine *E, uint32_t chain, uint32_t y, uint64_t epoch, uint32_t rounds);
int engine_cost(Engine *E, double *cost);                                                   // amx_swap.cu
int engine_render_prepare(Engine *E);                                                       // amx_render.cu
int engine_render(Engine *E, const double *times, uint32_t n, uint32_t *out, int out_is_device);  // amx_render.cu
int engine_render_blob(Engine *E, uint32_t blob, double t, uint64_t cap, uint16_t *xy, uint32_t *rgba, int64_t *n, uint64_t *group);
int engine_background(Engine *E, double t, uint32_t *out, int out_is_device);               // amx_render.cu
void engine_render_free(Engine *E);
int engine_render_fluid(Engine *E, double time, uint32_t f, double tl, const uint32_t *d_bg, uint32_t *d_dst);   // amx_fluiddraw.cu
void engine_fluid_free(Engine *E);
void engine_dist_table_gone(Engine *E);                                                     // amx_dist.cu
void engine_dist_free(Engine *E);

} // namespace amx

struct amx_ctx { amx::Engine e; };

#endif
