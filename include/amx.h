/*
 * amx.h -- C-ABI of libatomorph_b200.so: the sm_100a device engine underneath the am::morph
 * facade (include/atomorph/morph.h).  Plain C types, caller-owned host buffers, int status
 * codes, no exceptions, NO CPU FALLBACK: every entry point needs a CUDA device and fails with
 * AMX_ERR_CUDA when there is none.
 *
 * Each entry point names the reference interface it replaces (file:line under the reference
 * tree 1Hyena/atomorph).  A cgo / JNI / ctypes binding for this path would bind exactly these
 * symbols; INTEGRATION.md shows the stubs.
 *
 * Conventions
 *   - colours are packed little-endian  r | g<<8 | b<<16 | a<<24  (am::color, color.h:7-12)
 *   - key points are u64 words  x | y<<16 | x_fract<<32 | y_fract<<40 | flags<<48
 *     (am::point, atomorph.h:275-284); chain tables are COLUMN-major: word[j*width + x] is
 *     atom x at key frame j  (the reference's points[x][j], atomorph.h:286-298)
 *   - images are row-major, `canvas_w * canvas_h` for key-frame data, `width * height` for
 *     rendered output
 *   - frame INDEX = rank of the frame key in ascending key order (am::frame::index)
 */
#ifndef AMX_H
#define AMX_H

#include <stdint.h>
#include <stddef.h>
#include "amx_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amx_ctx amx_ctx;

enum amx_status {
    AMX_OK = 0,
    AMX_ERR_CUDA = 1,      /* no device / CUDA runtime error (see amx_last_error)        */
    AMX_ERR_ARG = 2,       /* bad argument                                                */
    AMX_ERR_STATE = 3,     /* call not valid in the current pipeline state                */
    AMX_ERR_NOMEM = 4,     /* device or host allocation failed                            */
    AMX_ERR_BUSY = 5
};

/* ---- lifetime ------------------------------------------------------------------------- */
/* replaces am::morph::morph() / ~morph() / clear()  (morph.cpp:9-50) on the device side    */
int  amx_create(amx_ctx **out, int device);
void amx_destroy(amx_ctx *ctx);
const char *amx_last_error(amx_ctx *ctx);
const char *amx_version(void);                                   /* am::get_version, atomorph.cpp:1013 */
/* Run every kernel / copy of this context on an externally owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream).
 * NULL (0) is taken literally: the legacy default stream.  AMX_STREAM_PRIVATE returns to the private non-blocking stream
 * the context was created with.  amx_get_stream returns the stream in use (cudaStream_t). */
#define AMX_STREAM_PRIVATE ((void *) (intptr_t) -1)
int  amx_set_stream(amx_ctx *ctx, void *cuda_stream);
void *amx_get_stream(amx_ctx *ctx);
int  amx_device_sync(amx_ctx *ctx);

/* ---- parameters: one id per setter, morph.h:52-76, pushed like synchronize() morph.cpp:105-120 */
int  amx_set_param(amx_ctx *ctx, int id, double value);
double amx_get_param(amx_ctx *ctx, int id);

/* ---- ingest (row a-I): replaces the worker-side set_frame, thread.cpp:89-111, and the
 *      RGB->HSP store / HSP->RGB fetch round trip of add_pixel / get_pixel (morph.cpp:298-301, 378-392).
 * amx_reset drops frames, blobs, chains (thread::clear, thread.cpp:36-81).
 * amx_set_canvas: width/height = set_resolution (morph.cpp:1517-1520); canvas covers every pixel
 * position; bbox = bbox_x1,y1,x2,y2 (morph.cpp:313-316).
 * amx_upload_frame: rgba = RAW input colours (as passed to add_pixel), present[i] != 0 where a
 * pixel was added; means = the facade's running means x,y,r,g,b,a (morph.cpp:318-334). */
int  amx_reset(amx_ctx *ctx);
int  amx_set_canvas(amx_ctx *ctx, uint32_t width, uint32_t height, uint32_t canvas_w, uint32_t canvas_h, const uint16_t bbox[4]);
int  amx_set_frame_count(amx_ctx *ctx, uint32_t nframes, const uint64_t *keys);
int  amx_upload_frame(amx_ctx *ctx, uint32_t index, const uint32_t *rgba, const uint8_t *present, const double means[6]);
/* device pointer variant (inputs already resident in HBM) */
int  amx_upload_frame_device(amx_ctx *ctx, uint32_t index, const uint32_t *d_rgba, const uint8_t *d_present, const double means[6]);
/* fetch colours after the store/fetch round trip (am::morph::get_pixel for every canvas position) */
int  amx_download_fetch(amx_ctx *ctx, uint32_t index, uint32_t *rgba_out);
int  amx_download_stored(amx_ctx *ctx, uint32_t index, uint32_t *rgba_out);

/* ---- pipeline state machine (row a-B..a-M): replaces thread::step, thread.cpp:139-171 ------- */
/* nsteps reference-equivalent steps: blob detection completes in one step; one matching step =
 * one parallel round of disjoint blob swaps; one morphing step = max(1,threads)*cycle_length atom
 * swap proposals on one chain, chains served round-robin (thread.cpp:1043-1064). */
int  amx_step(amx_ctx *ctx, uint64_t nsteps);
int  amx_next_state(amx_ctx *ctx);                               /* thread::next_state, thread.h:17 */
unsigned amx_get_state(amx_ctx *ctx);                            /* thread::get_state            */
double amx_get_energy(amx_ctx *ctx);                             /* thread::get_energy, thread.cpp:1176-1185 (true cost, not the reference's uninitialised sum) */

/* ---- K2 blob segmentation (row a-B1): thread::blobify_frame, thread.cpp:225-412 ------------- */
int  amx_blobify(amx_ctx *ctx);
int  amx_blob_count(amx_ctx *ctx, uint32_t index, uint32_t *count);
/* labels_out[canvas] = blob index in the frame's blob vector (-1 = no pixel);
 * stats_out[6*b..] = x,y,r,g,b,a ; meta_out[2*b..] = group, size */
int  amx_export_blobs(amx_ctx *ctx, uint32_t index, int32_t *labels_out, double *stats_out, uint64_t *meta_out);
int  amx_import_blobs(amx_ctx *ctx, uint32_t index, uint32_t nblobs, const int32_t *labels, const double *stats, const uint64_t *groups);

/* ---- K3 blob matching (row a-B3): thread::match, thread.cpp:598-738; blob_distance 1151-1174 -- */
int  amx_match_init(amx_ctx *ctx);
int  amx_match_rounds(amx_ctx *ctx, uint64_t rounds);
int  amx_match_energy(amx_ctx *ctx, double *energy);             /* thread::get_energy(blob***), thread.cpp:1087-1107 */

/* ---- K4 chain build (row a-C): thread::init_morph, thread.cpp:740-891 ------------------------ */
int  amx_init_chains(amx_ctx *ctx);
int  amx_chain_count(amx_ctx *ctx, uint32_t *count);
/* info4 = key(group), width, height, max_surface */
int  amx_chain_info(amx_ctx *ctx, uint32_t chain, uint64_t info4[4]);
int  amx_export_chain(amx_ctx *ctx, uint32_t chain, uint64_t *words_out);
/* Replace ALL chains: keys[n], widths[n], max_surface[n], height, words = chains concatenated, each column-major */
int  amx_import_chains(amx_ctx *ctx, uint32_t nchains, const uint64_t *keys, const uint64_t *widths,
                       const uint64_t *max_surface, uint32_t height, const uint64_t *words);
/* device access for collectives (NCCL all-gather of table columns): pointer to column j of the
 * concatenated table (total_atoms words) */
int  amx_table_device_ptr(amx_ctx *ctx, uint32_t column, void **d_ptr, uint64_t *total_atoms);

/* ---- K1 pair-swap optimal transport (row a-M): morph_asynch, thread.cpp:990-1041 -------------- */
/* `rounds` rounds of disjoint pairings on column `column` (or a counter-RNG chosen column when
 * column < 0) of chain `chain` (or all chains when chain < 0).  stats3 += proposals, accepted, gain. */
int  amx_swap_rounds(amx_ctx *ctx, int32_t chain, int32_t column, uint64_t rounds, uint64_t stats3[3]);
int  amx_swap_stats(amx_ctx *ctx, uint64_t stats3[3]);           /* cumulative since init / import  */
/* Multi-GPU atom-range sharding of one column: only atoms whose index bits under
 * `sel_mask` equal `sel_val` are paired, with pairing masks that are zero on those bits.  amx_pack_owned copies
 * the owned atoms of a column into a contiguous device buffer (count = 2^(free bits)), amx_unpack_owned scatters
 * the all-gathered buffers of `nranks` = 2^popcount(sel_mask) ranks (rank-major) back into the column. */
int  amx_swap_rounds_sharded(amx_ctx *ctx, uint32_t chain, int32_t column, uint64_t rounds, uint64_t sel_mask, uint64_t sel_val);
int  amx_pack_owned(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t sel_mask, uint64_t sel_val, void *d_out, uint64_t *count);
int  amx_unpack_owned(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t sel_mask, uint32_t nranks, const void *d_in);
/* Shared-memory tiled path (long single chains; amx_swap_rounds uses it by itself on one GPU).  An EPOCH spreads the
 * atoms of `chain` over tiles of 2048 through a bijection derived from (seed, chain, epoch) and runs `rounds` rounds
 * inside each tile (64 per launch).  Multi-GPU: rank r of n runs its contiguous share of the tiles; amx_pack_tiled copies the
 * slots the rank owns into a contiguous device buffer, the ranks all-gather those buffers (equal sizes when n divides
 * the tile count) and amx_unpack_tiled scatters the gathered slots (rank-major = slot order) back into the column. */
int  amx_swap_tiled_epoch(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rounds, uint32_t rank, uint32_t nranks);
/* Locality-biased proposals (extension, default off; the intent of the reference's experimental kd-tree variant,
 * thread.cpp:893-986): a locality epoch sorts the atoms by the Morton code of their current column position and refines
 * runs of 1024 spatial neighbours in shared memory.  amx_set_swap_locality(ctx, n) makes every n-th epoch of
 * amx_swap_rounds / amx_step a locality epoch (0 = never: the counted metric uses uniform partners only). */
int  amx_swap_local_epoch(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rounds);
int  amx_set_swap_locality(amx_ctx *ctx, uint32_t every);
int  amx_pack_tiled(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, uint32_t rank, uint32_t nranks, void *d_out, uint64_t *count);
int  amx_unpack_tiled(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t epoch, const void *d_in);
int  amx_cost(amx_ctx *ctx, double *cost);                       /* thread::get_energy(chain*), thread.cpp:1109-1125, summed over chains */

/* ---- multi-GPU matcher: one process (context) per GPU; the all-gather of trajectory-table columns
 *      is the only collective of the path and is issued by the library itself, on the context's stream -----------------
 * amx_comm_unique_id: rank 0 creates the id, the caller hands it to every rank (any transport: MPI, torch.distributed, a
 * file).  amx_comm_init joins the NCCL communicator (libnccl.so.2 is dlopen'ed on first use).  amx_comm_enable_p2p
 * (collective, after every rank holds a table of the same geometry) maps the table replicas into each other with
 * cudaIpc: from then on a step's refined tiles are written into every replica from inside the swap kernel over
 * NVLink and a flag barrier in peer memory closes the step -- no collective call at all.  It returns AMX_ERR_STATE and
 * leaves the NCCL exchange in place when peer mapping is not permitted.  Replacing the table (amx_import_chains with
 * another geometry, amx_init_chains, amx_reset) drops the mappings; call amx_comm_enable_p2p again.
 * amx_comm_info: [0] rank, [1] ranks, [2] 1 when the P2P exchange is active.  amx_comm_check: AMX_ERR_STATE when a peer
 * never arrived at a flag barrier (its process died): the replicas may differ from then on. */
#define AMX_UNIQUE_ID_BYTES 128
int  amx_comm_unique_id(uint8_t id[AMX_UNIQUE_ID_BYTES]);
int  amx_comm_init(amx_ctx *ctx, const uint8_t id[AMX_UNIQUE_ID_BYTES], uint32_t rank, uint32_t nranks);
int  amx_comm_destroy(amx_ctx *ctx);
int  amx_comm_enable_p2p(amx_ctx *ctx);
int  amx_comm_disable_p2p(amx_ctx *ctx);
int  amx_comm_info(amx_ctx *ctx, uint32_t info3[3]);
int  amx_comm_check(amx_ctx *ctx);
/* rank `root`'s table replaces everybody's (collective; before rendering / refining the same morph on every GPU) */
int  amx_table_broadcast(amx_ctx *ctx, uint32_t root);
/* h == 2 (a single free column): one STEP of the sharded matcher on `column` of `chain` -- morph_asynch, thread.cpp:990-1041,
 * over the N ranks.  A bijection of the atom index drawn from (seed, chain, step) gives rank r a pseudo-random 1/N of the
 * atoms; the rank refines its part for `sub_epochs` epochs (each re-tiling the part) of `rounds` (<= 64) rounds, then the
 * parts are exchanged (P2P write-through or pack -> ncclAllGather -> unpack).  Same call, same arguments on every rank.
 * With one rank (no communicator) it is `sub_epochs` ordinary epochs. */
int  amx_swap_part_step(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t step, uint32_t sub_epochs, uint32_t rounds);
/* h >= 3: the key-frame columns of one PHASE are refined concurrently, every N-th one on this rank, their neighbour
 * columns frozen (the objective couples column j to j-1 and j+1 only, thread.cpp:1007-1020); then the refined columns go
 * to every replica.  Phases: 0 = even columns, 1 = odd columns, 2 = the last column of an odd cycle (it neighbours
 * column 0); amx_swap_phase_count = 2 or 3.  chain < 0: every chain.  `epochs` x `rounds` rounds per owned column. */
int  amx_swap_columns_step(amx_ctx *ctx, int32_t chain, uint32_t phase, uint64_t step, uint32_t epochs, uint32_t rounds);
uint32_t amx_swap_phase_count(amx_ctx *ctx);
/* invariants of a sharded step: out2[0] = position-dependent hash of a column (equal on two ranks <=> same column),
 * out2[1] = position-independent hash (unchanged <=> still the same multiset of key points, thread.cpp:1023-1038) */
int  amx_column_hash(amx_ctx *ctx, uint32_t column, uint64_t out2[2]);

/* ---- K6 renderer (row a-R): morph::get_pixels / draw_atoms, morph.cpp:452-678, 1302-1421 ------- */
/* Refresh the per-atom render inputs from the chain table (the device analogue of the chain /
 * spline mirror of synchronize(), morph.cpp:174-205). */
int  amx_render_prepare(amx_ctx *ctx);
/* n frames at times t[i] -> out[i*width*height ...] packed RGBA.  out_is_device: out is a device pointer. */
int  amx_render(amx_ctx *ctx, const double *times, uint32_t n, uint32_t *out, int out_is_device);
/* one frame as width*height am::pixel records {u16 x, u16 y, u8 r, g, b, a} in row-major order: exactly what
 * morph::get_pixels(double, std::vector<pixel>*) returns (morph.cpp:1405-1421), built on the device so that the
 * host side is one copy into the caller's vector */
int  amx_render_pixels(amx_ctx *ctx, double t, uint64_t *pixels_out);
/* amx_render_pixels looks ahead: a call that is not served from its ring renders the requested frame and the next ones at
 * the caller's stride (times predicted as (f + k) / N, only when that reproduces the last two requests bit by bit) as one
 * batch into pinned host slots; later calls are served from there.  Served frames are bit-identical to direct renders; any
 * change of table, parameters or resolution empties the ring.  On by default (off: amx_set_lookahead(ctx, 0) or
 * AMX_LOOKAHEAD=0; fluid frames and frames above 128 MiB never use it).  stats2: [0] calls served from the ring, [1] misses. */
int  amx_set_lookahead(amx_ctx *ctx, int enable);
int  amx_lookahead_stats(amx_ctx *ctx, uint64_t stats2[2]);
/* renderer diagnostics, cumulative since the render buffers were (re)built: [0] pixels resolved by the ordered double
 * replay of morph.cpp:598-613 or left to the list kernels (three blobs or more, more than 32 records) instead of the fused
 * gather's exact integer sums, [1] of those the exact .5 ties, [2] A-buffer records beyond the four direct slots of their
 * home pixel (they go to the home's range of the overflow pool) */
int  amx_render_stats(amx_ctx *ctx, uint64_t stats3[3]);
/* device time of the render kernels of every batch ([0] k_bin2, or k_scatter + k_ovf_alloc + k_ovf_place; [1] k_acc (k_tile), or
 * k_gather_pixel + k_resolve), measured
 * with CUDA event pairs on the engine's stream without synchronising between launches.  Returns the milliseconds,
 * launches and frames accumulated since the previous call, then switches the recording on (enable != 0) or off.
 * Process-wide (one engine per process records). */
int  amx_kernel_times(amx_ctx *ctx, int enable, double ms2[2], uint64_t launches2[2], uint64_t frames2[2]);
/* frames rendered so far by [0] the tiled path (shared-memory tiles, feather == 0 without fluid) and [1] the general
 * A-buffer path (feather, per-blob fetch, several chains, or more than 7 atoms per pixel over a 32x32 tile); both are exact */
int  amx_render_path_frames(amx_ctx *ctx, uint64_t frames2[2]);
/* tiled-path diagnostics: [0..3] largest record count seen in a bin of class interior / last column / last row / corner,
 * [4] largest record total of a tile, [5] unused (0), [6] render calls repeated on the general path,
 * [7] 1 while a key-frame interval is kept off the tiled path for the current table (a bin overflowed there; the other
 * intervals stay on it) */
int  amx_render_tiled_stats(amx_ctx *ctx, uint64_t stats8[8]);
/* one blob of the frame active at time t (morph::get_pixels(size_t,double,vector*), morph.cpp:452-678):
 * returns count via *n (pixels in reference emission order), -1 in *n when the blob index is out of range */
int  amx_render_blob(amx_ctx *ctx, uint32_t blob, double t, uint64_t cap, uint16_t *xy_out, uint32_t *rgba_out, int64_t *n, uint64_t *group);
/* background cross-dissolve image at time t (morph::get_background for every pixel, morph.cpp:1431-1465) */
int  amx_background(amx_ctx *ctx, double t, uint32_t *out, int out_is_device);

/* ---- K5 fluid (row a-F): FluidModel::step, fluidmodel.cpp:165-580 ------------------------------ */
/* particle record = AMX_FP_STRIDE doubles:
 * 0 x 1 y 2 u 3 v 4 gravity_x 5 gravity_y 6 freedom_r 7 active 8 mature 9 R 10 G 11 B 12 A 13 r 14 g 15 b 16 a
 * 17 strength 18 source_owner 19 frame_key 20 source_pos 21 destination_pos 22 cx 23 cy */
#define AMX_FP_STRIDE 24
int  amx_fluid_create(amx_ctx *ctx, uint32_t gsize_x, uint32_t gsize_y, uint32_t particle_count);
int  amx_fluid_set_particles(amx_ctx *ctx, uint32_t n, const double *records);
int  amx_fluid_get_particles(amx_ctx *ctx, uint32_t n, double *records);
int  amx_fluid_step(amx_ctx *ctx, uint64_t steps_left, double freedom_radius, double t);
/* node record = 13 doubles m d gx gy u v ax ay r g b a weight ; out[(j*gsize_x+i)*13+k] */
int  amx_fluid_get_nodes(amx_ctx *ctx, double *out);

/* ---- measurement helpers -------------------------------------------------------------------- */
/* number of kernels this context launched since creation (bench.py "gpu_launches") */
uint64_t amx_launch_count(amx_ctx *ctx);
/* bracket a region with CUDA events on the engine's stream; amx_timer_stop returns milliseconds */
int  amx_timer_start(amx_ctx *ctx);
int  amx_timer_stop(amx_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif
