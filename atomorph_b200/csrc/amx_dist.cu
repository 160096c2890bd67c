/*
 * amx_dist.cu -- the multi-GPU matcher: one process per GPU, every collective of the path is
 * issued HERE, on the engine's own stream, so that it is ordered with the kernels around it by construction.
 *
 * What is exchanged is always the same thing: key-frame columns of the per-atom trajectory table (W * 8 B per column),
 * the only collective of the path.
 *
 *   h == 2 (BASELINE config 2): a single free column -> the ATOMS are split.  A STEP draws one bijection of the atom
 *     index (identical on every rank: it is a pure function of seed, chain, step); rank r owns the contiguous slot range
 *     [r 2^k / N, (r + 1) 2^k / N) of it, i.e. a pseudo-random 1/N of the atoms, and refines it for `sub_epochs` epochs of
 *     the shared-memory tiled kernel (amx_swap.cu), each epoch re-tiling the part through an inner bijection.  Pairs never
 *     leave a part, so no rank reads or writes another rank's atoms during a step; over steps every pair of atoms meets
 *     with equal probability, like the reference's uniform draw (thread.cpp:1002-1005).
 *   h >= 3 (BASELINE config 5): the objective couples column j only to j-1 and j+1 (thread.cpp:1007-1020), so columns of
 *     one PHASE (even / odd / the last column of an odd cycle) are refined concurrently with their neighbours frozen:
 *     rank g owns every N-th column of the phase ("partitioned by key-frame pair").
 *
 * The exchange has two implementations behind the same entry points:
 *   P2P   (amx_comm_enable_p2p): the table replicas and a flag block of every rank are mapped into every other rank with
 *         cudaIpc*; the LAST epoch of a step writes each refined tile straight into all replicas from inside k_swap_tiled
 *         (PeerCols) -- compute and transfer overlap tile by tile over NVLink / NVSwitch -- and a flag barrier in peer memory
 *         (k_peer_barrier, release/acquire at system scope) closes the step.  No pack, no collective call, no unpack.
 *   NCCL  (default until P2P is enabled, and the fallback when IPC mapping is not permitted): pack -> ncclAllGather ->
 *         unpack for parts, in-place ncclBroadcast per column for columns.
 * libnccl.so.2 is dlopen'ed on first use: a single-GPU user of the library never needs it.
 */
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "amx_engine.h"
#include "amx_swap.h"

namespace amx {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
static NcclApi g_nccl;

static bool nccl_load() {
    if (g_nccl.lib) return true;
    // a process that already carries an NCCL (e.g. torch's bundled one) gets that one: the loader matches the soname
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { g_nccl.err = std::string("libnccl.so.2 not found: ") + dlerror(); return false; }
#define AMX_SYM(field, name) do { *(void **) (&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { g_nccl.err = std::string("missing NCCL symbol ") + name; dlclose(lib); return false; } } while (0)
    AMX_SYM(GetUniqueId, "ncclGetUniqueId"); AMX_SYM(CommInitRank, "ncclCommInitRank"); AMX_SYM(CommDestroy, "ncclCommDestroy");
    AMX_SYM(AllGather, "ncclAllGather"); AMX_SYM(Broadcast, "ncclBroadcast"); AMX_SYM(AllReduce, "ncclAllReduce");
    AMX_SYM(GroupStart, "ncclGroupStart"); AMX_SYM(GroupEnd, "ncclGroupEnd"); AMX_SYM(GetErrorString, "ncclGetErrorString");
#undef AMX_SYM
    g_nccl.lib = lib;
    return true;
}

struct Dist {
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, nranks = 1;
    pword *send = nullptr, *recv = nullptr;       // NCCL path: packed part / gathered parts
    size_t cap = 0;
    // P2P
    bool p2p = false;
    pword *p2p_table = nullptr;                   // the local table the mappings belong to
    pword *peer_table[AMX_MAX_PEERS + 1] = {};    // by rank (own entry = local table)
    pword *stage = nullptr;                       // [2][stage_cap] staging buffers the peers write their parts into (slot order), by step parity
    pword *peer_stage[AMX_MAX_PEERS + 1] = {};
    size_t stage_cap = 0;
    unsigned long long *flags = nullptr;          // [nranks] arrival counters written by the peers
    unsigned long long *peer_flags[AMX_MAX_PEERS + 1] = {};
    unsigned long long **d_peer_flags = nullptr;  // device copy of peer_flags
    uint32_t *d_timeout = nullptr;                // raised by k_peer_barrier when a peer never arrived
    unsigned long long barrier_seq = 0;
    unsigned long long *d_hash = nullptr;
};

static bool nccl_fail(Engine *E, ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return false;
    E->err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return true;
}

static void drop_p2p(Engine *E) {
    Dist *D = E->dist;
    if (!D) return;
    for (uint32_t r = 0; r < D->nranks && r <= AMX_MAX_PEERS; ++r) {
        if (r == D->rank) continue;
        if (D->peer_table[r]) cudaIpcCloseMemHandle(D->peer_table[r]);
        if (D->peer_flags[r]) cudaIpcCloseMemHandle(D->peer_flags[r]);
        if (D->peer_stage[r]) cudaIpcCloseMemHandle(D->peer_stage[r]);
    }
    memset(D->peer_stage, 0, sizeof D->peer_stage);
    memset(D->peer_table, 0, sizeof D->peer_table);
    memset(D->peer_flags, 0, sizeof D->peer_flags);
    D->p2p = false; D->p2p_table = nullptr;
}

// the chain table is about to be freed: peers' views of it are meaningless from now on (amx_chain.cu, amx_core.cu)
void engine_dist_table_gone(Engine *E) {
    if (E->dist && E->dist->p2p) { cudaStreamSynchronize(E->stream); drop_p2p(E); }
}

void engine_dist_free(Engine *E) {
    Dist *D = E->dist;
    if (!D) return;
    cudaStreamSynchronize(E->stream);
    drop_p2p(E);
    if (D->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(D->comm);
    dev_free(D->send); dev_free(D->recv); dev_free(D->stage); dev_free(D->flags); dev_free(D->d_peer_flags); dev_free(D->d_timeout); dev_free(D->d_hash);
    delete D;
    E->dist = nullptr;
}

// ---- flag barrier in peer memory -------------------------------------------------------------------------------------
// Thread p signals rank p (its arrival counter of this rank := seq) and waits until rank p has signalled this rank.
// The kernels before it in the stream have completed, so their peer stores are performed; the release store orders them
// before the flag for the peer's acquiring load.  A peer that never arrives (a crashed process) must not hang the GPU:
// after ~10 s of spinning the kernel raises `timeout` and returns.
__global__ void k_peer_barrier(unsigned long long *const *__restrict__ peer_flags, volatile unsigned long long *my_flags, uint32_t rank, uint32_t n,
                               unsigned long long seq, uint32_t *timeout) {
    const uint32_t p = threadIdx.x;
    if (p >= n || p == rank) return;
    __threadfence_system();
    unsigned long long *dst = peer_flags[p] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(dst), "l"(seq) : "memory");
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + p) : "memory");
        if (v >= seq) break;
        if (clock64() - t0 > 20000000000ll) { atomicExch(timeout, 1u); break; }
        __nanosleep(200);
    }
}

static int peer_barrier(Engine *E) {
    Dist *D = E->dist;
    ++D->barrier_seq;
    k_peer_barrier<<<1, 32, 0, E->stream>>>(D->d_peer_flags, D->flags, D->rank, D->nranks, D->barrier_seq, D->d_timeout);
    E->launches++;
    return E->check("peer barrier") ? AMX_ERR_CUDA : AMX_OK;
}

static PeerCols peer_cols(Engine *E, uint32_t column) {
    PeerCols pc; pc.n = 0; pc.staged = 0;
    Dist *D = E->dist;
    if (!D || !D->p2p || D->p2p_table != E->table) return pc;
    for (uint32_t r = 0; r < D->nranks; ++r)
        if (r != D->rank) pc.dst[pc.n++] = D->peer_table[r] + (size_t) column * E->A;
    return pc;
}
// the peers' staging buffers of parity `par` (a rank's part lands at the slot indices it owns)
static PeerCols peer_stages(Engine *E, uint32_t par) {
    PeerCols pc; pc.n = 0; pc.staged = 1;
    Dist *D = E->dist;
    for (uint32_t r = 0; r < D->nranks; ++r)
        if (r != D->rank) pc.dst[pc.n++] = D->peer_stage[r] + (size_t) par * D->stage_cap;
    return pc;
}

static bool p2p_ready(Engine *E) { return E->dist && E->dist->p2p && E->dist->p2p_table == E->table; }

// column y of the local table -> every replica (short chains, whose kernels do not write through themselves)
__global__ void __launch_bounds__(256)
k_push_column(const pword *__restrict__ col, uint64_t n, PeerCols peers) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const pword w = col[i];
        for (uint32_t p = 0; p < peers.n; ++p) peers.dst[p][i] = w;
    }
}

// ---- hashes for the invariants a sharded step must keep --------------------------------------------------------------
// out[0]: position-dependent hash of the column (equal on two ranks <=> same table, w.h.p.)
// out[1]: position-independent hash (unchanged by a step <=> the column is still the same multiset of key points)
__global__ void __launch_bounds__(256)
k_column_hash(const pword *__restrict__ col, uint64_t n, unsigned long long *__restrict__ out) {
    unsigned long long a = 0, b = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t w = col[i];
        a += mix64(w ^ mix64(i + 0x1234567ull));
        b += mix64(w + 0x9e3779b97f4a7c15ull);
    }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}

// tile size (log2 atoms) for a part of 2^kk slots.  The tiled kernel is latency-bound (one barrier per round): it needs
// about six resident CTAs per SM to hide it, so a smaller part is cut into smaller tiles (1024 -> 512 -> 256 atoms)
// rather than into fewer CTAs.  Measured at N = 2 with 1024-atom tiles: half the tiles took the time of all of them.
static int pick_tile_bits(Engine *E, unsigned kk) {
    for (int tb = TILE_BITS; tb > 8; --tb)
        if (kk >= (unsigned) tb && (1ull << (kk - tb)) >= 6ull * (uint64_t) E->sm_count) return tb;
    return kk >= 8 ? 8 : -1;
}

static unsigned ilog2(uint32_t v) { unsigned s = 0; while ((1u << s) < v) ++s; return s; }

// ---- h == 2: one step on the rank's part of column y ------------------------------------------------------------------
int engine_swap_part_step(Engine *E, uint32_t chain, uint32_t y, uint64_t step, uint32_t sub_epochs, uint32_t rounds) {
    if (chain >= E->nchains || y >= E->h || E->h < 2 || rounds == 0 || rounds > TILE_MAX_ROUNDS || sub_epochs == 0) return AMX_ERR_ARG;
    Dist *D = E->dist;
    const uint32_t N = D ? D->nranks : 1u, rank = D ? D->rank : 0u;
    const uint64_t off = E->chain_off[chain], w = E->chain_off[chain + 1] - off;
    if (w < 2 || w > 0x80000000ull) return AMX_ERR_ARG;
    const unsigned k = ceil_log2(w), s = ilog2(N);
    if ((1u << s) != N || k < s + 8u) { E->err = "sharded step: the rank count must be a power of two and every part at least 256 atoms"; return AMX_ERR_ARG; }
    const unsigned kk = k - s;
    const int tb = pick_tile_bits(E, kk);
    if (tb < 0) return AMX_ERR_ARG;
    const uint32_t ntl = 1u << (kk - (unsigned) tb), tile0 = rank << (kk - (unsigned) tb);
    TileMap tm = make_tilemap(E->p.seed, 0x100u + chain, step, k);
    const uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
    pword *col = E->table + (size_t) y * E->A;
    const pword *prev = E->table + (size_t) yp * E->A, *next = E->table + (size_t) yn * E->A;
    const bool h2 = E->h == 2;
    const bool p2p = N > 1 && p2p_ready(E) && D->stage && D->stage_cap >= ((size_t) 1 << k);
    const uint32_t par = (uint32_t) (step & 1ull);
    PeerCols nopeers; nopeers.n = 0; nopeers.staged = 0;
    for (uint32_t j = 0; j < sub_epochs; ++j) {
        tilemap_set_inner(tm, E->p.seed, 0x100u + chain, step * 4096ull + j, kk);
        const bool last = j + 1 == sub_epochs;
        // the last sub-epoch also writes every refined tile, as one contiguous run, into the staging buffer of every peer
        launch_swap_tiled(E, h2, tb, col, prev, next, off, (uint32_t) w, tm, tile0, ntl, rounds, (step << 24) + ((uint64_t) j << 8),
                          (last && p2p) ? peer_stages(E, par) : nopeers);
    }
    E->render_ready = false;
    if (E->check("sharded swap step")) return AMX_ERR_CUDA;
    if (N == 1) return AMX_OK;
    if (p2p) {
        // everybody's tiles have arrived once the flag barrier is passed; the slots of the other ranks go to their atoms (the
        // slot -> atom map of the step's LAST sub-epoch, identical on every rank).  Staging buffers alternate with the step
        // parity, so a fast peer's next step never lands in the buffer this rank is still reading.
        int rc = peer_barrier(E);
        if (rc != AMX_OK) return rc;
        const uint32_t n = 1u << kk;
        const pword *stage = D->stage + (size_t) par * D->stage_cap;
        launch_unpack_tiled(E, col, off, (uint32_t) w, tm, 0u, n * N, rank * n, (rank + 1u) * n, stage);     // one launch: every slot but the own part
        return E->check("sharded swap unpack") ? AMX_ERR_CUDA : AMX_OK;
    }
    // NCCL: the part is a contiguous slot range of the OUTER bijection -> pack, all-gather, unpack, all on E->stream
    if (!D->comm) { E->err = "sharded step: amx_comm_init first"; return AMX_ERR_STATE; }
    const size_t n = (size_t) 1 << kk;
    if (D->cap < n * N) {
        dev_free(D->send); dev_free(D->recv); D->send = D->recv = nullptr; D->cap = 0;
        if (!dev_alloc(E, (void **) &D->send, n * 8, "dist send") || !dev_alloc(E, (void **) &D->recv, n * N * 8, "dist recv")) return AMX_ERR_NOMEM;
        D->cap = n * N;
    }
    tm.imask = 0u;                                   // slots of the outer bijection
    launch_pack_tiled(E, col, off, (uint32_t) w, tm, (uint32_t) (rank * n), (uint32_t) n, D->send);
    if (nccl_fail(E, g_nccl.AllGather(D->send, D->recv, n, ncclUint64, D->comm, E->stream), "ncclAllGather")) return AMX_ERR_CUDA;
    launch_unpack_tiled(E, col, off, (uint32_t) w, tm, 0u, (uint32_t) (n * N), (uint32_t) (rank * n), (uint32_t) ((rank + 1u) * n), D->recv);
    return E->check("sharded swap exchange") ? AMX_ERR_CUDA : AMX_OK;
}

// ---- h >= 3: the columns of one phase, every N-th one on this rank -----------------------------------------------------
// phases: 0 = even columns (without the last column of an odd cycle, which neighbours column 0), 1 = odd columns,
// 2 = that last column (odd h only)
uint32_t dist_phase_count(uint32_t h) { return h < 2 ? 0u : (h & 1u) ? 3u : 2u; }
void dist_phase_columns(uint32_t h, uint32_t phase, std::vector<uint32_t> &cols) {
    cols.clear();
    if (phase == 2) { if ((h & 1u) && h >= 3) cols.push_back(h - 1); return; }
    for (uint32_t j = phase; j < h; j += 2)
        if (!(phase == 0 && (h & 1u) && h >= 3 && j == h - 1)) cols.push_back(j);
}

int engine_swap_columns_step(Engine *E, int32_t chain, uint32_t phase, uint64_t step, uint32_t epochs, uint32_t rounds) {
    if (E->h < 2 || phase >= dist_phase_count(E->h) || rounds == 0 || epochs == 0 || chain >= (int32_t) E->nchains) return AMX_ERR_ARG;
    Dist *D = E->dist;
    const uint32_t N = D ? D->nranks : 1u, rank = D ? D->rank : 0u;
    std::vector<uint32_t> cols;
    dist_phase_columns(E->h, phase, cols);
    const bool p2p = N > 1 && p2p_ready(E);
    const bool h2 = E->h == 2;
    PeerCols nopeers; nopeers.n = 0; nopeers.staged = 0;
    for (size_t i = 0; i < cols.size(); ++i) {
        if (i % N != rank) continue;
        const uint32_t y = cols[i];
        const uint32_t yn = (y + 1) % E->h, yp = (y + E->h - 1) % E->h;
        pword *col = E->table + (size_t) y * E->A;
        const pword *prev = E->table + (size_t) yp * E->A, *next = E->table + (size_t) yn * E->A;
        bool pushed = false;
        const uint32_t c0 = chain >= 0 ? (uint32_t) chain : 0u, c1 = chain >= 0 ? (uint32_t) chain + 1u : E->nchains;
        for (uint32_t c = c0; c < c1; ++c) {
            const uint64_t off = E->chain_off[c], w = E->chain_off[c + 1] - off;
            if (w < 2) continue;
            if (tiled_ok(E, c) && rounds <= TILE_MAX_ROUNDS) {
                const unsigned k = ceil_log2(w);
                const int tb = pick_tile_bits(E, k);
                for (uint32_t e = 0; e < epochs; ++e) {
                    TileMap tm = make_tilemap(E->p.seed, 0x200u + c, (step * 64ull + y) * 4096ull + e, k);
                    // (whole columns travel as one contiguous copy below: scattered 8-byte peer stores from inside the kernel
                    // would fill NVLink packets to a quarter)
                    launch_swap_tiled(E, h2, tb, col, prev, next, off, (uint32_t) w, tm, 0u, 1u << (k - (unsigned) tb), rounds,
                                      (step << 24) + ((uint64_t) y << 16) + ((uint64_t) e << 8), nopeers);
                }
            } else {
                int rc = engine_swap_rounds(E, (int32_t) c, (int32_t) y, (uint64_t) epochs * rounds);
                if (rc != AMX_OK) return rc;
            }
        }
        if (p2p && !pushed) {
            k_push_column<<<std::min<uint32_t>(div_up(E->A, 256), (uint32_t) E->sm_count * 8u), 256, 0, E->stream>>>(col, E->A, peer_cols(E, y));
            E->launches++;
        }
    }
    E->render_ready = false;
    if (E->check("column swap step")) return AMX_ERR_CUDA;
    if (N == 1) return AMX_OK;
    if (p2p) return peer_barrier(E);
    if (!D->comm) { E->err = "column step: amx_comm_init first"; return AMX_ERR_STATE; }
    // every column of the phase travels from its owner to everybody, in place
    if (nccl_fail(E, g_nccl.GroupStart(), "ncclGroupStart")) return AMX_ERR_CUDA;
    for (size_t i = 0; i < cols.size(); ++i) {
        pword *col = E->table + (size_t) cols[i] * E->A;
        if (nccl_fail(E, g_nccl.Broadcast(col, col, E->A, ncclUint64, (int) (i % N), D->comm, E->stream), "ncclBroadcast")) { g_nccl.GroupEnd(); return AMX_ERR_CUDA; }
    }
    if (nccl_fail(E, g_nccl.GroupEnd(), "ncclGroupEnd")) return AMX_ERR_CUDA;
    return AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_comm_unique_id(uint8_t id[AMX_UNIQUE_ID_BYTES]) {
    if (!id) return AMX_ERR_ARG;
    if (!nccl_load()) return AMX_ERR_STATE;
    static_assert(sizeof(ncclUniqueId) <= AMX_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) return AMX_ERR_CUDA;
    memset(id, 0, AMX_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return AMX_OK;
}

int amx_comm_init(amx_ctx *ctx, const uint8_t id[AMX_UNIQUE_ID_BYTES], uint32_t rank, uint32_t nranks) {
    if (!ctx || !id || nranks == 0 || rank >= nranks || nranks > AMX_MAX_PEERS + 1) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    if (!nccl_load()) { E->err = g_nccl.err; return AMX_ERR_STATE; }
    engine_dist_free(E);
    Dist *D = new Dist();
    D->rank = rank; D->nranks = nranks;
    E->dist = D;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    if (nccl_fail(E, g_nccl.CommInitRank(&D->comm, (int) nranks, u, (int) rank), "ncclCommInitRank")) { D->comm = nullptr; engine_dist_free(E); return AMX_ERR_CUDA; }
    if (!dev_alloc(E, (void **) &D->d_hash, 16, "hash")) return AMX_ERR_NOMEM;
    return AMX_OK;
}

int amx_comm_destroy(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    engine_dist_free(&ctx->e);
    return AMX_OK;
}

int amx_comm_info(amx_ctx *ctx, uint32_t info3[3]) {
    if (!ctx || !info3) return AMX_ERR_ARG;
    Dist *D = ctx->e.dist;
    info3[0] = D ? D->rank : 0u; info3[1] = D ? D->nranks : 1u; info3[2] = (D && D->p2p && D->p2p_table == ctx->e.table) ? 1u : 0u;
    return AMX_OK;
}

// Collective.  Maps every rank's table and flag block into every other rank (cudaIpc handles travel through one
// ncclAllGather).  AMX_ERR_STATE (with the NCCL path still usable) when this box does not allow it.
int amx_comm_enable_p2p(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    Dist *D = E->dist;
    cudaSetDevice(E->device);
    if (!D || !D->comm) { E->err = "amx_comm_init first"; return AMX_ERR_STATE; }
    if (!E->table) { E->err = "no chain table to share"; return AMX_ERR_STATE; }
    if (D->nranks == 1) return AMX_OK;
    cudaStreamSynchronize(E->stream);
    drop_p2p(E);
    const uint32_t N = D->nranks;
    if (!D->flags) {
        if (!dev_alloc(E, (void **) &D->flags, 64 * sizeof(unsigned long long), "peer flags") ||
            !dev_alloc(E, (void **) &D->d_peer_flags, (AMX_MAX_PEERS + 1) * sizeof(void *), "peer flag pointers") ||
            !dev_alloc(E, (void **) &D->d_timeout, 4, "barrier timeout")) return AMX_ERR_NOMEM;
        cudaMemset(D->flags, 0, 64 * sizeof(unsigned long long));
        cudaMemset(D->d_timeout, 0, 4);
    }
    // staging for the parts of a step (h == 2): two buffers of 2^k slots for the widest chain
    if (E->h == 2) {
        uint64_t wmax = 2;
        for (uint32_t c = 0; c < E->nchains; ++c) wmax = std::max<uint64_t>(wmax, E->chain_off[c + 1] - E->chain_off[c]);
        const size_t cap = (size_t) 1 << ceil_log2(wmax);
        if (D->stage_cap < cap) {
            dev_free(D->stage); D->stage = nullptr; D->stage_cap = 0;
            if (!dev_alloc(E, (void **) &D->stage, 2 * cap * sizeof(pword), "peer staging")) return AMX_ERR_NOMEM;
            D->stage_cap = cap;
        }
    }
    // handles: [table | flags | stage | ok] per rank
    struct Rec { cudaIpcMemHandle_t table, flags, stage; uint64_t ok, seq, stage_cap; };
    Rec mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = (cudaIpcGetMemHandle(&mine.table, E->table) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, D->flags) == cudaSuccess &&
               (!D->stage || cudaIpcGetMemHandle(&mine.stage, D->stage) == cudaSuccess)) ? 1u : 0u;
    mine.stage_cap = D->stage ? D->stage_cap : 0;
    mine.seq = D->barrier_seq;
    cudaGetLastError();
    Rec *d_recs = nullptr;
    if (!dev_alloc(E, (void **) &d_recs, sizeof(Rec) * (N + 1), "ipc handles")) return AMX_ERR_NOMEM;
    std::vector<Rec> all(N);
    bool bad = E->fail(cudaMemcpyAsync(d_recs + N, &mine, sizeof mine, cudaMemcpyHostToDevice, E->stream), "ipc H2D") ||
               nccl_fail(E, g_nccl.AllGather(d_recs + N, d_recs, sizeof(Rec), ncclUint8, D->comm, E->stream), "ncclAllGather(handles)") ||
               E->fail(cudaMemcpyAsync(all.data(), d_recs, sizeof(Rec) * N, cudaMemcpyDeviceToHost, E->stream), "ipc D2H") ||
               E->fail(cudaStreamSynchronize(E->stream), "ipc exchange");
    dev_free(d_recs);
    if (bad) return AMX_ERR_CUDA;
    bool ok = true;
    unsigned long long seq = 0;
    for (uint32_t r = 0; r < N; ++r) { ok = ok && all[r].ok; seq = std::max<unsigned long long>(seq, all[r].seq); }
    for (uint32_t r = 0; r < N && ok; ++r) {
        if (r == D->rank) { D->peer_table[r] = E->table; D->peer_flags[r] = D->flags; D->peer_stage[r] = D->stage; continue; }
        void *pt = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pt, all[r].table, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pf, all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; if (pt) cudaIpcCloseMemHandle(pt); break; }
        D->peer_table[r] = (pword *) pt; D->peer_flags[r] = (unsigned long long *) pf;
        if (D->stage) {
            void *ps = nullptr;
            if (all[r].stage_cap != D->stage_cap || cudaIpcOpenMemHandle(&ps, all[r].stage, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
            D->peer_stage[r] = (pword *) ps;
        }
    }
    cudaGetLastError();
    // everybody must agree: one rank that cannot map means the NCCL path for all
    int *d_ok = nullptr;
    if (!dev_alloc(E, (void **) &d_ok, 8, "p2p vote")) return AMX_ERR_NOMEM;
    int h_ok = ok ? 1 : 0;
    bad = E->fail(cudaMemcpyAsync(d_ok, &h_ok, 4, cudaMemcpyHostToDevice, E->stream), "vote H2D") ||
          nccl_fail(E, g_nccl.AllReduce(d_ok, d_ok + 1, 1, ncclInt32, ncclMin, D->comm, E->stream), "ncclAllReduce(vote)") ||
          E->fail(cudaMemcpyAsync(&h_ok, d_ok + 1, 4, cudaMemcpyDeviceToHost, E->stream), "vote D2H") ||
          E->fail(cudaStreamSynchronize(E->stream), "p2p vote");
    dev_free(d_ok);
    if (bad) return AMX_ERR_CUDA;
    if (!h_ok) { drop_p2p(E); E->err = "peer mapping (cudaIpc) is not available on this box: the NCCL exchange stays in use"; return AMX_ERR_STATE; }
    cudaMemcpy(D->d_peer_flags, D->peer_flags, sizeof D->peer_flags, cudaMemcpyHostToDevice);
    D->barrier_seq = seq;                          // all ranks continue from the same sequence number
    D->p2p = true; D->p2p_table = E->table;
    return peer_barrier(E) == AMX_OK && !E->fail(cudaStreamSynchronize(E->stream), "first peer barrier") ? AMX_OK : AMX_ERR_CUDA;
}

int amx_comm_disable_p2p(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    if (ctx->e.dist) { cudaStreamSynchronize(ctx->e.stream); drop_p2p(&ctx->e); }
    return AMX_OK;
}

// Collective: rank `root`'s table (every column) replaces everybody's.  Same geometry on all ranks.
int amx_table_broadcast(amx_ctx *ctx, uint32_t root) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    Dist *D = E->dist;
    cudaSetDevice(E->device);
    if (!D || D->nranks == 1) return AMX_OK;
    if (!D->comm || !E->table || root >= D->nranks) return AMX_ERR_STATE;
    if (nccl_fail(E, g_nccl.Broadcast(E->table, E->table, (size_t) E->h * E->A, ncclUint64, (int) root, D->comm, E->stream), "ncclBroadcast(table)")) return AMX_ERR_CUDA;
    E->render_ready = false;
    return E->fail(cudaStreamSynchronize(E->stream), "table broadcast") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_swap_part_step(amx_ctx *ctx, uint32_t chain, uint32_t column, uint64_t step, uint32_t sub_epochs, uint32_t rounds) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_swap_part_step(&ctx->e, chain, column, step, sub_epochs, rounds);
}

int amx_swap_columns_step(amx_ctx *ctx, int32_t chain, uint32_t phase, uint64_t step, uint32_t epochs, uint32_t rounds) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_swap_columns_step(&ctx->e, chain, phase, step, epochs, rounds);
}

uint32_t amx_swap_phase_count(amx_ctx *ctx) { return ctx ? dist_phase_count(ctx->e.h) : 0u; }

int amx_column_hash(amx_ctx *ctx, uint32_t column, uint64_t out2[2]) {
    if (!ctx || !out2 || column >= ctx->e.h || !ctx->e.table) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    unsigned long long *d = nullptr;
    if (!dev_alloc(E, (void **) &d, 16, "hash")) return AMX_ERR_NOMEM;
    cudaMemsetAsync(d, 0, 16, E->stream);
    k_column_hash<<<std::min<uint32_t>(div_up(E->A, 256), (uint32_t) E->sm_count * 8u), 256, 0, E->stream>>>(E->table + (size_t) column * E->A, E->A, d);
    E->launches++;
    unsigned long long h[2] = {0, 0};
    const bool bad = E->fail(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, E->stream), "hash D2H") || E->fail(cudaStreamSynchronize(E->stream), "hash");
    dev_free(d);
    out2[0] = h[0]; out2[1] = h[1];
    return bad ? AMX_ERR_CUDA : AMX_OK;
}

// 1 when a peer barrier of this context ever timed out (a peer process died): the tables may differ from then on
int amx_comm_check(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    Dist *D = E->dist;
    if (!D || !D->d_timeout) return AMX_OK;
    cudaSetDevice(E->device);
    uint32_t t = 0;
    if (E->fail(cudaMemcpyAsync(&t, D->d_timeout, 4, cudaMemcpyDeviceToHost, E->stream), "timeout D2H") || E->fail(cudaStreamSynchronize(E->stream), "comm check")) return AMX_ERR_CUDA;
    if (t) { E->err = "a peer never arrived at a barrier"; return AMX_ERR_STATE; }
    return AMX_OK;
}

}
