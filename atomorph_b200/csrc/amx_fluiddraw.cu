/*
 * amx_fluiddraw.cu -- the fluid render path (SURVEY.md row a-F, driver part): morph::step_fluid (morph.cpp:807-1055),
 * update_particle (680-805) and draw_fluid (1057-1300) on the device.
 *
 * The reference keeps one particle slot per atom (density is forced to 1 with fluid, so slots and atoms of a chain
 * coincide), walks std::maps of "sources" (key points of column y that carry HAS_FLUID) and "destinations" (column
 * y+1), creates / deletes particles one at a time with draws from its host RNG until round((1-t) |blob_before| +
 * t |blob_after|) of them are active, steers every active particle towards its atom's interpolated position, runs
 * `fluid` MPM steps and splats the particles.  Here every stage is a kernel over particles, atoms or pixels:
 *
 *   index      per atom: canvas position -> atom for the sources (column y) and destinations (column y+1)
 *   retire     per particle: stay active only inside the same key-frame interval (morph.cpp:868-878); occupancy of
 *              sources / destinations, per-chain lists of active and free slots
 *   candidates per atom: unoccupied sources / destinations of its chain
 *   create     per candidate: a counter-RNG Feistel permutation of the candidate list picks `to_create` unoccupied
 *              sources uniformly without replacement (the reference draws them one by one: morph.cpp:897-925); when
 *              the sources run out the remaining particles take unoccupied destinations and start as a copy of a
 *              particle that shares their source (926-1000)
 *   delete     per active particle: the same permutation trick removes `to_delete` at random (1021-1034)
 *   update     per active particle: update_particle -- attractor = the atom's position at t (linear / Catmull-Rom),
 *              ideal colour = faded end colours (linear / cosine / Perlin)
 *   step       `fluid` x FluidModel::step (amx_fluid.cu)
 *   splat      per active particle: colour pull, bilinear splat with double weights into per-pixel sums
 *   resolve    per pixel: round(sum(c w) / sum(w)), background fade-out factor 1 - t^24
 *   feather    `feather` x one layer GROWN outwards: a border pixel takes the mean of its blob neighbours (1204-1266)
 *   final      per pixel: blend over the background (1268-1299)
 *
 * Which slot a new particle lands in, which of the unoccupied sources it takes and the order of the floating-point
 * sums are the reference's only through its host RNG / iteration order; they are statistically equivalent here, and a
 * frame in which every source gets exactly one particle (t = 0 of an interval) is the same image up to summation order.
 */
#include <algorithm>
#include <cmath>
#include <random>
#include "amx_fluid.h"

namespace amx {

#define FNIL 0xffffffffu
#define PFI(k) pf[(size_t) (k) * n + i]
enum { FC_ACTIVE, FC_NACT, FC_NFREE, FC_NSRC, FC_NDST, FC_STRIDE = 8 };

struct FluidDraw {
    uint32_t *pkey = nullptr, *psrc = nullptr, *pdst = nullptr;          // [n] frame key, source / destination canvas position
    uint32_t *src_atom = nullptr, *dst_atom = nullptr;                    // [canvas]
    uint32_t *src_occ = nullptr, *dst_occ = nullptr, *src_rep = nullptr;  // [canvas]
    uint32_t *act_list = nullptr, *free_list = nullptr, *cand_src = nullptr, *cand_dst = nullptr;   // [n], per-chain segments
    uint32_t *cnt = nullptr;                                              // [nchains][FC_STRIDE]
    int32_t  *limit = nullptr;                                            // [nchains]
    double   *surf = nullptr;                                             // [nchains][2] |blob_before|, |blob_after|
    uint32_t *avgcol = nullptr;                                           // unused slot (AVERAGE colours come as doubles below)
    double   *avgd = nullptr;                                             // [nchains][4] blob_before r g b a (show_blobs AVERAGE)
    double   *distinct = nullptr;                                         // [nchains][3] DISTINCT colours (host mt19937)
    double   *acc = nullptr;                                              // [5][np] sum(r w) sum(g w) sum(b w) sum(a w) sum(w)
    uint32_t *img[2] = {nullptr, nullptr};                                // feather ping-pong: colour
    uint8_t  *has[2] = {nullptr, nullptr};                                //                    presence
    size_t canvas = 0, np = 0;
    uint32_t nchains = 0;
    uint64_t frame_serial = 0;
};

void fluid_draw_free(Fluid *F) {
    if (!F || !F->draw) return;
    FluidDraw *D = F->draw;
    dev_free(D->pkey); dev_free(D->psrc); dev_free(D->pdst); dev_free(D->src_atom); dev_free(D->dst_atom);
    dev_free(D->src_occ); dev_free(D->dst_occ); dev_free(D->src_rep); dev_free(D->act_list); dev_free(D->free_list);
    dev_free(D->cand_src); dev_free(D->cand_dst); dev_free(D->cnt); dev_free(D->limit); dev_free(D->surf);
    dev_free(D->avgd); dev_free(D->distinct); dev_free(D->acc);
    dev_free(D->img[0]); dev_free(D->img[1]); dev_free(D->has[0]); dev_free(D->has[1]);
    delete D;
    F->draw = nullptr;
}

struct FParams {
    uint32_t A, n, nchains, h, y, yn, frame_key;
    uint32_t cw, ch, width, height, bx1, by1, bx2, by2;
    uint32_t motion, fading, show_blobs, keep_background, feather;
    double t, time, w;                 // local t, global time, 1 - t
    double b1, b2, b3, b4;             // Catmull-Rom basis
    double str_cos;                    // cosine-eased colour weight (host libm, like the atom renderer)
    int p0, p1, p2, p3;                // Catmull-Rom control columns
    uint64_t seed, serial;
    uint32_t gsx, gsy;
};

// random bijection of [0, n) by cycle-walking a 4-round Feistel network (same construction as amx_chain.cu)
__device__ __forceinline__ uint32_t fperm(uint32_t i, uint32_t n, uint64_t seed) {
    unsigned bits = 32 - __clz(n | 1u);
    unsigned hb = (bits + 1) / 2;
    if (hb == 0) hb = 1;
    uint32_t mask = (1u << hb) - 1u;
    uint32_t x = i;
    do {
        uint32_t L = x >> hb, R = x & mask;
        for (int r = 0; r < 4; ++r) {
            uint32_t Fv = (uint32_t) mix64((uint64_t) R ^ (seed + 0x9e37u * (uint64_t) r)) & mask;
            uint32_t nl = R;
            R = L ^ Fv;
            L = nl;
        }
        x = (L << hb) | R;
    } while (x >= n);
    return x;
}

__device__ __forceinline__ uint32_t canvas_pos(pword p, uint32_t cw, uint32_t ch) {
    uint32_t x = pw_x(p), y = pw_y(p);
    return (x < cw && y < ch) ? y * cw + x : FNIL;
}

// ---- index: sources of column y, destinations of column y+1 (morph.cpp:846-865)
__global__ void __launch_bounds__(256)
k_fd_index(const pword *__restrict__ table, FParams P, uint32_t *__restrict__ src_atom, uint32_t *__restrict__ dst_atom) {
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.A) return;
    pword pt1 = table[(size_t) P.y * P.A + a], pt2 = table[(size_t) P.yn * P.A + a];
    if (pw_flags(pt1) & F_HAS_FLUID) { uint32_t c = canvas_pos(pt1, P.cw, P.ch); if (c != FNIL) src_atom[c] = a; }
    if (pw_flags(pt2) & F_HAS_FLUID) { uint32_t c = canvas_pos(pt2, P.cw, P.ch); if (c != FNIL) dst_atom[c] = a; }
}

// Append to a per-chain list: the lanes of a warp that append to the same list (same counter) claim their slots with ONE
// atomicAdd (a chain-wide counter bumped by every particle separately serialises at one L2 address: 1.7 ms per frame at
// 1 M particles).  Must be called by all 32 lanes; returns the slot of a lane with `pred`, undefined otherwise.
__device__ __forceinline__ uint32_t warp_claim(uint32_t *counter_base, uint32_t which, bool pred) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(0xffffffffu, pred ? which : 0xffffffffu);
    const uint32_t leader = (uint32_t) __ffs((int) peers) - 1u;
    uint32_t base = 0u;
    if (pred && lane == leader) base = atomicAdd(counter_base + which, (uint32_t) __popc(peers));
    base = __shfl_sync(0xffffffffu, base, (int) leader);
    return base + (uint32_t) __popc(peers & ((1u << lane) - 1u));
}

// ---- retire + occupancy + lists
__global__ void __launch_bounds__(256)
k_fd_retire(FParams P, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off, uint8_t *__restrict__ active,
            const uint32_t *__restrict__ pkey, const uint32_t *__restrict__ psrc, const uint32_t *__restrict__ pdst,
            uint32_t *__restrict__ src_occ, uint32_t *__restrict__ dst_occ, uint32_t *__restrict__ src_rep,
            uint32_t *__restrict__ act_list, uint32_t *__restrict__ free_list, uint32_t *__restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < P.n;                                  // (the grid is a whole number of warps: every lane takes part in the claims)
    uint32_t c = valid ? chain_of[i] : 0u;
    uint32_t off = (uint32_t) chain_off[c];
    bool keep = valid && active[i] && pkey[i] == P.frame_key;
    const uint32_t ka = warp_claim(cnt, c * FC_STRIDE + FC_NACT, keep);
    const uint32_t kf = warp_claim(cnt, c * FC_STRIDE + FC_NFREE, valid && !keep);
    warp_claim(cnt, c * FC_STRIDE + FC_ACTIVE, keep);
    if (!valid) return;
    if (keep) {
        if (psrc[i] != FNIL) { atomicAdd(&src_occ[psrc[i]], 1u); src_rep[psrc[i]] = i; }
        if (pdst[i] != FNIL) atomicAdd(&dst_occ[pdst[i]], 1u);
        act_list[off + ka] = i;
    } else {
        active[i] = 0;
        free_list[off + kf] = i;
    }
}

__global__ void __launch_bounds__(256)
k_fd_candidates(const pword *__restrict__ table, FParams P, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off,
                const uint32_t *__restrict__ src_occ, const uint32_t *__restrict__ dst_occ, const uint32_t *__restrict__ src_atom,
                const uint32_t *__restrict__ dst_atom, uint32_t *__restrict__ cand_src, uint32_t *__restrict__ cand_dst, uint32_t *__restrict__ cnt) {
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = a < P.A;
    uint32_t c = valid ? chain_of[a] : 0u;
    uint32_t off = (uint32_t) chain_off[c];
    bool is_src = false, is_dst = false;
    if (valid) {
        pword pt1 = table[(size_t) P.y * P.A + a], pt2 = table[(size_t) P.yn * P.A + a];
        uint32_t c1 = canvas_pos(pt1, P.cw, P.ch), c2 = canvas_pos(pt2, P.cw, P.ch);
        is_src = (pw_flags(pt1) & F_HAS_FLUID) && c1 != FNIL && src_atom[c1] == a && src_occ[c1] == 0;
        is_dst = (pw_flags(pt2) & F_HAS_FLUID) && c2 != FNIL && dst_atom[c2] == a && dst_occ[c2] == 0;
    }
    const uint32_t ks = warp_claim(cnt, c * FC_STRIDE + FC_NSRC, is_src);
    const uint32_t kd = warp_claim(cnt, c * FC_STRIDE + FC_NDST, is_dst);
    if (is_src) cand_src[off + ks] = a;
    if (is_dst) cand_dst[off + kd] = a;
}

// ---- create (morph.cpp:881-1018): thread j of a chain's segment handles candidate j
__global__ void __launch_bounds__(256)
k_fd_create(const pword *__restrict__ table, FParams P, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off,
            const uint32_t *__restrict__ cnt, const int32_t *__restrict__ limit, const uint32_t *__restrict__ free_list,
            const uint32_t *__restrict__ cand_src, const uint32_t *__restrict__ cand_dst, const uint32_t *__restrict__ src_rep,
            double *__restrict__ pf, uint8_t *__restrict__ active, uint8_t *__restrict__ mature, uint8_t *__restrict__ owner,
            uint32_t *__restrict__ pkey, uint32_t *__restrict__ psrc, uint32_t *__restrict__ pdst) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n) return;
    const uint32_t n = P.n;
    uint32_t c = chain_of[g];
    uint32_t off = (uint32_t) chain_off[c], j = g - off;
    const uint32_t *k = cnt + c * FC_STRIDE;
    int64_t want = (int64_t) limit[c] - (int64_t) k[FC_ACTIVE];
    if (want <= 0) return;
    uint32_t to_create = want < (int64_t) k[FC_NFREE] ? (uint32_t) want : k[FC_NFREE];
    uint32_t nsrc = k[FC_NSRC], ndst = k[FC_NDST];
    uint64_t seed = rng64(P.seed, 0xf1d0u + c, P.serial);
    if (j < nsrc) {
        // an unoccupied source: taken when its rank in a random order is below to_create
        uint32_t r = fperm(j, nsrc, seed);
        if (r < to_create) {
            uint32_t a = cand_src[off + j], i = free_list[off + r];
            pword pt1 = table[(size_t) P.y * P.A + a], pt2 = table[(size_t) P.yn * P.A + a];
            pkey[i] = P.frame_key;
            psrc[i] = canvas_pos(pt1, P.cw, P.ch);
            pdst[i] = canvas_pos(pt2, P.cw, P.ch);       // its default destination (morph.cpp:913-923)
            PFI(PF_U) = 0.0; PFI(PF_V) = 0.0; PFI(PF_STRENGTH) = 1.0;     // Particle::clear()
            mature[i] = 0; owner[i] = 1; active[i] = 1;
        }
    }
    if (to_create > nsrc && j < ndst) {
        // sources exhausted: unoccupied destinations, the new particle starts as a copy of one that shares its source
        uint32_t extra = min(to_create - nsrc, ndst);
        if (j < extra) {
            uint32_t a = cand_dst[off + j], i = free_list[off + nsrc + j];
            pword pt1 = table[(size_t) P.y * P.A + a], pt2 = table[(size_t) P.yn * P.A + a];
            uint32_t sp = canvas_pos(pt1, P.cw, P.ch);
            pkey[i] = P.frame_key;
            psrc[i] = sp;
            pdst[i] = canvas_pos(pt2, P.cw, P.ch);
            PFI(PF_U) = 0.0; PFI(PF_V) = 0.0; PFI(PF_STRENGTH) = 1.0;
            uint32_t parent = sp != FNIL ? src_rep[sp] : FNIL;
            if (parent != FNIL && parent < n) {
                // copy_from(parent) + fuzz (morph.cpp:975-990)
                const int copy[] = {PF_X, PF_Y, PF_GX, PF_GY, PF_FREE, PF_RI, PF_GI, PF_BI, PF_AI, PF_R, PF_G, PF_B, PF_A, PF_STRENGTH};
                for (int q = 0; q < 14; ++q) pf[(size_t) copy[q] * n + i] = pf[(size_t) copy[q] * n + parent];
                uint64_t rr = rng64(P.seed, 0xf1d1u + c, P.serial * 0x100000000ull + i);
                double f1 = (double) (rr & 0xffffffffu) / 4294967296.0, f2 = (double) (rr >> 32) / 4294967296.0;
                PFI(PF_X) = fmin(PFI(PF_X) + f1, (double) P.width - 10.0);
                PFI(PF_Y) = fmin(PFI(PF_Y) + f2, (double) P.height - 10.0);
                mature[i] = mature[parent]; owner[i] = 0;
            } else {
                mature[i] = 0; owner[i] = 1;                 // nobody to copy from: start at the attractor like a source owner
            }
            active[i] = 1;
        }
    }
}

// ---- delete (morph.cpp:1021-1034)
__global__ void __launch_bounds__(256)
k_fd_delete(FParams P, const uint32_t *__restrict__ chain_of, const uint64_t *__restrict__ chain_off, const uint32_t *__restrict__ cnt,
            const int32_t *__restrict__ limit, const uint32_t *__restrict__ act_list, uint8_t *__restrict__ active) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n) return;
    uint32_t c = chain_of[g];
    uint32_t off = (uint32_t) chain_off[c], j = g - off;
    const uint32_t *k = cnt + c * FC_STRIDE;
    int64_t over = (int64_t) k[FC_ACTIVE] - (int64_t) limit[c];
    if (over <= 0 || j >= k[FC_NACT]) return;
    if (fperm(j, k[FC_NACT], rng64(P.seed, 0xf1d2u + c, P.serial)) < (uint64_t) over) active[act_list[off + j]] = 0;
}

struct FDevCos { __device__ double operator()(double x) const { return cos(x); } };

// ---- update_particle (morph.cpp:680-805)
__global__ void __launch_bounds__(128)
k_fd_update(const pword *__restrict__ table, FParams P, const uint32_t *__restrict__ chain_of, const uint32_t *__restrict__ fetch_y,
            const uint32_t *__restrict__ fetch_yn, const int32_t *__restrict__ perlin, const uint32_t *__restrict__ src_atom,
            const uint32_t *__restrict__ dst_atom, const double *__restrict__ surf, const double *__restrict__ avgd,
            const double *__restrict__ distinct, double *__restrict__ pf, const uint8_t *__restrict__ active, const uint8_t *__restrict__ mature,
            const uint8_t *__restrict__ owner, const uint32_t *__restrict__ psrc, const uint32_t *__restrict__ pdst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n || !active[i]) return;
    const uint32_t n = P.n;
    uint32_t c = chain_of[i];
    double previous_surface = surf[2 * c], next_surface = surf[2 * c + 1];
    uint32_t sp = psrc[i], dp = pdst[i];
    uint32_t src_x = sp != FNIL ? src_atom[sp] : FNIL, dst_x = dp != FNIL ? dst_atom[dp] : FNIL;
    if (next_surface > previous_surface) src_x = dst_x;        // growing area: the real source comes from the destination
    if (src_x == FNIL) src_x = dst_x;
    if (dst_x == FNIL) dst_x = src_x;
    if (src_x == FNIL) return;                                 // a std::map default (atom 0) in the reference; not reachable with valid tables
    pword pt1 = table[(size_t) P.y * P.A + src_x], pt2 = table[(size_t) P.yn * P.A + dst_x];
    bool has1 = pw_flags(pt1) & F_HAS_PIXEL, has2 = pw_flags(pt2) & F_HAS_PIXEL;
    auto fetch = [&](const uint32_t *img, pword p) -> uint32_t {
        uint32_t x = pw_x(p), yy = pw_y(p);
        return (x < P.cw && yy < P.ch) ? img[(size_t) yy * P.cw + x] : 0u;
    };
    uint32_t c1, c2;
    if (has1 && !has2) { c1 = fetch(fetch_y, pt1); c2 = c1; if (next_surface > 0.0) c2 &= 0x00ffffffu; }
    else if (has2 && !has1) { c2 = fetch(fetch_yn, pt2); c1 = c2; if (previous_surface > 0.0) c1 &= 0x00ffffffu; }
    else { c1 = fetch(fetch_y, pt1); c2 = fetch(fetch_yn, pt2); }
    double str = P.w;
    if (P.fading == K_PERLIN) {
        double f = 8.0;
        double bbox_w = (double) ((int) P.bx2 - (int) P.bx1) + 1.0, bbox_h = (double) ((int) P.by2 - (int) P.by1) + 1.0;
        double px = ((double) (((int) pw_x(pt1) - (int) P.bx1) * 256 + (int) pw_xf(pt1)) / (bbox_w * 256.0)) * f;
        double py = ((double) (((int) pw_y(pt1) - (int) P.by1) * 256 + (int) pw_yf(pt1)) / (bbox_h * 256.0)) * f;
        double lag = pn_octave2(perlin, px, py, 8) * 0.5 + 0.5, slope = pn_octave2(perlin + 512, px, py, 8) * 0.5 + 0.5;
        str = ease_strength(lag, slope, P.w, FDevCos());
    } else if (P.fading == K_COSINE) str = P.str_cos;
    uint32_t col = lerp_color(c1, c2, str);
    // attractor position
    uint32_t x = pw_x(pt1), y = pw_y(pt1), xf = pw_xf(pt1), yf = pw_yf(pt1);
    if (P.motion == K_LINEAR) lerp_point(pt1, pt2, P.w, &x, &y, &xf, &yf);
    else if (P.motion == K_SPLINE) {
        // the spline of atom src_x (morph.cpp:742-748)
        pword q0 = table[(size_t) P.p0 * P.A + src_x], q1 = table[(size_t) P.p1 * P.A + src_x];
        pword q2 = table[(size_t) P.p2 * P.A + src_x], q3 = table[(size_t) P.p3 * P.A + src_x];
        double vx = cr_eval(pw_xd(q0), pw_xd(q1), pw_xd(q2), pw_xd(q3), P.b1, P.b2, P.b3, P.b4);
        double vy = cr_eval(pw_yd(q0), pw_yd(q1), pw_yd(q2), pw_yd(q3), P.b1, P.b2, P.b3, P.b4);
        split_spline_coord(vx, &x, &xf);
        split_spline_coord(vy, &y, &yf);
    }
    float ptx = (float) x + (float) xf / 256.0f, pty = (float) y + (float) yf / 256.0f;      // point2xy (atomorph.h:315-318)
    double gx = fmin(fmax(((double) ptx * 1.0) + 10.0, 1.0), (double) P.width + 10.0);
    double gy = fmin(fmax(((double) pty * 1.0) + 10.0, 1.0), (double) P.height + 10.0);
    PFI(PF_GX) = gx; PFI(PF_GY) = gy;
    double R = c_r(col) / 255.0, G = c_g(col) / 255.0, B = c_b(col) / 255.0, Av = c_a(col) / 255.0;
    if (P.show_blobs == SHOW_DISTINCT) { R = distinct[3 * c]; G = distinct[3 * c + 1]; B = distinct[3 * c + 2]; Av = 1.0; }
    else if (P.show_blobs == SHOW_AVERAGE) { R = avgd[4 * c]; G = avgd[4 * c + 1]; B = avgd[4 * c + 2]; Av = avgd[4 * c + 3]; }
    PFI(PF_RI) = R; PFI(PF_GI) = G; PFI(PF_BI) = B; PFI(PF_AI) = Av;
    double t = P.t;
    double r = t * R + (1.0 - t) * PFI(PF_R), g = t * G + (1.0 - t) * PFI(PF_G);
    double b = t * B + (1.0 - t) * PFI(PF_B), a = t * Av + (1.0 - t) * PFI(PF_A);
    if (!mature[i]) {
        if (owner[i]) { PFI(PF_X) = gx; PFI(PF_Y) = gy; r = R; g = G; b = B; a = Av; }
        PFI(PF_STRENGTH) = previous_surface == 0.0 ? 0.1 : 1.0;
        PFI(PF_FREE) = 1.0;
    }
    PFI(PF_R) = r; PFI(PF_G) = g; PFI(PF_B) = b; PFI(PF_A) = a;
}

// ---- splat (morph.cpp:1082-1160)
__global__ void __launch_bounds__(128)
k_fd_splat(FParams P, double *__restrict__ pf, const uint8_t *__restrict__ active, uint8_t *__restrict__ mature, double *__restrict__ acc, size_t np) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n || !active[i]) return;
    const uint32_t n = P.n;
    double t = P.t, strength = PFI(PF_STRENGTH);
    {
        double d = t - (1.0 * (1.0 - strength));
        double w = (1.0 / (100.0 * (d * d) + 1.0)) * (strength * strength);
        PFI(PF_R) = (w * PFI(PF_RI)) + (1.0 - w) * PFI(PF_R);
        PFI(PF_G) = (w * PFI(PF_GI)) + (1.0 - w) * PFI(PF_G);
        PFI(PF_B) = (w * PFI(PF_BI)) + (1.0 - w) * PFI(PF_B);
        PFI(PF_A) = (w * PFI(PF_AI)) + (1.0 - w) * PFI(PF_A);
    }
    double X = PFI(PF_X), Y = PFI(PF_Y), px, py;
    if (X <= 10.0) px = 0.0; else if (X >= (double) P.width + 10.0) px = (double) (P.width - 1); else px = (X - 10.0);
    if (Y <= 10.0) py = 0.0; else if (Y >= (double) P.height + 10.0) py = (double) (P.height - 1); else py = (Y - 10.0);
    uint32_t x = (uint32_t) min((int) px, (int) P.width - 1), y = (uint32_t) min((int) py, (int) P.height - 1);
    mature[i] = 1;
    double ipx, ipy;
    double x_fract = modf(px, &ipx), y_fract = modf(py, &ipy);
    if (ipx >= (double) P.width || ipy >= (double) P.height) {
        if (ipx > (double) P.bx2 || ipx < (double) P.bx1 || ipy > (double) P.by2 || ipy < (double) P.by1) return;
    }
    double w11 = (1.0 - x_fract) * (1.0 - y_fract), w21 = x_fract * (1.0 - y_fract), w12 = (1.0 - x_fract) * y_fract, w22 = x_fract * y_fract;
    uint32_t col = create_color_d(PFI(PF_R), PFI(PF_G), PFI(PF_B), PFI(PF_A));
    double cr = (double) c_r(col), cg = (double) c_g(col), cb = (double) c_b(col), ca = (double) c_a(col);
    auto put = [&](uint32_t xx, uint32_t yy, double w) {
        if (xx >= P.width || yy >= P.height) return;      // the image holds width x height pixels (morph.cpp:1187: imgpos = y*width + x)
        size_t p = (size_t) yy * P.width + xx;
        atomicAdd(&acc[p], cr * w); atomicAdd(&acc[np + p], cg * w); atomicAdd(&acc[2 * np + p], cb * w);
        atomicAdd(&acc[3 * np + p], ca * w); atomicAdd(&acc[4 * np + p], w);
    };
    if (w11 > 0.0) put(x, y, w11);
    if ((x < P.bx2 || x + 1 < P.width) && w21 > 0.0) put(x + 1, y, w21);
    if ((y < P.by2 || y + 1 < P.height) && w12 > 0.0) put(x, y + 1, w12);
    if (w22 > 0.0 && ((y < P.by2 && x < P.bx2) || (y + 1 < P.height && x + 1 < P.width))) put(x + 1, y + 1, w22);
}

// ---- resolve (morph.cpp:1162-1202): pixel colour or "absent"
__global__ void __launch_bounds__(256)
k_fd_resolve(FParams P, const double *__restrict__ acc, size_t np, uint32_t *__restrict__ img, uint8_t *__restrict__ has) {
    size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double ws = acc[4 * np + p];
    if (!(ws > 0.0)) { has[p] = 0; img[p] = 0; return; }
    double r = round(acc[p] / ws), g = round(acc[np + p] / ws), b = round(acc[2 * np + p] / ws), a = round(acc[3 * np + p] / ws);
    if (P.keep_background) a *= 1.0 - pow(P.t, 24.0);
    img[p] = c_make(to_u8(r), to_u8(g), to_u8(b), to_u8(a));      // create_pixel(x, y, r, g, b, a): doubles truncate to unsigned char
    has[p] = 1;
}

// ---- one feather layer grown outwards (morph.cpp:1204-1256)
__global__ void __launch_bounds__(256)
k_fd_feather(FParams P, double alpha, const uint32_t *__restrict__ img_in, const uint8_t *__restrict__ has_in, uint32_t *__restrict__ img_out,
             uint8_t *__restrict__ has_out, size_t np) {
    size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    if (has_in[p]) { img_out[p] = img_in[p]; has_out[p] = 1; return; }
    uint32_t x = (uint32_t) (p % P.width), y = (uint32_t) (p / P.width);
    // neighbours in the reference's std::set order of xy2pos: (x, y-1), (x-1, y), (x+1, y), (x, y+1)
    size_t nb[4]; int k = 0;
    if (y > 0 && has_in[p - P.width]) nb[k++] = p - P.width;
    if (x > 0 && has_in[p - 1]) nb[k++] = p - 1;
    if (x + 1 < P.width && has_in[p + 1]) nb[k++] = p + 1;
    if (y + 1 < P.height && has_in[p + P.width]) nb[k++] = p + P.width;
    if (k == 0) { img_out[p] = 0; has_out[p] = 0; return; }
    double r = 0.0, g = 0.0, b = 0.0, a = 0.0, w = (double) k;
    for (int q = 0; q < k; ++q) {
        uint32_t c = img_in[nb[q]];
        r += (c_r(c) / 255.0) / w; g += (c_g(c) / 255.0) / w; b += (c_b(c) / 255.0) / w; a += (c_a(c) / 255.0) / w;
    }
    a *= alpha;
    img_out[p] = create_color_d(r, g, b, a);
    has_out[p] = 1;
}

// ---- final blend over the background (morph.cpp:1268-1299)
__global__ void __launch_bounds__(256)
k_fd_final(FParams P, const uint32_t *__restrict__ img, const uint8_t *__restrict__ has, const uint32_t *__restrict__ bg, uint32_t *__restrict__ out, size_t np) {
    size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    uint32_t bgc = P.keep_background ? bg[p] : 0u;
    if (!has[p]) { out[p] = bgc; return; }
    uint32_t c = img[p];
    if (P.keep_background) {
        double bgr = c_r(bgc) / 255.0, bgg = c_g(bgc) / 255.0, bgb = c_b(bgc) / 255.0, bga = c_a(bgc) / 255.0;
        double r = c_r(c) / 255.0, g = c_g(c) / 255.0, b = c_b(c) / 255.0, a = c_a(c) / 255.0;
        r = a * r + (1.0 - a) * bgr; g = a * g + (1.0 - a) * bgg; b = a * b + (1.0 - a) * bgb; a = bga + (1.0 - bga) * a;
        c = create_color_d(r, g, b, a);
    }
    out[p] = c;
}

__global__ void __launch_bounds__(256) k_fd_init_particles(double *__restrict__ pf, uint32_t n, double x0, double y0) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // morph.cpp:245-262
    PFI(PF_X) = x0; PFI(PF_Y) = y0; PFI(PF_U) = 0.0; PFI(PF_V) = 0.0; PFI(PF_GX) = x0; PFI(PF_GY) = y0; PFI(PF_FREE) = 1.0;
    PFI(PF_RI) = 0.0; PFI(PF_GI) = 1.0; PFI(PF_BI) = 0.0; PFI(PF_AI) = 1.0;
    PFI(PF_R) = 0.0; PFI(PF_G) = 1.0; PFI(PF_B) = 0.0; PFI(PF_A) = 1.0; PFI(PF_STRENGTH) = 1.0;
}

// ------------------------------------------------------------------------------------------ host side
static int fluid_draw_ensure(Engine *E) {
    const size_t np = (size_t) E->width * E->height, cv = E->canvas();
    Fluid *F = E->fluid;
    if (F && (F->n != E->A || F->gx != E->width + 20 || F->gy != E->height + 20)) { engine_fluid_free(E); F = nullptr; }
    if (!F) {
        // FluidModel(width + 20, height + 20, sum of max_surface) (morph.cpp:235-243); density is 1 with fluid, so slots = atoms
        int rcode = fluid_alloc(E, E->width + 20, E->height + 20, (uint32_t) E->A);
        if (rcode != AMX_OK) return rcode;
        F = E->fluid;
        k_fd_init_particles<<<div_up(F->n, 256), 256, 0, E->stream>>>(F->pf, F->n, E->width / 2.0 - 10.0, E->height / 2.0 - 10.0);
        E->launches++;
    }
    if (F->draw && (F->draw->canvas != cv || F->draw->np != np || F->draw->nchains != E->nchains)) fluid_draw_free(F);
    if (!F->draw) {
        FluidDraw *D = new FluidDraw();
        F->draw = D;
        D->canvas = cv; D->np = np; D->nchains = E->nchains;
        const size_t n = F->n;
        bool ok = dev_alloc(E, (void **) &D->pkey, n * 4, "fd") && dev_alloc(E, (void **) &D->psrc, n * 4, "fd") && dev_alloc(E, (void **) &D->pdst, n * 4, "fd") &&
                  dev_alloc(E, (void **) &D->src_atom, cv * 4, "fd") && dev_alloc(E, (void **) &D->dst_atom, cv * 4, "fd") &&
                  dev_alloc(E, (void **) &D->src_occ, cv * 4, "fd") && dev_alloc(E, (void **) &D->dst_occ, cv * 4, "fd") && dev_alloc(E, (void **) &D->src_rep, cv * 4, "fd") &&
                  dev_alloc(E, (void **) &D->act_list, n * 4, "fd") && dev_alloc(E, (void **) &D->free_list, n * 4, "fd") &&
                  dev_alloc(E, (void **) &D->cand_src, n * 4, "fd") && dev_alloc(E, (void **) &D->cand_dst, n * 4, "fd") &&
                  dev_alloc(E, (void **) &D->cnt, (size_t) E->nchains * FC_STRIDE * 4, "fd") && dev_alloc(E, (void **) &D->limit, (size_t) E->nchains * 4, "fd") &&
                  dev_alloc(E, (void **) &D->surf, (size_t) E->nchains * 16, "fd") && dev_alloc(E, (void **) &D->avgd, (size_t) E->nchains * 32, "fd") &&
                  dev_alloc(E, (void **) &D->distinct, (size_t) E->nchains * 24, "fd") && dev_alloc(E, (void **) &D->acc, np * 5 * 8, "fd") &&
                  dev_alloc(E, (void **) &D->img[0], np * 4, "fd") && dev_alloc(E, (void **) &D->img[1], np * 4, "fd") &&
                  dev_alloc(E, (void **) &D->has[0], np, "fd") && dev_alloc(E, (void **) &D->has[1], np, "fd");
        if (!ok) { fluid_draw_free(F); return AMX_ERR_NOMEM; }
        cudaMemsetAsync(D->pkey, 0xff, n * 4, E->stream);
        cudaMemsetAsync(D->psrc, 0xff, n * 4, E->stream);
        cudaMemsetAsync(D->pdst, 0xff, n * 4, E->stream);
    }
    return AMX_OK;
}

struct FLibmCos { double operator()(double x) const { return ::cos(x); } };

// one frame of the fluid path into d_dst (width*height packed RGBA); d_bg = background image or nullptr
int engine_render_fluid(Engine *E, double time, uint32_t f, double tl, const uint32_t *d_bg, uint32_t *d_dst) {
    if (E->p.density != 1) { E->err = "fluid needs density 1 (morph.cpp:1553-1559)"; return AMX_ERR_STATE; }
    int rcode = fluid_draw_ensure(E);
    if (rcode != AMX_OK) return rcode;
    Fluid *F = E->fluid;
    FluidDraw *D = F->draw;
    const size_t np = D->np, cv = D->canvas;
    const uint32_t n = F->n, nch = E->nchains, h = E->h;

    FParams P;
    P.A = (uint32_t) E->A; P.n = n; P.nchains = nch; P.h = h; P.y = f; P.yn = (f + 1) % h;
    P.frame_key = (uint32_t) E->frames[f].key;
    P.cw = E->cw; P.ch = E->ch; P.width = E->width; P.height = E->height;
    P.bx1 = E->bbox[0]; P.by1 = E->bbox[1]; P.bx2 = E->bbox[2]; P.by2 = E->bbox[3];
    P.motion = E->p.motion; P.fading = E->p.fading; P.show_blobs = E->p.show_blobs;
    P.keep_background = E->p.keep_background ? 1u : 0u;
    P.feather = (uint32_t) std::min<uint64_t>(E->p.feather, 4096);
    P.t = tl; P.time = time; P.w = 1.0 - tl;
    double lt;
    cr_locate(time, (int) h, &P.p0, &P.p1, &P.p2, &P.p3, &lt);
    cr_basis(lt, &P.b1, &P.b2, &P.b3, &P.b4);
    P.str_cos = ease_strength(0.5, 0.5, P.w, FLibmCos());
    P.seed = E->p.seed; P.serial = D->frame_serial++;
    P.gsx = F->gx; P.gsy = F->gy;

    // per chain: blob sizes before / after, active limit (morph.cpp:827-835), AVERAGE / DISTINCT colours
    std::vector<double> surf(2 * nch, 0.0), avgd(4 * nch, 0.0), distinct(3 * nch, 0.0);
    std::vector<int32_t> limit(nch, 0);
    const FrameDev &fa = E->frames[f], &fb = E->frames[(f + 1) % E->frames.size()];
    for (uint32_t c = 0; c < nch; ++c) {
        uint64_t g = E->chain_key[c];
        for (const BlobHost &b : fa.blobs) if (b.group == g) { surf[2 * c] = (double) b.size; for (int q = 0; q < 4; ++q) avgd[4 * c + q] = b.stats[2 + q]; break; }
        for (const BlobHost &b : fb.blobs) if (b.group == g) { surf[2 * c + 1] = (double) b.size; break; }
        limit[c] = (int32_t) std::round((1.0 - tl) * surf[2 * c] + tl * surf[2 * c + 1]);
        std::mt19937 gen((unsigned) g);                                   // morph.cpp:763-768
        std::uniform_real_distribution<double> dist(0.0, 1.0);
        distinct[3 * c] = dist(gen); distinct[3 * c + 1] = dist(gen); distinct[3 * c + 2] = dist(gen);
    }
    cudaMemcpyAsync(D->surf, surf.data(), surf.size() * 8, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(D->avgd, avgd.data(), avgd.size() * 8, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(D->distinct, distinct.data(), distinct.size() * 8, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(D->limit, limit.data(), limit.size() * 4, cudaMemcpyHostToDevice, E->stream);
    cudaStreamSynchronize(E->stream);           // the host vectors above go out of scope

    cudaMemsetAsync(D->src_atom, 0xff, cv * 4, E->stream);
    cudaMemsetAsync(D->dst_atom, 0xff, cv * 4, E->stream);
    cudaMemsetAsync(D->src_rep, 0xff, cv * 4, E->stream);
    cudaMemsetAsync(D->src_occ, 0, cv * 4, E->stream);
    cudaMemsetAsync(D->dst_occ, 0, cv * 4, E->stream);
    cudaMemsetAsync(D->cnt, 0, (size_t) nch * FC_STRIDE * 4, E->stream);
    const unsigned ga = div_up(P.A, 256), gn = div_up(n, 256);
    k_fd_index<<<ga, 256, 0, E->stream>>>(E->table, P, D->src_atom, D->dst_atom);
    k_fd_retire<<<gn, 256, 0, E->stream>>>(P, E->chain_of, E->d_chain_off, F->active, D->pkey, D->psrc, D->pdst, D->src_occ, D->dst_occ, D->src_rep,
                                          D->act_list, D->free_list, D->cnt);
    k_fd_candidates<<<ga, 256, 0, E->stream>>>(E->table, P, E->chain_of, E->d_chain_off, D->src_occ, D->dst_occ, D->src_atom, D->dst_atom, D->cand_src, D->cand_dst, D->cnt);
    k_fd_create<<<gn, 256, 0, E->stream>>>(E->table, P, E->chain_of, E->d_chain_off, D->cnt, D->limit, D->free_list, D->cand_src, D->cand_dst, D->src_rep,
                                          F->pf, F->active, F->mature, F->owner, D->pkey, D->psrc, D->pdst);
    k_fd_delete<<<gn, 256, 0, E->stream>>>(P, E->chain_of, E->d_chain_off, D->cnt, D->limit, D->act_list, F->active);
    k_fd_update<<<div_up(n, 128), 128, 0, E->stream>>>(E->table, P, E->chain_of, E->frames[f].fetch, E->frames[(f + 1) % E->frames.size()].fetch, E->d_perlin,
                                                      D->src_atom, D->dst_atom, D->surf, D->avgd, D->distinct, F->pf, F->active, F->mature, F->owner, D->psrc, D->pdst);
    E->launches += 6;
    // morph.cpp:1050-1054
    double freedom = std::sqrt((double) E->width * E->width + (double) E->height * E->height) * 1.0 / 4.0;
    double freedom_radius = freedom * (1.0 - tl);
    for (unsigned s = 0; s < E->p.fluid; ++s) {
        rcode = fluid_step(E, E->p.fluid - (s + 1), freedom_radius);
        if (rcode != AMX_OK) return rcode;
    }
    cudaMemsetAsync(D->acc, 0, np * 5 * 8, E->stream);
    k_fd_splat<<<div_up(n, 128), 128, 0, E->stream>>>(P, F->pf, F->active, F->mature, D->acc, np);
    k_fd_resolve<<<div_up(np, 256), 256, 0, E->stream>>>(P, D->acc, np, D->img[0], D->has[0]);
    E->launches += 2;
    int cur = 0;
    for (uint32_t layer = 0; layer < P.feather; ++layer) {
        double alpha = 1.0 - ((double) (layer + 1) / (double) (P.feather + 1));
        k_fd_feather<<<div_up(np, 256), 256, 0, E->stream>>>(P, alpha, D->img[cur], D->has[cur], D->img[cur ^ 1], D->has[cur ^ 1], np);
        E->launches++;
        cur ^= 1;
    }
    k_fd_final<<<div_up(np, 256), 256, 0, E->stream>>>(P, D->img[cur], D->has[cur], d_bg, d_dst, np);
    E->launches++;
    return E->check("fluid frame") ? AMX_ERR_CUDA : AMX_OK;
}

} // namespace amx
