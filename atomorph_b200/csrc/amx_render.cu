/*
 * amx_render.cu -- K6: the per-frame renderer (SURVEY.md row a-R).
 *
 * Reference: morph::get_pixels(t) -> draw_atoms -> get_pixels(blob,t)  (morph.cpp:452-678,
 * 1302-1421), get_background (1431-1465).  The reference is a scatter into per-pixel std::map
 * lists followed by per-pixel normalisation IN ATOM ORDER, per-blob feather peeling and
 * cross-blob "over" compositing.  Here a frame is two kernels:
 *
 *   prepare  (once per table refresh)  per atom & interval: end colours with the one-sided
 *            alpha rule resolved, Perlin lag/slope  -> coalesced SoA, 24 B/atom/interval
 *   scatter  per atom: trajectory (linear / Catmull-Rom in double, reference operation order),
 *            colour fade; the atom is linked into the list of its HOME pixel (top-left splat
 *            target) with ONE 32-bit atomicExch; its 16-byte record {next, colour, fract} is
 *            written at index = atom (coalesced, no allocator)
 *   gather   per pixel: walk the lists of the 4 homes that can reach the pixel, keep the
 *            contributions (atom, colour, integer bilinear numerator n <= 65025), sort them by
 *            blob order and atom index and replay the reference's double sums in ITS order
 *            (morph.cpp:598-613): bit-exact, including the exact .5 ties where an order-free
 *            integer sum would differ by 1 LSB.  Without feather the same thread composites the
 *            blobs "over" each other and blends the background (morph.cpp:1357-1401) and
 *            writes the RGBA pixel: no accumulator ever touches HBM.
 *   feather  (only when feather > 0) per-(pixel, blob) entries, 4-neighbour erosion layers
 *            (morph.cpp:625-674), then a composite kernel.
 *
 * A pixel that receives more than MAXK contributions (atoms of a volatile blob collapsing
 * onto one point) falls back to exact integer sums sum(c*n)/sum(n) with exact rational
 * rounding -- identical to the reference except on exact .5 ties (<= 1 LSB there).
 *
 * Algorithmic bytes per frame: A*24 B (h <= 3 or linear) or A*40 B
 * (spline, h >= 4) read + P*4 B written.
 */
#include <cmath>
#include <cstring>
#include <random>
#include <numeric>
#include <algorithm>
#include "amx_engine.h"

namespace amx {

#define MAXK 32
#define NIL 0xffffffffu

struct RConst {
    uint32_t width, height, cw, ch;
    uint32_t bx1, by1, bx2, by2;
    uint32_t motion, fading, density, show_blobs, keep_background;
    uint32_t nchains, h;
    uint64_t A;
    uint32_t ovf_mask;     // ovf_cap - 1
    uint32_t feather;
};

struct RFrame {
    uint32_t y, yn;
    int      p0, p1, p2, p3;      // Catmull-Rom control columns
    double   b1, b2, b3, b4;      // Catmull-Rom basis at the local time
    double   w;                   // c1 / pt1 weight = 1 - local t
    double   str_cos;             // eased weight for COSINE fading (host libm)
    int32_t  chain_only;          // >= 0: only this chain (per-blob fetch)
};

// per-(pixel, blob) entries, used by the feather / per-blob paths
struct Acc {
    int32_t *owner;               // chain stored in the layer-0 entry of a pixel, -1 none
    uint8_t *hasovf;
    unsigned long long *ovf_key;  // open addressing, key = pixel<<32 | chain+1
    uint32_t *ovf_used;
    size_t canvas;
    size_t ovf_cap;
};

__device__ __forceinline__ uint32_t hash_pix(uint32_t ci) { return (uint32_t) (mix64(ci) >> 17); }

// find or insert the overflow slot of (pixel ci, chain c)
__device__ __forceinline__ uint32_t ovf_slot(const Acc &ac, uint32_t mask, uint32_t ci, uint32_t c, bool insert) {
    unsigned long long key = ((unsigned long long) ci << 32) | (unsigned long long) (c + 1u);
    uint32_t s = hash_pix(ci) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        unsigned long long k = ac.ovf_key[s];
        if (k == key) return s;
        if (k == 0ull) {
            if (!insert) return 0xffffffffu;
            unsigned long long prev = atomicCAS(&ac.ovf_key[s], 0ull, key);
            if (prev == 0ull) { atomicAdd(ac.ovf_used, 1u); return s; }
            if (prev == key) return s;
        }
        s = (s + 1) & mask;
    }
    return 0xffffffffu;
}

struct DevCos { __device__ double operator()(double x) const { return cos(x); } };

// ---------------------------------------------------------------------------------------- scatter
__global__ void __launch_bounds__(256)
k_scatter(const pword *__restrict__ table, const uint32_t *__restrict__ rc1, const uint32_t *__restrict__ rc2,
          const double *__restrict__ rlag, const double *__restrict__ rslope, const uint32_t *__restrict__ chain_of,
          RConst rc, RFrame rf, uint32_t *__restrict__ head, uint4 *__restrict__ rec) {
    size_t a = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= rc.A) return;
    if (rf.chain_only >= 0 && chain_of[a] != (uint32_t) rf.chain_only) return;
    size_t A = rc.A;
    pword pt1 = table[(size_t) rf.y * A + a];
    pword pt2 = table[(size_t) rf.yn * A + a];
    bool has1 = pw_flags(pt1) & F_HAS_PIXEL, has2 = pw_flags(pt2) & F_HAS_PIXEL;
    if (!has1 && !has2) return;

    // trajectory (morph.cpp:523-531)
    uint32_t x, y, xf, yf;
    if (rc.motion == K_LINEAR) {
        lerp_point(pt1, pt2, rf.w, &x, &y, &xf, &yf);
    } else if (rc.motion == K_SPLINE) {
        pword q0 = table[(size_t) rf.p0 * A + a], q1 = table[(size_t) rf.p1 * A + a];
        pword q2 = table[(size_t) rf.p2 * A + a], q3 = table[(size_t) rf.p3 * A + a];
        double vx = cr_eval(pw_xd(q0), pw_xd(q1), pw_xd(q2), pw_xd(q3), rf.b1, rf.b2, rf.b3, rf.b4);
        double vy = cr_eval(pw_yd(q0), pw_yd(q1), pw_yd(q2), pw_yd(q3), rf.b1, rf.b2, rf.b3, rf.b4);
        split_spline_coord(vx, &x, &xf);
        split_spline_coord(vy, &y, &yf);
    } else {
        x = pw_x(pt1); y = pw_y(pt1); xf = pw_xf(pt1); yf = pw_yf(pt1);
    }
    // clip (morph.cpp:552-555)
    if (x >= rc.width || y >= rc.height) {
        if (x > rc.bx2 || x < rc.bx1 || y > rc.by2 || y < rc.by1) return;
    }
    // colour (morph.cpp:537-550)
    uint32_t c1 = rc1[(size_t) rf.y * A + a], c2 = rc2[(size_t) rf.y * A + a];
    double str = rf.w;
    if (rc.fading == K_COSINE) str = rf.str_cos;
    else if (rc.fading == K_PERLIN) str = ease_strength(rlag[(size_t) rf.y * A + a], rslope[(size_t) rf.y * A + a], rf.w, DevCos());
    uint32_t col = lerp_color(c1, c2, str);

    uint32_t home = y * rc.cw + x;
    uint32_t next = atomicExch(&head[home], (uint32_t) a);
    rec[a] = make_uint4(next, col, xf | (yf << 8), 0u);
}

// ---------------------------------------------------------------------------------------- gather
// Visit every contribution to pixel (px, py): f(atom, colour, n) with n the integer bilinear numerator.
// Splat targets and their edge rules: morph.cpp:558-588.
template <typename F>
__device__ __forceinline__ void visit_contributions(const uint32_t *__restrict__ head, const uint4 *__restrict__ rec, const RConst &rc,
                                                    uint32_t px, uint32_t py, F f) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t dx = k & 1, dy = k >> 1;
        if (px < dx || py < dy) continue;
        uint32_t hx = px - dx, hy = py - dy;
        bool ok;
        if (k == 0) ok = true;
        else if (k == 1) ok = (hx < rc.bx2 || hx + 1 < rc.width);
        else if (k == 2) ok = (hy < rc.by2 || hy + 1 < rc.height);
        else ok = (hy < rc.by2 && hx < rc.bx2) || (hy + 1 < rc.height && hx + 1 < rc.width);
        if (!ok) continue;
        uint32_t idx = head[hy * rc.cw + hx];
        while (idx != NIL) {
            uint4 r = rec[idx];
            uint32_t xf = r.z & 255u, yf = (r.z >> 8) & 255u;
            uint32_t n = (dx ? xf : 255u - xf) * (dy ? yf : 255u - yf);
            if (n) f(idx, r.y, n);
            idx = r.x;
        }
    }
}

// the reference's per-position normalisation, contributions already in atom order (morph.cpp:598-613)
__device__ __forceinline__ uint32_t resolve_fp(const uint32_t *cc, const uint32_t *cn, int first, int last, uint32_t density) {
    double r = 0.0, g = 0.0, b = 0.0, a = 0.0, weight_sum = 0.0;
    for (int i = first; i < last; ++i) {
        double w = (double) cn[i] / 65025.0;
        uint32_t c = cc[i];
        weight_sum += w;
        r += (double) c_r(c) * w;
        g += (double) c_g(c) * w;
        b += (double) c_b(c) * w;
        a += (double) c_a(c) * w;
    }
    double wd = density > 0 ? (double) (last - first) / (double) density : 0.0;
    if (wd > 1.0) wd = 1.0;
    r = round(r / weight_sum);
    g = round(g / weight_sum);
    b = round(b / weight_sum);
    a = round(wd * (a / weight_sum));
    return c_make(to_u8(r), to_u8(g), to_u8(b), to_u8(a));
}

// exact round-half-up of num/den for non-negative integers (fallback for pixels with > MAXK contributions)
__device__ __forceinline__ uint32_t rdiv(unsigned long long num, unsigned long long den) {
    return (uint32_t) ((2ull * num + den) / (2ull * den));
}
__device__ __forceinline__ uint32_t resolve_int(unsigned long long R, unsigned long long G, unsigned long long B, unsigned long long Av,
                                                unsigned long long N, unsigned long long cnt, uint32_t density) {
    uint32_t r = rdiv(R, N), g = rdiv(G, N), b = rdiv(B, N), a;
    if (density == 0) a = 0;
    else if (cnt >= density) a = rdiv(Av, N);
    else a = (uint32_t) ((2ull * Av * cnt + N * density) / (2ull * N * density));
    return c_make(r, g, b, a);
}

// final colour of an entry: feather alpha (morph.cpp:658-669) and show_blobs substitution (1320-1340)
__device__ __forceinline__ uint32_t entry_color(uint32_t px, uint32_t layer, uint32_t c, const RConst &rc, uint32_t y_frame,
                                               const uint32_t *blob_avg, const uint32_t *blob_distinct) {
    if (rc.feather > 0 && layer < 254u) {
        double a = round((double) c_a(px) * ((double) (layer + 1u) / (double) (rc.feather + 1u)));
        px = (px & 0x00ffffffu) | (to_u8(a) << 24);
    }
    if (rc.show_blobs == SHOW_DISTINCT) return blob_distinct[c];
    if (rc.show_blobs == SHOW_AVERAGE) return blob_avg[(size_t) y_frame * rc.nchains + c];
    return px;
}

// cross-blob "over" accumulation in arrival order (morph.cpp:1342-1380)
struct Over {
    double r = 0, g = 0, b = 0, a = 0;
    bool first = true;
    __device__ __forceinline__ void add(uint32_t col) {
        if (c_a(col) == 0) return;
        double sr = c_r(col) / 255.0, sg = c_g(col) / 255.0, sb = c_b(col) / 255.0, sa = c_a(col) / 255.0;
        if (first) { r = sr; g = sg; b = sb; a = sa; first = false; }
        else {
            r = sa * sr + (1.0 - sa) * r;
            g = sa * sg + (1.0 - sa) * g;
            b = sa * sb + (1.0 - sa) * b;
            a = a + (1.0 - a) * sa;
        }
    }
    __device__ __forceinline__ uint32_t finish(uint32_t bgc, bool keep_background) {
        if (first) return bgc;
        if (keep_background) {                      // morph.cpp:1388-1399
            double bgr = c_r(bgc) / 255.0, bgg = c_g(bgc) / 255.0, bgb = c_b(bgc) / 255.0, bga = c_a(bgc) / 255.0;
            r = a * r + (1.0 - a) * bgr;
            g = a * g + (1.0 - a) * bgg;
            b = a * b + (1.0 - a) * bgb;
            a = bga + (1.0 - bga) * a;
        }
        return create_color_d(r, g, b, a);
    }
};

// Emits the resolved blob pixels of one position in ascending blob order: emit(chain, px)
template <bool SINGLE, typename E>
__device__ __forceinline__ void resolve_position(const uint32_t *__restrict__ head, const uint4 *__restrict__ rec, const RConst &rc,
                                                 const uint32_t *__restrict__ chain_of, const int32_t *__restrict__ boc,
                                                 uint32_t px, uint32_t py, E emit) {
    unsigned long long key[MAXK];
    uint32_t cc[MAXK], cn[MAXK];
    int count = 0;
    visit_contributions(head, rec, rc, px, py, [&](uint32_t a, uint32_t col, uint32_t n) {
        if (count < MAXK) {
            unsigned long long k = a;
            if (!SINGLE) k |= (unsigned long long) (uint32_t) boc[chain_of[a]] << 32;
            // insertion sort by (blob order, atom)
            int i = count;
            while (i > 0 && key[i - 1] > k) { key[i] = key[i - 1]; cc[i] = cc[i - 1]; cn[i] = cn[i - 1]; --i; }
            key[i] = k; cc[i] = col; cn[i] = n;
        }
        ++count;
    });
    if (count == 0) return;
    if (count <= MAXK) {
        int first = 0;
        while (first < count) {
            int last = first + 1;
            if (!SINGLE) while (last < count && (key[last] >> 32) == (key[first] >> 32)) ++last;
            else last = count;
            uint32_t chain = SINGLE ? 0u : chain_of[(uint32_t) key[first]];
            emit(chain, resolve_fp(cc, cn, first, last, rc.density));
            first = last;
        }
        return;
    }
    // heavy position: exact integer sums, one chain at a time in ascending blob order
    long long prev = -1;
    for (;;) {
        long long best = LLONG_MAX;
        uint32_t bchain = 0;
        if (SINGLE) { if (prev < 0) best = 0; }
        else visit_contributions(head, rec, rc, px, py, [&](uint32_t a, uint32_t, uint32_t) {
            uint32_t c = chain_of[a];
            long long k = boc[c];
            if (k > prev && k < best) { best = k; bchain = c; }
        });
        if (best == LLONG_MAX) break;
        unsigned long long R = 0, G = 0, B = 0, Av = 0, N = 0, cnt = 0;
        visit_contributions(head, rec, rc, px, py, [&](uint32_t a, uint32_t col, uint32_t n) {
            if (!SINGLE && chain_of[a] != bchain) return;
            R += c_r(col) * n; G += c_g(col) * n; B += c_b(col) * n; Av += c_a(col) * n; N += n; ++cnt;
        });
        emit(bchain, resolve_int(R, G, B, Av, N, cnt, rc.density));
        prev = best;
    }
}

// exact round(num/den) (half up) for num < 2^40, den < 2^31, quotient <= 255: float estimate + integer fix-up.
// *tie is set when num/den is an exact .5 tie (the only place where the reference's double sums can differ).
__device__ __forceinline__ uint32_t rdiv_small(unsigned long long num, unsigned long long den, bool *tie) {
    unsigned long long n2 = 2ull * num + den, d2 = 2ull * den;
    uint32_t q = (uint32_t) __fmul_rz(__ull2float_rz(n2), __frcp_rz(__ull2float_ru(d2)));   // every step rounds down: never above the true quotient
    unsigned long long rem = n2 - (unsigned long long) q * d2;
    while (rem >= d2) { rem -= d2; ++q; }
    *tie |= (rem == 0ull);
    return q;
}

// FAST PATH of the fused gather: one chain at the position and no exact tie -> integer sums, no sort, no
// per-contribution double math.  Returns false when the generic ordered replay is needed.
template <bool SINGLE>
__device__ __forceinline__ bool resolve_fast(const uint32_t *__restrict__ head, const uint4 *__restrict__ rec, const RConst &rc,
                                             const uint32_t *__restrict__ chain_of, uint32_t px, uint32_t py, uint32_t *chain_out,
                                             uint32_t *px_out, bool *empty) {
    unsigned long long R = 0, G = 0, B = 0, Av = 0;
    uint32_t N = 0, cnt = 0, chain = 0xffffffffu;
    bool mixed = false;
    visit_contributions(head, rec, rc, px, py, [&](uint32_t a, uint32_t col, uint32_t n) {
        if (!SINGLE) {
            uint32_t c = chain_of[a];
            if (chain == 0xffffffffu) chain = c;
            else if (c != chain) mixed = true;
        }
        R += c_r(col) * n; G += c_g(col) * n; B += c_b(col) * n; Av += c_a(col) * n; N += n; ++cnt;
    });
    *empty = (cnt == 0);
    if (cnt == 0) return true;
    if (mixed || cnt > 30000u) return false;       // N = sum(n) must stay below 2^31
    bool tie = false;
    uint32_t r = rdiv_small(R, N, &tie), g = rdiv_small(G, N, &tie), b = rdiv_small(B, N, &tie), a;
    if (rc.density == 0) a = 0;
    else if (cnt >= rc.density) a = rdiv_small(Av, N, &tie);
    else {
        unsigned long long num = Av * cnt, den = (unsigned long long) N * rc.density;   // round(cnt*A / (density*N))
        unsigned long long n2 = 2ull * num + den, d2 = 2ull * den;
        unsigned long long q = n2 / d2;
        tie |= (n2 - q * d2 == 0ull);
        a = (uint32_t) q;
    }
    if (tie) return false;
    *chain_out = SINGLE ? 0u : chain;
    *px_out = c_make(r, g, b, a);
    return true;
}

// fused gather + composite (feather == 0): one thread per OUTPUT pixel
template <bool SINGLE>
__global__ void __launch_bounds__(256)
k_gather_composite(const uint32_t *__restrict__ head, const uint4 *__restrict__ rec, RConst rc, uint32_t y_frame,
                   const uint32_t *__restrict__ chain_of, const int32_t *__restrict__ blob_of_chain,
                   const uint32_t *__restrict__ blob_avg, const uint32_t *__restrict__ blob_distinct,
                   const uint32_t *__restrict__ bg, uint32_t *__restrict__ out) {
    uint32_t px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= rc.width || py >= rc.height) return;
    size_t i = (size_t) py * rc.width + px;
    uint32_t bgc = rc.keep_background ? bg[i] : 0u;
    uint32_t chain = 0, pxl = 0;
    bool empty = false;
    if (resolve_fast<SINGLE>(head, rec, rc, chain_of, px, py, &chain, &pxl, &empty)) {
        if (empty) { out[i] = bgc; return; }
        uint32_t col = entry_color(pxl, 255u, chain, rc, y_frame, blob_avg, blob_distinct);
        if (!rc.keep_background) { out[i] = c_a(col) ? col : 0u; return; }   // round((c/255.0)*255.0) == c for every byte c
        Over ov;
        ov.add(col);
        out[i] = ov.finish(bgc, true);
        return;
    }
    // generic path: several blobs at the position, an exact tie, or a very long list
    const int32_t *boc = blob_of_chain + (size_t) y_frame * rc.nchains;
    Over ov;
    resolve_position<SINGLE>(head, rec, rc, chain_of, boc, px, py, [&](uint32_t ch, uint32_t p) {
        ov.add(entry_color(p, 255u, ch, rc, y_frame, blob_avg, blob_distinct));
    });
    out[i] = ov.finish(bgc, rc.keep_background != 0);
}

// gather into per-(pixel, blob) entries (feather / per-blob fetch): one thread per CANVAS pixel
template <bool SINGLE>
__global__ void __launch_bounds__(256)
k_gather_entries(const uint32_t *__restrict__ head, const uint4 *__restrict__ rec, RConst rc, uint32_t y_frame,
                 const uint32_t *__restrict__ chain_of, const int32_t *__restrict__ blob_of_chain, Acc ac,
                 uint32_t *__restrict__ px0, uint8_t *__restrict__ layer0, uint32_t *__restrict__ pxo, uint8_t *__restrict__ layero) {
    uint32_t px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= rc.cw || py >= rc.ch) return;
    uint32_t ci = py * rc.cw + px;
    const int32_t *boc = blob_of_chain + (size_t) y_frame * rc.nchains;
    int emitted = 0;
    resolve_position<SINGLE>(head, rec, rc, chain_of, boc, px, py, [&](uint32_t chain, uint32_t pxl) {
        if (emitted == 0) { ac.owner[ci] = (int32_t) chain; px0[ci] = pxl; layer0[ci] = 255; }
        else {
            uint32_t s = ovf_slot(ac, rc.ovf_mask, ci, chain, true);
            if (s != 0xffffffffu) { pxo[s] = pxl; layero[s] = 255; ac.hasovf[ci] = 1; }
        }
        ++emitted;
    });
    if (emitted == 0) { ac.owner[ci] = -1; layer0[ci] = 254; }
}

// layer value of entry (pixel ci, chain c): 254 none, 255 unpeeled, else peel iteration
template <bool SINGLE>
__device__ __forceinline__ uint32_t entry_layer(const Acc &ac, uint32_t mask, const uint8_t *layer0, const uint8_t *layero, uint32_t ci, uint32_t c) {
    if (SINGLE) return layer0[ci];
    if (ac.owner[ci] == (int32_t) c) return layer0[ci];
    if (!ac.hasovf[ci]) return 254;
    uint32_t s = ovf_slot(ac, mask, ci, c, false);
    return s == 0xffffffffu ? 254u : layero[s];
}

// one erosion pass l (morph.cpp:627-656): an unpeeled entry is border if x==0 || y==0 or any
// 4-neighbour entry of the same blob is missing or was peeled in an earlier pass
template <bool SINGLE>
__global__ void __launch_bounds__(256)
k_feather_pass(Acc ac, RConst rc, uint8_t *layer0, uint8_t *layero, uint32_t l, size_t total) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t ci, c;
    uint8_t *mine;
    if (i < ac.canvas) {
        ci = (uint32_t) i; mine = layer0 + i;
        if (*mine != 255) return;
        c = SINGLE ? 0u : (uint32_t) ac.owner[ci];
    } else {
        size_t s = i - ac.canvas;
        unsigned long long k = ac.ovf_key[s];
        if (k == 0ull) return;
        mine = layero + s;
        if (*mine != 255) return;
        ci = (uint32_t) (k >> 32); c = (uint32_t) (k & 0xffffffffu) - 1u;
    }
    uint32_t x = ci % rc.cw, y = ci / rc.cw;
    bool border = (x == 0 || y == 0 || x == 65535u || y == 65535u);
    if (!border) {
        uint32_t nb[4];
        nb[0] = (x + 1 < rc.cw) ? entry_layer<SINGLE>(ac, rc.ovf_mask, layer0, layero, ci + 1, c) : 254u;
        nb[1] = entry_layer<SINGLE>(ac, rc.ovf_mask, layer0, layero, ci - 1, c);
        nb[2] = (y + 1 < rc.ch) ? entry_layer<SINGLE>(ac, rc.ovf_mask, layer0, layero, ci + rc.cw, c) : 254u;
        nb[3] = entry_layer<SINGLE>(ac, rc.ovf_mask, layer0, layero, ci - rc.cw, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) border |= (nb[k] == 254u) || (nb[k] < l);
    }
    if (border) *mine = (uint8_t) l;
}

template <bool SINGLE>
__global__ void __launch_bounds__(256)
k_composite(Acc ac, RConst rc, uint32_t y_frame, const uint32_t *__restrict__ px0, const uint8_t *__restrict__ layer0,
            const uint32_t *__restrict__ pxo, const uint8_t *__restrict__ layero, const int32_t *__restrict__ blob_of_chain,
            const uint32_t *__restrict__ blob_avg, const uint32_t *__restrict__ blob_distinct,
            const uint32_t *__restrict__ bg, uint32_t *__restrict__ out) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) rc.width * rc.height) return;
    uint32_t x = (uint32_t) (i % rc.width), y = (uint32_t) (i / rc.width);
    uint32_t ci = y * rc.cw + x;
    uint32_t bgc = rc.keep_background ? bg[i] : 0u;
    Over ov;
    auto over = [&](uint32_t col) { ov.add(col); };

    if (SINGLE) {
        uint32_t l = layer0[ci];
        if (l != 254u) over(entry_color(px0[ci], l, 0, rc, y_frame, blob_avg, blob_distinct));
    } else {
        int32_t o = ac.owner[ci];
        if (o >= 0) {
            if (!ac.hasovf[ci]) {
                uint32_t l = layer0[ci];
                if (l != 254u) over(entry_color(px0[ci], l, (uint32_t) o, rc, y_frame, blob_avg, blob_distinct));
            } else {
                // several blobs: visit in ascending blob-vector index (morph.cpp:1309: b = 0,1,...)
                const int32_t *boc = blob_of_chain + (size_t) y_frame * rc.nchains;
                int64_t prev = -1;
                for (;;) {
                    int64_t best = INT64_MAX; uint32_t bpx = 0, bl = 254, bc = 0;
                    if (layer0[ci] != 254u) {
                        int64_t k = boc[o];
                        if (k > prev && k < best) { best = k; bpx = px0[ci]; bl = layer0[ci]; bc = (uint32_t) o; }
                    }
                    uint32_t s = hash_pix(ci) & rc.ovf_mask;
                    for (uint32_t probe = 0; probe <= rc.ovf_mask; ++probe) {
                        unsigned long long key = ac.ovf_key[s];
                        if (key == 0ull) break;
                        if ((uint32_t) (key >> 32) == ci && layero[s] != 254u) {
                            uint32_t c = (uint32_t) (key & 0xffffffffu) - 1u;
                            int64_t k = boc[c];
                            if (k > prev && k < best) { best = k; bpx = pxo[s]; bl = layero[s]; bc = c; }
                        }
                        s = (s + 1) & rc.ovf_mask;
                    }
                    if (best == INT64_MAX) break;
                    over(entry_color(bpx, bl, bc, rc, y_frame, blob_avg, blob_distinct));
                    prev = best;
                }
            }
        }
    }
    out[i] = ov.finish(bgc, rc.keep_background != 0);
}

// clear ownership after a frame
__global__ void __launch_bounds__(256) k_clear_owner(int32_t *owner, uint8_t *hasovf, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { owner[i] = -1; hasovf[i] = 0; }
}

// background cross-dissolve (morph.cpp:1431-1465)
__global__ void __launch_bounds__(256)
k_background(const uint32_t *__restrict__ f1, const uint32_t *__restrict__ f2, RConst rc, double w, double str_cos,
             const int32_t *__restrict__ perlin, uint32_t *__restrict__ out) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) rc.width * rc.height) return;
    uint32_t x = (uint32_t) (i % rc.width), y = (uint32_t) (i / rc.width);
    uint32_t ci = y * rc.cw + x;
    uint32_t c1 = f1[ci], c2 = f2[ci];
    double str = w;
    if (rc.fading == K_COSINE) str = str_cos;
    else if (rc.fading == K_PERLIN) {
        double f = 8.0;
        double bbox_w = (double) ((int) rc.bx2 - (int) rc.bx1) + 1.0;
        double bbox_h = (double) ((int) rc.by2 - (int) rc.by1) + 1.0;
        double px = ((double) ((int) x - (int) rc.bx1) / bbox_w) * f;
        double py = ((double) ((int) y - (int) rc.by1) / bbox_h) * f;
        double lag = pn_octave2(perlin, px, py, 8) * 0.5 + 0.5;
        double slope = pn_octave2(perlin + 512, px, py, 8) * 0.5 + 0.5;
        str = ease_strength(lag, slope, w, DevCos());
    }
    out[i] = lerp_color(c1, c2, str);
}

// prepare: per atom & interval end colours + Perlin lag/slope (morph.cpp:495-548)
__global__ void __launch_bounds__(256)
k_prepare(const pword *__restrict__ table, const uint32_t *__restrict__ fetch_y, const uint32_t *__restrict__ fetch_yn,
          uint32_t y, uint32_t yn, RConst rc, const int32_t *__restrict__ perlin, uint32_t *__restrict__ rc1,
          uint32_t *__restrict__ rc2, double *__restrict__ rlag, double *__restrict__ rslope) {
    size_t a = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= rc.A) return;
    pword pt1 = table[(size_t) y * rc.A + a], pt2 = table[(size_t) yn * rc.A + a];
    bool has1 = pw_flags(pt1) & F_HAS_PIXEL, has2 = pw_flags(pt2) & F_HAS_PIXEL;
    uint32_t c1 = 0, c2 = 0;
    auto fetch = [&](const uint32_t *img, pword p) -> uint32_t {
        uint32_t x = pw_x(p), yy = pw_y(p);
        return (x < rc.cw && yy < rc.ch) ? img[(size_t) yy * rc.cw + x] : 0u;
    };
    if (has1 && !has2) { c1 = fetch(fetch_y, pt1); c2 = c1 & 0x00ffffffu; }
    else if (has2 && !has1) { c2 = fetch(fetch_yn, pt2); c1 = c2 & 0x00ffffffu; }
    else if (has1 && has2) { c1 = fetch(fetch_y, pt1); c2 = fetch(fetch_yn, pt2); }
    rc1[(size_t) y * rc.A + a] = c1;
    rc2[(size_t) y * rc.A + a] = c2;
    if (rlag) {
        double f = 8.0;
        double bbox_w = (double) ((int) rc.bx2 - (int) rc.bx1) + 1.0;
        double bbox_h = (double) ((int) rc.by2 - (int) rc.by1) + 1.0;
        double px = ((double) (((int) pw_x(pt1) - (int) rc.bx1) * 256 + (int) pw_xf(pt1)) / (bbox_w * 256.0)) * f;
        double py = ((double) (((int) pw_y(pt1) - (int) rc.by1) * 256 + (int) pw_yf(pt1)) / (bbox_h * 256.0)) * f;
        rlag[(size_t) y * rc.A + a] = pn_octave2(perlin, px, py, 8) * 0.5 + 0.5;
        rslope[(size_t) y * rc.A + a] = pn_octave2(perlin + 512, px, py, 8) * 0.5 + 0.5;
    }
}

// ------------------------------------------------------------------------------------------ host side

void engine_render_free(Engine *E) {
    dev_free(E->rc1); dev_free(E->rc2); dev_free(E->rlag); dev_free(E->rslope);
    E->rc1 = E->rc2 = nullptr; E->rlag = E->rslope = nullptr;
    dev_free(E->d_blob_of_chain); dev_free(E->d_blob_avg); dev_free(E->d_blob_distinct);
    E->d_blob_of_chain = nullptr; E->d_blob_avg = nullptr; E->d_blob_distinct = nullptr;
    dev_free(E->acc_owner); dev_free(E->acc_hasovf); dev_free(E->ovf_key);
    dev_free(E->d_ovf_used); dev_free(E->blob_px); dev_free(E->ab_head); dev_free(E->ab_rec);
    E->acc_owner = nullptr; E->acc_hasovf = nullptr; E->ovf_key = nullptr;
    E->d_ovf_used = nullptr; E->blob_px = nullptr; E->ovf_cap = 0; E->ab_head = nullptr; E->ab_rec = nullptr;
    E->render_ready = false;
}

static void perlin_table(unsigned seed, int32_t *p) {
    // reference perlin.cpp:11-17 -- libstdc++ std::shuffle with mt19937, generated on the host (SURVEY.md section 9 note 11)
    if (seed == 0) seed = std::mt19937::default_seed;
    int tmp[256];
    std::iota(tmp, tmp + 256, 0);
    std::shuffle(tmp, tmp + 256, std::mt19937(seed));
    for (int i = 0; i < 256; ++i) { p[i] = tmp[i]; p[256 + i] = tmp[i]; }
}

static RConst make_rconst(Engine *E) {
    RConst rc;
    rc.width = E->width; rc.height = E->height; rc.cw = E->cw; rc.ch = E->ch;
    rc.bx1 = E->bbox[0]; rc.by1 = E->bbox[1]; rc.bx2 = E->bbox[2]; rc.by2 = E->bbox[3];
    rc.motion = E->p.motion; rc.fading = E->p.fading; rc.density = E->p.density; rc.show_blobs = E->p.show_blobs;
    rc.keep_background = E->p.keep_background ? 1u : 0u;
    rc.nchains = E->nchains; rc.h = E->h; rc.A = E->A;
    rc.ovf_mask = E->ovf_cap ? E->ovf_cap - 1 : 0;
    rc.feather = (uint32_t) std::min<uint64_t>(E->p.feather, 253);
    return rc;
}

static int ensure_perlin(Engine *E) {
    if (!E->d_perlin && !dev_alloc(E, (void **) &E->d_perlin, 1024 * sizeof(int32_t), "perlin")) return AMX_ERR_NOMEM;
    if (E->perlin_seed_loaded != E->p.seed) {
        int32_t host[1024];
        perlin_table(E->p.seed, host);            // lag_map   = PerlinNoise(seed)    morph.cpp:436-444
        perlin_table(E->p.seed + 1, host + 512);  // slope_map = PerlinNoise(seed+1)
        if (E->fail(cudaMemcpyAsync(E->d_perlin, host, sizeof host, cudaMemcpyHostToDevice, E->stream), "perlin H2D")) return AMX_ERR_CUDA;
        cudaStreamSynchronize(E->stream);
        E->perlin_seed_loaded = E->p.seed;
    }
    return AMX_OK;
}

int engine_render_prepare(Engine *E) {
    if (E->nchains == 0 || E->A == 0 || E->h == 0) { E->err = "no chains"; return AMX_ERR_STATE; }
    if (E->frames.size() != E->h) { E->err = "chain height != frame count"; return AMX_ERR_STATE; }
    size_t n = (size_t) E->h * E->A;
    bool perlin = E->p.fading == K_PERLIN;
    if (!E->rc1) {
        if (!dev_alloc(E, (void **) &E->rc1, n * 4, "rc1") || !dev_alloc(E, (void **) &E->rc2, n * 4, "rc2")) return AMX_ERR_NOMEM;
    }
    if (perlin && !E->rlag) {
        if (!dev_alloc(E, (void **) &E->rlag, n * 8, "rlag") || !dev_alloc(E, (void **) &E->rslope, n * 8, "rslope")) return AMX_ERR_NOMEM;
    }
    int rcode = ensure_perlin(E);
    if (rcode != AMX_OK) return rcode;
    size_t cv = E->canvas();
    if (!E->ab_head) {
        // A-buffer: list heads per canvas position + one 16-byte record per atom
        if (!dev_alloc(E, (void **) &E->ab_head, cv * 4, "abuf heads") || !dev_alloc(E, (void **) &E->ab_rec, E->A * 16, "abuf records"))
            return AMX_ERR_NOMEM;
    }
    // blob order / colours per (frame, chain)
    size_t m = (size_t) E->h * E->nchains;
    std::vector<int32_t> boc(m, 0x7fffffff);
    std::vector<uint32_t> avg(m, 0), distinct(E->nchains, 0);
    for (uint32_t y = 0; y < E->h; ++y) {
        FrameDev &f = E->frames[y];
        for (size_t b = 0; b < f.blobs.size(); ++b) {
            uint64_t g = f.blobs[b].group;
            // chain with key g
            auto it = std::find(E->chain_key.begin(), E->chain_key.end(), g);
            if (it == E->chain_key.end()) continue;
            size_t c = it - E->chain_key.begin();
            if (boc[y * E->nchains + c] != 0x7fffffff) continue;
            boc[y * E->nchains + c] = (int32_t) b;
            const double *s = f.blobs[b].stats;
            uint32_t col = create_color_d(s[2], s[3], s[4], s[5]);        // blob2pixel, morph.cpp:1423-1429
            if (E->p.blob_delimiter == K_HSP) col = hsp_to_rgb(col);
            avg[y * E->nchains + c] = col;
        }
        if (f.blobs.empty() && E->nchains == 1) boc[y] = 0;
    }
    for (uint32_t c = 0; c < E->nchains; ++c) {
        std::mt19937 gen((unsigned) E->chain_key[c]);                       // morph.cpp:1322-1326
        std::uniform_int_distribution<unsigned char> dist(0, 255);
        unsigned rr = dist(gen), gg = dist(gen), bb = dist(gen);
        distinct[c] = c_make(rr, gg, bb, 255);
    }
    if (!E->d_blob_of_chain) {
        if (!dev_alloc(E, (void **) &E->d_blob_of_chain, m * 4, "boc") || !dev_alloc(E, (void **) &E->d_blob_avg, m * 4, "avg") ||
            !dev_alloc(E, (void **) &E->d_blob_distinct, (size_t) E->nchains * 4, "distinct"))
            return AMX_ERR_NOMEM;
    }
    cudaMemcpyAsync(E->d_blob_of_chain, boc.data(), m * 4, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(E->d_blob_avg, avg.data(), m * 4, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(E->d_blob_distinct, distinct.data(), (size_t) E->nchains * 4, cudaMemcpyHostToDevice, E->stream);

    RConst rc = make_rconst(E);
    for (uint32_t y = 0; y < E->h; ++y) {
        uint32_t yn = (y + 1) % E->h;
        k_prepare<<<div_up(E->A, 256), 256, 0, E->stream>>>(E->table, E->frames[y].fetch, E->frames[yn].fetch, y, yn, rc, E->d_perlin,
                                                           E->rc1, E->rc2, perlin ? E->rlag : nullptr, perlin ? E->rslope : nullptr);
        E->launches++;
    }
    if (E->fail(cudaStreamSynchronize(E->stream), "render prepare") || E->check("render prepare")) return AMX_ERR_CUDA;
    E->render_ready = true;
    return AMX_OK;
}

struct LibmCos { double operator()(double x) const { return ::cos(x); } };

// time -> frame index, local t (morph.cpp:404-423, 446-464)
static bool locate_frame(Engine *E, double t, double *time_out, uint32_t *f_out, double *tl_out) {
    double integ, time = std::modf(t, &integ);
    if (time < 0.0) time += 1.0;
    size_t nf = E->frames.size();
    if (nf == 0) return false;
    size_t f = (size_t) (time * (double) nf);
    if (f >= nf) return false;                         // reference: get_frame_key -> SIZE_MAX -> nothing drawn
    double dt = 1.0 / double(nf);
    double tl = std::max(0.0, (time - ((double) E->frames[f].key * dt)) / dt);   // key used as a number (morph.cpp:463-464)
    *time_out = time; *f_out = (uint32_t) f; *tl_out = tl;
    return true;
}

static RFrame make_rframe(Engine *E, double time, uint32_t f, double tl) {
    RFrame rf;
    rf.y = f; rf.yn = (f + 1) % E->h;
    double lt;
    cr_locate(time, (int) E->h, &rf.p0, &rf.p1, &rf.p2, &rf.p3, &lt);
    cr_basis(lt, &rf.b1, &rf.b2, &rf.b3, &rf.b4);
    rf.w = 1.0 - tl;
    rf.str_cos = ease_strength(0.5, 0.5, rf.w, LibmCos());
    rf.chain_only = -1;
    return rf;
}

// entry storage for the feather / per-blob paths, allocated on first use
static int ensure_entries(Engine *E) {
    if (E->acc_owner) return AMX_OK;
    size_t cv = E->canvas();
    E->ovf_cap = 1u << 12;
    if (E->nchains > 1) while (E->ovf_cap < cv && E->ovf_cap < (1u << 26)) E->ovf_cap <<= 1;
    if (!dev_alloc(E, (void **) &E->acc_owner, cv * 4, "owner") || !dev_alloc(E, (void **) &E->acc_hasovf, cv, "hasovf") ||
        !dev_alloc(E, (void **) &E->ovf_key, (size_t) E->ovf_cap * 8, "ovf_key") || !dev_alloc(E, (void **) &E->d_ovf_used, 4, "ovf_used") ||
        !dev_alloc(E, (void **) &E->blob_px, (cv + E->ovf_cap) * 5, "blob_px"))
        return AMX_ERR_NOMEM;
    cudaMemsetAsync(E->acc_owner, 0xff, cv * 4, E->stream);
    cudaMemsetAsync(E->acc_hasovf, 0, cv, E->stream);
    cudaMemsetAsync(E->ovf_key, 0, (size_t) E->ovf_cap * 8, E->stream);
    cudaMemsetAsync(E->d_ovf_used, 0, 4, E->stream);
    return AMX_OK;
}

static Acc make_acc(Engine *E) {
    Acc ac;
    ac.owner = E->acc_owner; ac.hasovf = E->acc_hasovf; ac.ovf_key = E->ovf_key;
    ac.ovf_used = E->d_ovf_used; ac.canvas = E->canvas(); ac.ovf_cap = E->ovf_cap;
    return ac;
}

static int ensure_out(Engine *E, uint64_t words) {
    if (E->d_out_cap >= words) return AMX_OK;
    dev_free(E->d_out); E->d_out = nullptr; E->d_out_cap = 0;
    if (!dev_alloc(E, (void **) &E->d_out, words * 4, "out staging")) return AMX_ERR_NOMEM;
    E->d_out_cap = words;
    return AMX_OK;
}

// background of one frame into d_dst (width*height)
static void launch_background(Engine *E, const RConst &rc, const RFrame &rf, uint32_t *d_dst) {
    size_t np = (size_t) E->width * E->height;
    k_background<<<div_up(np, 256), 256, 0, E->stream>>>(E->frames[rf.y].fetch, E->frames[rf.yn].fetch, rc, rf.w, rf.str_cos, E->d_perlin, d_dst);
    E->launches++;
}

static dim3 grid2d(uint32_t w, uint32_t h) { return dim3(div_up(w, 32), div_up(h, 8)); }

// scatter one frame into the A-buffer
static void launch_scatter(Engine *E, const RConst &rc, const RFrame &rf) {
    cudaMemsetAsync(E->ab_head, 0xff, E->canvas() * 4, E->stream);
    k_scatter<<<div_up(E->A, 256), 256, 0, E->stream>>>(E->table, E->rc1, E->rc2, E->rlag, E->rslope, E->chain_of, rc, rf, E->ab_head, E->ab_rec);
    E->launches++;
}

// scatter + gather into entries + feather of one frame; leaves entries (px/layer) valid and ownership set
static void launch_frame_entries(Engine *E, const RConst &rc, const RFrame &rf, const Acc &ac) {
    size_t cv = E->canvas();
    uint32_t *px0 = E->blob_px, *pxo = E->blob_px + cv;
    uint8_t *layer0 = (uint8_t *) (E->blob_px + cv + E->ovf_cap), *layero = layer0 + cv;
    bool single = E->nchains == 1;
    launch_scatter(E, rc, rf);
    const int32_t *boc = E->d_blob_of_chain;
    if (single) k_gather_entries<true><<<grid2d(rc.cw, rc.ch), dim3(32, 8), 0, E->stream>>>(E->ab_head, E->ab_rec, rc, rf.y, E->chain_of, boc, ac, px0, layer0, pxo, layero);
    else        k_gather_entries<false><<<grid2d(rc.cw, rc.ch), dim3(32, 8), 0, E->stream>>>(E->ab_head, E->ab_rec, rc, rf.y, E->chain_of, boc, ac, px0, layer0, pxo, layero);
    E->launches++;
    size_t total = single ? cv : cv + E->ovf_cap;
    for (uint32_t l = 0; l < rc.feather; ++l) {
        if (single) k_feather_pass<true><<<div_up(total, 256), 256, 0, E->stream>>>(ac, rc, layer0, layero, l, total);
        else        k_feather_pass<false><<<div_up(total, 256), 256, 0, E->stream>>>(ac, rc, layer0, layero, l, total);
        E->launches++;
    }
}

static void launch_frame_cleanup(Engine *E) {
    if (E->nchains == 1) return;
    size_t cv = E->canvas();
    cudaMemsetAsync(E->acc_hasovf, 0, cv, E->stream);
    cudaMemsetAsync(E->ovf_key, 0, (size_t) E->ovf_cap * 8, E->stream);
}

int engine_render(Engine *E, const double *times, uint32_t n, uint32_t *out, int out_is_device) {
    if (E->width == 0 || E->height == 0) { E->err = "resolution not set"; return AMX_ERR_STATE; }
    size_t np = (size_t) E->width * E->height;
    bool have_chains = E->nchains > 0 && E->A > 0;
    if (have_chains && !E->render_ready) {
        int rcode = engine_render_prepare(E);
        if (rcode != AMX_OK) return rcode;
    }
    if (E->p.keep_background) { int rcode = ensure_perlin(E); if (rcode != AMX_OK) return rcode; }
    uint32_t *d_dst = out;
    if (!out_is_device) {
        int rcode = ensure_out(E, np * n);
        if (rcode != AMX_OK) return rcode;
        d_dst = E->d_out;
    }
    uint32_t *d_bg = nullptr;
    if (E->p.keep_background && !dev_alloc(E, (void **) &d_bg, np * 4, "bg")) return AMX_ERR_NOMEM;
    if (have_chains && E->p.feather > 0) { int rcode = ensure_entries(E); if (rcode != AMX_OK) { dev_free(d_bg); return rcode; } }
    RConst rc = make_rconst(E);
    Acc ac = make_acc(E);
    size_t cv = E->canvas();
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t *dst = d_dst + (size_t) i * np;
        double time, tl; uint32_t f;
        if (!locate_frame(E, times[i], &time, &f, &tl)) { cudaMemsetAsync(dst, 0, np * 4, E->stream); continue; }
        RFrame rf;
        if (have_chains) rf = make_rframe(E, time, f, tl);
        else { rf.y = f; rf.yn = (f + 1) % (uint32_t) E->frames.size(); rf.w = 1.0 - tl; rf.str_cos = ease_strength(0.5, 0.5, rf.w, LibmCos()); }
        if (E->p.keep_background) launch_background(E, rc, rf, d_bg);
        if (!have_chains) {
            if (E->p.keep_background) cudaMemcpyAsync(dst, d_bg, np * 4, cudaMemcpyDeviceToDevice, E->stream);
            else cudaMemsetAsync(dst, 0, np * 4, E->stream);
            continue;
        }
        bool single = E->nchains == 1;
        if (rc.feather == 0) {
            launch_scatter(E, rc, rf);
            if (single) k_gather_composite<true><<<grid2d(rc.width, rc.height), dim3(32, 8), 0, E->stream>>>(E->ab_head, E->ab_rec, rc, rf.y, E->chain_of, E->d_blob_of_chain, E->d_blob_avg, E->d_blob_distinct, d_bg, dst);
            else        k_gather_composite<false><<<grid2d(rc.width, rc.height), dim3(32, 8), 0, E->stream>>>(E->ab_head, E->ab_rec, rc, rf.y, E->chain_of, E->d_blob_of_chain, E->d_blob_avg, E->d_blob_distinct, d_bg, dst);
            E->launches++;
            continue;
        }
        launch_frame_entries(E, rc, rf, ac);
        uint32_t *px0 = E->blob_px, *pxo = E->blob_px + cv;
        uint8_t *layer0 = (uint8_t *) (E->blob_px + cv + E->ovf_cap), *layero = layer0 + cv;
        if (single)
            k_composite<true><<<div_up(np, 256), 256, 0, E->stream>>>(ac, rc, rf.y, px0, layer0, pxo, layero, E->d_blob_of_chain, E->d_blob_avg, E->d_blob_distinct, d_bg, dst);
        else
            k_composite<false><<<div_up(np, 256), 256, 0, E->stream>>>(ac, rc, rf.y, px0, layer0, pxo, layero, E->d_blob_of_chain, E->d_blob_avg, E->d_blob_distinct, d_bg, dst);
        E->launches++;
        launch_frame_cleanup(E);
    }
    int rcode = AMX_OK;
    if (!out_is_device) {
        if (E->fail(cudaMemcpyAsync(out, d_dst, np * n * 4, cudaMemcpyDeviceToHost, E->stream), "render D2H")) rcode = AMX_ERR_CUDA;
    }
    if (!out_is_device || d_bg) {
        if (E->fail(cudaStreamSynchronize(E->stream), "render")) rcode = AMX_ERR_CUDA;
    }
    if (E->check("render")) rcode = AMX_ERR_CUDA;
    dev_free(d_bg);
    if (rcode == AMX_OK && E->nchains > 1 && !out_is_device && E->d_ovf_used) {
        uint32_t used = 0;
        cudaMemcpy(&used, E->d_ovf_used, 4, cudaMemcpyDeviceToHost);
        cudaMemsetAsync(E->d_ovf_used, 0, 4, E->stream);
        if ((uint64_t) used > (uint64_t) E->ovf_cap * n) { E->err = "overflow table exhausted"; rcode = AMX_ERR_NOMEM; }
    }
    return rcode;
}

int engine_background(Engine *E, double t, uint32_t *out, int out_is_device) {
    if (E->width == 0 || E->height == 0) { E->err = "resolution not set"; return AMX_ERR_STATE; }
    size_t np = (size_t) E->width * E->height;
    int rcode = ensure_perlin(E);
    if (rcode != AMX_OK) return rcode;
    uint32_t *d_dst = out;
    if (!out_is_device) { rcode = ensure_out(E, np); if (rcode != AMX_OK) return rcode; d_dst = E->d_out; }
    double time, tl; uint32_t f;
    if (!locate_frame(E, t, &time, &f, &tl)) cudaMemsetAsync(d_dst, 0, np * 4, E->stream);
    else {
        RConst rc = make_rconst(E);
        RFrame rf; rf.y = f; rf.yn = (f + 1) % (uint32_t) E->frames.size(); rf.w = 1.0 - tl; rf.str_cos = ease_strength(0.5, 0.5, rf.w, LibmCos());
        launch_background(E, rc, rf, d_dst);
    }
    if (!out_is_device) {
        if (E->fail(cudaMemcpyAsync(out, d_dst, np * 4, cudaMemcpyDeviceToHost, E->stream), "bg D2H") || E->fail(cudaStreamSynchronize(E->stream), "bg")) return AMX_ERR_CUDA;
    }
    return E->check("background") ? AMX_ERR_CUDA : AMX_OK;
}

// per-blob fetch (morph.cpp:452-678): entries of one chain, emitted in the reference's order
int engine_render_blob(Engine *E, uint32_t blob, double t, uint64_t cap, uint16_t *xy, uint32_t *rgba, int64_t *n_out, uint64_t *group) {
    *n_out = -1;
    double time, tl; uint32_t f;
    if (!locate_frame(E, t, &time, &f, &tl)) return AMX_OK;
    FrameDev &fr = E->frames[f];
    if (blob >= fr.blobs.size()) return AMX_OK;
    uint64_t g = fr.blobs[blob].group;
    if (group) *group = g;
    if (E->nchains == 0) {
        // no chains yet: the blob's own pixels (morph.cpp:469-475)
        if (!fr.blob_pix || fr.blob_pix_off.size() <= blob + 1) { *n_out = 0; return AMX_OK; }
        uint64_t b0 = fr.blob_pix_off[blob], b1 = fr.blob_pix_off[blob + 1];
        std::vector<uint32_t> pos(b1 - b0), img(E->canvas());
        cudaMemcpy(pos.data(), fr.blob_pix + b0, (b1 - b0) * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(img.data(), fr.fetch, E->canvas() * 4, cudaMemcpyDeviceToHost);
        uint64_t k = 0;
        for (uint32_t ci : pos) { if (k < cap) { xy[2 * k] = ci % E->cw; xy[2 * k + 1] = ci / E->cw; rgba[k] = img[ci]; } ++k; }
        *n_out = (int64_t) k;
        return AMX_OK;
    }
    auto it = std::find(E->chain_key.begin(), E->chain_key.end(), g);
    if (it == E->chain_key.end()) return AMX_OK;        // reference returns nullptr
    uint32_t c = (uint32_t) (it - E->chain_key.begin());
    if (!E->render_ready) { int rcode = engine_render_prepare(E); if (rcode != AMX_OK) return rcode; }
    { int rcode = ensure_entries(E); if (rcode != AMX_OK) return rcode; }
    RConst rc = make_rconst(E);
    Acc ac = make_acc(E);
    RFrame rf = make_rframe(E, time, f, tl);
    rf.chain_only = (int32_t) c;
    launch_frame_entries(E, rc, rf, ac);
    size_t cv = E->canvas();
    std::vector<uint32_t> px(cv);
    std::vector<uint8_t> layer(cv);
    // with chain_only every entry of this chain owns its pixel, so layer-0 arrays hold the whole blob
    cudaMemcpyAsync(px.data(), E->blob_px, cv * 4, cudaMemcpyDeviceToHost, E->stream);
    cudaMemcpyAsync(layer.data(), (uint8_t *) (E->blob_px + cv + E->ovf_cap), cv, cudaMemcpyDeviceToHost, E->stream);
    launch_frame_cleanup(E);
    if (E->fail(cudaStreamSynchronize(E->stream), "render blob") || E->check("render blob")) return AMX_ERR_CUDA;
    uint64_t k = 0;
    auto emit = [&](uint32_t ci, uint32_t col) {
        if (k < cap) { xy[2 * k] = (uint16_t) (ci % E->cw); xy[2 * k + 1] = (uint16_t) (ci / E->cw); rgba[k] = col; }
        ++k;
    };
    uint32_t F = rc.feather;
    if (F == 0) {
        for (size_t ci = 0; ci < cv; ++ci) if (layer[ci] != 254) emit((uint32_t) ci, px[ci]);
    } else {
        for (uint32_t l = 0; l < F; ++l)
            for (size_t ci = 0; ci < cv; ++ci)
                if (layer[ci] == l) {
                    double a = std::round((double) c_a(px[ci]) * ((double) (l + 1) / (double) (F + 1)));
                    emit((uint32_t) ci, (px[ci] & 0x00ffffffu) | (to_u8(a) << 24));
                }
        for (size_t ci = 0; ci < cv; ++ci) if (layer[ci] == 255) emit((uint32_t) ci, px[ci]);
    }
    *n_out = (int64_t) k;
    return AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_render_prepare(amx_ctx *ctx) {
    if (!ctx) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_render_prepare(&ctx->e);
}
int amx_render(amx_ctx *ctx, const double *times, uint32_t n, uint32_t *out, int out_is_device) {
    if (!ctx || !times || !out) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_render(&ctx->e, times, n, out, out_is_device);
}
int amx_render_blob(amx_ctx *ctx, uint32_t blob, double t, uint64_t cap, uint16_t *xy_out, uint32_t *rgba_out, int64_t *n, uint64_t *group) {
    if (!ctx || !n) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_render_blob(&ctx->e, blob, t, cap, xy_out, rgba_out, n, group);
}
int amx_background(amx_ctx *ctx, double t, uint32_t *out, int out_is_device) {
    if (!ctx || !out) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return engine_background(&ctx->e, t, out, out_is_device);
}

}
