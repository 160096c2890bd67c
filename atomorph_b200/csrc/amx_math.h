/*
 * amx_math.h -- pure value functions shared by the sm_100a kernels and the host facade.
 *
 * Everything here is a restatement, from the algorithm, of the reference's small numeric
 * helpers (SURVEY.md row a-N).  All double arithmetic keeps the reference's operation order
 * and is compiled WITHOUT fused multiply-add contraction (nvcc -fmad=false, g++
 * -ffp-contract=off) so floor/round decisions agree bit for bit with the x86-64 build of the
 * reference.  IEEE add/mul/div/sqrt are correctly rounded on both sides.
 */
#ifndef AMX_MATH_H
#define AMX_MATH_H

#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define AMX_HD __host__ __device__ __forceinline__
#else
#define AMX_HD inline
#endif

namespace amx {

// ------------------------------------------------------------------ constants (reference atomorph.h:235-248,304-306)
enum : unsigned { K_RGB = 0, K_HSP = 1, K_NONE = 2, K_LINEAR = 3, K_SPLINE = 4, K_COSINE = 5, K_PERLIN = 6 };
enum : unsigned { ST_BLOB_DETECTION = 0, ST_BLOB_UNIFICATION = 1, ST_BLOB_MATCHING = 2, ST_ATOM_MORPHING = 3, ST_DONE = 4 };
enum : unsigned { SHOW_TEXTURE = 0, SHOW_AVERAGE = 1, SHOW_DISTINCT = 2 };
enum : unsigned { F_HAS_PIXEL = 1, F_HAS_FLUID = 2 };

// ------------------------------------------------------------------ key point word (reference atomorph.h:275-284)
// u64: x | y<<16 | x_fract<<32 | y_fract<<40 | flags<<48   (byte 7 unused, kept 0)
typedef uint64_t pword;
AMX_HD uint32_t pw_x(pword w) { return (uint32_t) (w & 0xffffu); }
AMX_HD uint32_t pw_y(pword w) { return (uint32_t) ((w >> 16) & 0xffffu); }
AMX_HD uint32_t pw_xf(pword w) { return (uint32_t) ((w >> 32) & 0xffu); }
AMX_HD uint32_t pw_yf(pword w) { return (uint32_t) ((w >> 40) & 0xffu); }
AMX_HD uint32_t pw_flags(pword w) { return (uint32_t) ((w >> 48) & 0xffu); }
AMX_HD pword pw_make(uint32_t x, uint32_t y, uint32_t xf, uint32_t yf, uint32_t flags) {
    return (pword) (x & 0xffffu) | ((pword) (y & 0xffffu) << 16) | ((pword) (xf & 0xffu) << 32) |
           ((pword) (yf & 0xffu) << 40) | ((pword) (flags & 0xffu) << 48);
}
// sub-pixel coordinates in 1/256 px units
AMX_HD int32_t pw_x256(pword w) { return (int32_t) (pw_x(w) * 256u + pw_xf(w)); }
AMX_HD int32_t pw_y256(pword w) { return (int32_t) (pw_y(w) * 256u + pw_yf(w)); }

// squared travel distance in 1/256 px units -- reference atomorph.h:334-339 (point_distance)
AMX_HD uint64_t point_distance(pword a, pword b) {
    int64_t dx = (int64_t) pw_x256(a) - (int64_t) pw_x256(b);
    int64_t dy = (int64_t) pw_y256(a) - (int64_t) pw_y256(b);
    return (uint64_t) (dx * dx) + (uint64_t) (dy * dy);
}

// ------------------------------------------------------------------ colours, packed r | g<<8 | b<<16 | a<<24
AMX_HD uint32_t c_r(uint32_t c) { return c & 255u; }
AMX_HD uint32_t c_g(uint32_t c) { return (c >> 8) & 255u; }
AMX_HD uint32_t c_b(uint32_t c) { return (c >> 16) & 255u; }
AMX_HD uint32_t c_a(uint32_t c) { return (c >> 24) & 255u; }
AMX_HD uint32_t c_make(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
    return (r & 255u) | ((g & 255u) << 8) | ((b & 255u) << 16) | ((a & 255u) << 24);
}
// double -> uint8_t the way the reference's implicit conversions behave on x86-64 for in-range values
AMX_HD uint32_t to_u8(double v) { return (uint32_t) ((int32_t) v) & 255u; }

// reference color.cpp:22-29  create_color(double...) = round(v*255)
AMX_HD uint32_t create_color_d(double r, double g, double b, double a) {
    return c_make(to_u8(round(r * 255.0)), to_u8(round(g * 255.0)), to_u8(round(b * 255.0)), to_u8(round(a * 255.0)));
}

// reference color.h:17-24
AMX_HD double color_distance(uint32_t c1, uint32_t c2) {
    int rd = (int) c_r(c1) - (int) c_r(c2);
    int gd = (int) c_g(c1) - (int) c_g(c2);
    int bd = (int) c_b(c1) - (int) c_b(c2);
    int ad = (int) c_a(c1) - (int) c_a(c2);
    return sqrt((double) (rd * rd + gd * gd + bd * bd + ad * ad)) / 510.0;
}

// HSP colour model (Darel Rex Finley 2006, public domain; alienryderflex.com/hsp.html), as used by
// reference color.cpp:31-59 (byte wrappers) and 71-178 (the model).  Perceived-brightness weights:
#define AMX_PR 0.299
#define AMX_PG 0.587
#define AMX_PB 0.114

AMX_HD void rgb_to_hsp_d(double R, double G, double B, double *H, double *S, double *P) {
    *P = sqrt(R * R * AMX_PR + G * G * AMX_PG + B * B * AMX_PB);
    if (R == G && R == B) { *H = 0.; *S = 0.; return; }
    if (R >= G && R >= B) {          // R largest
        if (B >= G) { *H = 6. / 6. - 1. / 6. * (B - G) / (R - G); *S = 1. - G / R; }
        else        { *H = 0. / 6. + 1. / 6. * (G - B) / (R - B); *S = 1. - B / R; }
    } else if (G >= R && G >= B) {   // G largest
        if (R >= B) { *H = 2. / 6. - 1. / 6. * (R - B) / (G - B); *S = 1. - B / G; }
        else        { *H = 2. / 6. + 1. / 6. * (B - R) / (G - R); *S = 1. - R / G; }
    } else {                         // B largest
        if (G >= R) { *H = 4. / 6. - 1. / 6. * (G - R) / (B - R); *S = 1. - R / B; }
        else        { *H = 4. / 6. + 1. / 6. * (R - G) / (B - G); *S = 1. - G / B; }
    }
}

// One sextant of the inverse: returns lo (smallest), hi (largest), mid channel.
AMX_HD void hsp_sextant(double Hs, double P, double mom, double wl, double wm, double wh, double *lo, double *mid, double *hi) {
    // lo = P / sqrt(wh/mom/mom + wm*part*part + wl), hi = lo/mom, mid = lo + Hs*(hi-lo)
    double part = 1. + Hs * (1. / mom - 1.);
    *lo = P / sqrt(wh / mom / mom + wm * part * part + wl);
    *hi = (*lo) / mom;
    *mid = (*lo) + Hs * ((*hi) - (*lo));
}
AMX_HD void hsp_sextant0(double Hs, double P, double wm, double wh, double *mid, double *hi) {
    // saturation 1: lo = 0, hi = sqrt(P*P/(wh + wm*Hs*Hs)), mid = hi*Hs
    *hi = sqrt(P * P / (wh + wm * Hs * Hs));
    *mid = (*hi) * Hs;
}

AMX_HD void hsp_to_rgb_d(double H, double S, double P, double *R, double *G, double *B) {
    double mom = 1. - S;   // minOverMax
    if (mom > 0.) {
        if (H < 1. / 6.)      { H = 6. * (H - 0. / 6.);  hsp_sextant(H, P, mom, AMX_PB, AMX_PG, AMX_PR, B, G, R); }   // R>G>B
        else if (H < 2. / 6.) { H = 6. * (-H + 2. / 6.); hsp_sextant(H, P, mom, AMX_PB, AMX_PR, AMX_PG, B, R, G); }   // G>R>B
        else if (H < 3. / 6.) { H = 6. * (H - 2. / 6.);  hsp_sextant(H, P, mom, AMX_PR, AMX_PB, AMX_PG, R, B, G); }   // G>B>R
        else if (H < 4. / 6.) { H = 6. * (-H + 4. / 6.); hsp_sextant(H, P, mom, AMX_PR, AMX_PG, AMX_PB, R, G, B); }   // B>G>R
        else if (H < 5. / 6.) { H = 6. * (H - 4. / 6.);  hsp_sextant(H, P, mom, AMX_PG, AMX_PR, AMX_PB, G, R, B); }   // B>R>G
        else                  { H = 6. * (-H + 6. / 6.); hsp_sextant(H, P, mom, AMX_PG, AMX_PB, AMX_PR, G, B, R); }   // R>B>G
    } else {
        if (H < 1. / 6.)      { H = 6. * (H - 0. / 6.);  hsp_sextant0(H, P, AMX_PG, AMX_PR, G, R); *B = 0.; }
        else if (H < 2. / 6.) { H = 6. * (-H + 2. / 6.); hsp_sextant0(H, P, AMX_PR, AMX_PG, R, G); *B = 0.; }
        else if (H < 3. / 6.) { H = 6. * (H - 2. / 6.);  hsp_sextant0(H, P, AMX_PB, AMX_PG, B, G); *R = 0.; }
        else if (H < 4. / 6.) { H = 6. * (-H + 4. / 6.); hsp_sextant0(H, P, AMX_PG, AMX_PB, G, B); *R = 0.; }
        else if (H < 5. / 6.) { H = 6. * (H - 4. / 6.);  hsp_sextant0(H, P, AMX_PR, AMX_PB, R, B); *G = 0.; }
        else                  { H = 6. * (-H + 6. / 6.); hsp_sextant0(H, P, AMX_PB, AMX_PR, B, R); *G = 0.; }
    }
}

// byte wrappers, reference color.cpp:31-59 (alpha passes through; inverse clamps to 255)
AMX_HD uint32_t rgb_to_hsp(uint32_t c) {
    double h, s, p;
    rgb_to_hsp_d(c_r(c) / 255.0, c_g(c) / 255.0, c_b(c) / 255.0, &h, &s, &p);
    return c_make(to_u8(round(h * 255.0)), to_u8(round(s * 255.0)), to_u8(round(p * 255.0)), c_a(c));
}
AMX_HD uint32_t hsp_to_rgb(uint32_t c) {
    double r, g, b;
    hsp_to_rgb_d(c_r(c) / 255.0, c_g(c) / 255.0, c_b(c) / 255.0, &r, &g, &b);
    return c_make(to_u8(fmin(round(r * 255.0), 255.0)), to_u8(fmin(round(g * 255.0), 255.0)),
                  to_u8(fmin(round(b * 255.0), 255.0)), c_a(c));
}

// ------------------------------------------------------------------ interpolation (reference morph.cpp:1467-1515)
// colour: channels round(w*c1 + (1-w)*c2)
AMX_HD uint32_t lerp_color(uint32_t c1, uint32_t c2, double w) {
    double iw = 1.0 - w;
    return c_make(to_u8(round(w * (double) c_r(c1) + iw * (double) c_r(c2))),
                  to_u8(round(w * (double) c_g(c1) + iw * (double) c_g(c2))),
                  to_u8(round(w * (double) c_b(c1) + iw * (double) c_b(c2))),
                  to_u8(round(w * (double) c_a(c1) + iw * (double) c_a(c2))));
}
// cosine / Perlin easing of the c1 weight `str` (morph.cpp:1467-1476); cos_fn lets the caller pick libm / device cos
#define AMX_PI_REF 3.14159265358
template <typename CosFn>
AMX_HD double ease_strength(double lag, double slope, double str, CosFn cos_fn) {
    double s = (slope + 0.1) / 1.1;
    double l = (1.0 - s) * lag;
    if (str <= l) return 0.0;
    if (str >= (l + s)) return 1.0;
    return ((-cos_fn((str - l) * (AMX_PI_REF / s)) + 1.0) / 2.0);
}
// key point: linear in 1/256 px units with truncation (morph.cpp:1501-1515); flags of the result are 0 here
AMX_HD void lerp_point(pword p1, pword p2, double w, uint32_t *x, uint32_t *y, uint32_t *xf, uint32_t *yf) {
    double x1 = 256.0 * (double) pw_x(p1) + (double) pw_xf(p1);
    double y1 = 256.0 * (double) pw_y(p1) + (double) pw_yf(p1);
    double x2 = 256.0 * (double) pw_x(p2) + (double) pw_xf(p2);
    double y2 = 256.0 * (double) pw_y(p2) + (double) pw_yf(p2);
    double xx = (w * x1 + (1.0 - w) * x2);
    double yy = (w * y1 + (1.0 - w) * y2);
    uint32_t ix = (uint32_t) ((int32_t) (xx / 256.0)) & 0xffffu;
    uint32_t iy = (uint32_t) ((int32_t) (yy / 256.0)) & 0xffffu;
    *x = ix; *y = iy;
    *xf = to_u8(xx - (double) ((int32_t) ix * 256));
    *yf = to_u8(yy - (double) ((int32_t) iy * 256));
}

// ------------------------------------------------------------------ Catmull-Rom (reference spline.cpp:29-57)
// One coordinate of Eq(): p1*b1 + p2*b2 + p3*b3 + p4*b4, left to right.
AMX_HD void cr_basis(double t, double *b1, double *b2, double *b3, double *b4) {
    double t2 = t * t;
    double t3 = t2 * t;
    *b1 = 0.5 * (-t3 + 2.0 * t2 - t);
    *b2 = 0.5 * (3.0 * t3 - 5.0 * t2 + 2.0);
    *b3 = 0.5 * (-3.0 * t3 + 4.0 * t2 + t);
    *b4 = 0.5 * (t3 - t2);
}
AMX_HD double cr_eval(double p1, double p2, double p3, double p4, double b1, double b2, double b3, double b4) {
    return ((p1 * b1 + p2 * b2) + p3 * b3) + p4 * b4;
}
// interval index and local time of GetInterpolatedSplinePoint for n control points
AMX_HD void cr_locate(double t, int n, int *p0, int *p1, int *p2, int *p3, double *lt) {
    double delta_t = 1.0 / (double) n;
    int p = (int) (t / delta_t);
    int q;
    q = p - 1; *p0 = (q < 0 ? n - 1 : (q >= n ? q - n : q));
    q = p;     *p1 = (q < 0 ? n - 1 : (q >= n ? q - n : q));
    q = p + 1; *p2 = (q < 0 ? n - 1 : (q >= n ? q - n : q));
    q = p + 2; *p3 = (q < 0 ? n - 1 : (q >= n ? q - n : q));
    *lt = (t - delta_t * (double) p) / delta_t;
}
// control point coordinate of a key point (morph.cpp:195-196): x + x_fract/256
AMX_HD double pw_xd(pword w) { return (double) pw_x(w) + (double) pw_xf(w) / 256.0; }
AMX_HD double pw_yd(pword w) { return (double) pw_y(w) + (double) pw_yf(w) / 256.0; }
// spline sample -> integer pixel + fract*255 rounded (morph.cpp:527-530)
AMX_HD void split_spline_coord(double v, uint32_t *i, uint32_t *f) {
    double integ;
    double fract = modf(v, &integ);
    *i = (uint32_t) ((int32_t) integ) & 0xffffu;
    *f = to_u8(round(fract * 255.0));
}

// ------------------------------------------------------------------ improved Perlin noise (reference perlin.cpp:19-89)
// p = 512-entry permutation generated on the HOST with std::shuffle(mt19937(seed)) (perlin.cpp:11-17).
AMX_HD double pn_fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
AMX_HD double pn_lerp(double t, double a, double b) { return a + t * (b - a); }
AMX_HD double pn_grad(int hash, double x, double y, double z) {
    int h = hash & 15;
    double u = h < 8 ? x : y, v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
    return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}
template <typename P>
AMX_HD double pn_noise(const P *p, double x, double y, double z) {
    int X = (int) floor(x) & 255;
    int Y = (int) floor(y) & 255;
    int Z = (int) floor(z) & 255;
    x -= floor(x); y -= floor(y); z -= floor(z);
    double u = pn_fade(x), v = pn_fade(y), w = pn_fade(z);
    int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
    int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
    return pn_lerp(w,
                   pn_lerp(v, pn_lerp(u, pn_grad(p[AA], x, y, z), pn_grad(p[BA], x - 1, y, z)),
                           pn_lerp(u, pn_grad(p[AB], x, y - 1, z), pn_grad(p[BB], x - 1, y - 1, z))),
                   pn_lerp(v, pn_lerp(u, pn_grad(p[AA + 1], x, y, z - 1), pn_grad(p[BA + 1], x - 1, y, z - 1)),
                           pn_lerp(u, pn_grad(p[AB + 1], x, y - 1, z - 1), pn_grad(p[BB + 1], x - 1, y - 1, z - 1))));
}
template <typename P>
AMX_HD double pn_octave2(const P *p, double x, double y, int octaves) {
    double result = 0.0, amp = 1.0;
    for (int i = 0; i < octaves; ++i) {
        result += pn_noise(p, x, y, 0.0) * amp;
        x *= 2.0; y *= 2.0; amp *= 0.5;
    }
    return result;
}

// ------------------------------------------------------------------ counter-based RNG (device-side pairing masks, fracts, shuffles)
// splitmix64 finaliser over (seed, stream, counter): stateless, reproducible on host and device.
AMX_HD uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
AMX_HD uint64_t rng64(uint64_t seed, uint64_t stream, uint64_t counter) {
    return mix64(mix64(seed ^ (stream * 0xd1342543de82ef95ull)) ^ (counter * 0x2545f4914f6cdd1dull));
}

} // namespace amx
#endif
