/*
 * am_oracle.cpp -- TEST INFRASTRUCTURE ONLY: a plain, serial CPU restatement of the reference's
 * algorithms on the morph hot path (1Hyena/atomorph), written from the algorithm, each function
 * citing the reference file:line it follows.  Built by `make -C oracle port` into
 * oracle/libamoracle.so and bound by oracle/amoracle.py.
 *
 * PINNING: every function here is checked against the UNMODIFIED reference compiled from source
 * (oracle/_ref/libamref.so) in tests/test_oracle_pin.py, and against the golden vectors under
 * tests/golden/ that tests/golden/make_golden.py generated from that same reference build
 * (the reference ships no golden vectors of its own -- SURVEY.md section 8c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (atomorph_b200/) never does.
 *
 * Conventions: colours packed r | g<<8 | b<<16 | a<<24; key points packed
 * x | y<<16 | x_fract<<32 | y_fract<<40 | flags<<48; chain tables column-major words[j*width+x].
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------ small helpers
struct Col { int r, g, b, a; };
inline Col unpack(uint32_t c) { return Col{int(c & 255), int((c >> 8) & 255), int((c >> 16) & 255), int((c >> 24) & 255)}; }
inline uint32_t pack(int r, int g, int b, int a) { return uint32_t(r & 255) | (uint32_t(g & 255) << 8) | (uint32_t(b & 255) << 16) | (uint32_t(a & 255) << 24); }
inline uint8_t u8(double v) { return (uint8_t) v; }          // the reference's implicit double -> unsigned char

struct Pt { int x, y, xf, yf, flags; };
inline Pt unpackp(uint64_t w) { return Pt{int(w & 0xffff), int((w >> 16) & 0xffff), int((w >> 32) & 255), int((w >> 40) & 255), int((w >> 48) & 255)}; }
inline uint64_t packp(int x, int y, int xf, int yf, int fl) {
    return uint64_t(x & 0xffff) | (uint64_t(y & 0xffff) << 16) | (uint64_t(xf & 255) << 32) | (uint64_t(yf & 255) << 40) | (uint64_t(fl & 255) << 48);
}

// color.cpp:22-29
uint32_t color_from_doubles(double r, double g, double b, double a) {
    return pack(u8(round(r * 255.0)), u8(round(g * 255.0)), u8(round(b * 255.0)), u8(round(a * 255.0)));
}

// HSP model, color.cpp:71-178 (Darel Rex Finley, public domain)
const double Pr = 0.299, Pg = 0.587, Pb = 0.114;

void rgb2hsp(double R, double G, double B, double *H, double *S, double *P) {
    *P = sqrt(R * R * Pr + G * G * Pg + B * B * Pb);
    if (R == G && R == B) { *H = 0.; *S = 0.; return; }
    if (R >= G && R >= B) {
        if (B >= G) { *H = 6. / 6. - 1. / 6. * (B - G) / (R - G); *S = 1. - G / R; }
        else { *H = 0. / 6. + 1. / 6. * (G - B) / (R - B); *S = 1. - B / R; }
    } else if (G >= R && G >= B) {
        if (R >= B) { *H = 2. / 6. - 1. / 6. * (R - B) / (G - B); *S = 1. - B / G; }
        else { *H = 2. / 6. + 1. / 6. * (B - R) / (G - R); *S = 1. - R / G; }
    } else {
        if (G >= R) { *H = 4. / 6. - 1. / 6. * (G - R) / (B - R); *S = 1. - R / B; }
        else { *H = 4. / 6. + 1. / 6. * (R - G) / (B - G); *S = 1. - G / B; }
    }
}

// the six sextants share one shape: (lo, mid, hi) channels with their perceived weights
void hsp2rgb(double H, double S, double P, double *R, double *G, double *B) {
    double mom = 1. - S;
    double *lo, *mid, *hi, wl, wm, wh, h;
    if (H < 1. / 6.)      { h = 6. * (H - 0. / 6.);  hi = R; mid = G; lo = B; wh = Pr; wm = Pg; wl = Pb; }
    else if (H < 2. / 6.) { h = 6. * (-H + 2. / 6.); hi = G; mid = R; lo = B; wh = Pg; wm = Pr; wl = Pb; }
    else if (H < 3. / 6.) { h = 6. * (H - 2. / 6.);  hi = G; mid = B; lo = R; wh = Pg; wm = Pb; wl = Pr; }
    else if (H < 4. / 6.) { h = 6. * (-H + 4. / 6.); hi = B; mid = G; lo = R; wh = Pb; wm = Pg; wl = Pr; }
    else if (H < 5. / 6.) { h = 6. * (H - 4. / 6.);  hi = B; mid = R; lo = G; wh = Pb; wm = Pr; wl = Pg; }
    else                  { h = 6. * (-H + 6. / 6.); hi = R; mid = B; lo = G; wh = Pr; wm = Pb; wl = Pg; }
    if (mom > 0.) {
        double part = 1. + h * (1. / mom - 1.);
        *lo = P / sqrt(wh / mom / mom + wm * part * part + wl);
        *hi = (*lo) / mom;
        *mid = (*lo) + h * ((*hi) - (*lo));
    } else {
        *hi = sqrt(P * P / (wh + wm * h * h));
        *mid = (*hi) * h;
        *lo = 0.;
    }
}

// Perlin, perlin.cpp:11-89
struct Perlin {
    int p[512];
    explicit Perlin(unsigned seed) {
        if (seed == 0) seed = std::mt19937::default_seed;
        std::iota(p, p + 256, 0);
        std::shuffle(p, p + 256, std::mt19937(seed));
        for (int i = 0; i < 256; ++i) p[256 + i] = p[i];
    }
    static double fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    static double lerp(double t, double a, double b) { return a + t * (b - a); }
    static double grad(int hash, double x, double y, double z) {
        int h = hash & 15;
        double u = h < 8 ? x : y, v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    double noise(double x, double y, double z) const {
        int X = (int) floor(x) & 255, Y = (int) floor(y) & 255, Z = (int) floor(z) & 255;
        x -= floor(x); y -= floor(y); z -= floor(z);
        double u = fade(x), v = fade(y), w = fade(z);
        int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z, B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
        return lerp(w, lerp(v, lerp(u, grad(p[AA], x, y, z), grad(p[BA], x - 1, y, z)), lerp(u, grad(p[AB], x, y - 1, z), grad(p[BB], x - 1, y - 1, z))),
                    lerp(v, lerp(u, grad(p[AA + 1], x, y, z - 1), grad(p[BA + 1], x - 1, y, z - 1)),
                         lerp(u, grad(p[AB + 1], x, y - 1, z - 1), grad(p[BB + 1], x - 1, y - 1, z - 1))));
    }
    double octave2(double x, double y, int octaves) const {
        double result = 0.0, amp = 1.0;
        for (int i = 0; i < octaves; ++i) { result += noise(x, y, 0.0) * amp; x *= 2.0; y *= 2.0; amp *= 0.5; }
        return result;
    }
};

// Catmull-Rom, spline.cpp:29-57
void spline_at(const std::vector<double> &xs, const std::vector<double> &ys, double t, double *ox, double *oy) {
    int n = (int) xs.size();
    double delta_t = 1.0 / (double) n;
    int p = (int) (t / delta_t);
    auto wrap = [n](int q) { return q < 0 ? n - 1 : (q >= n ? q - n : q); };
    int p0 = wrap(p - 1), p1 = wrap(p), p2 = wrap(p + 1), p3 = wrap(p + 2);
    double lt = (t - delta_t * (double) p) / delta_t;
    double t2 = lt * lt, t3 = t2 * lt;
    double b1 = 0.5 * (-t3 + 2.0 * t2 - lt), b2 = 0.5 * (3.0 * t3 - 5.0 * t2 + 2.0), b3 = 0.5 * (-3.0 * t3 + 4.0 * t2 + lt), b4 = 0.5 * (t3 - t2);
    *ox = ((xs[p0] * b1 + xs[p1] * b2) + xs[p2] * b3) + xs[p3] * b4;
    *oy = ((ys[p0] * b1 + ys[p1] * b2) + ys[p2] * b3) + ys[p3] * b4;
}

// morph.cpp:1467-1488
uint32_t mix_colors(uint32_t c1, uint32_t c2, double w) {
    Col a = unpack(c1), b = unpack(c2);
    return pack(u8(round(w * double(a.r) + (1.0 - w) * double(b.r))), u8(round(w * double(a.g) + (1.0 - w) * double(b.g))),
                u8(round(w * double(a.b) + (1.0 - w) * double(b.b))), u8(round(w * double(a.a) + (1.0 - w) * double(b.a))));
}
double eased(double lag, double slope, double str) {
    const double pi = 3.14159265358;
    double s = (slope + 0.1) / 1.1, l = (1.0 - s) * lag;
    if (str <= l) return 0.0;
    if (str >= (l + s)) return 1.0;
    return ((-cos((str - l) * (pi / s)) + 1.0) / 2.0);
}
// morph.cpp:1501-1515
Pt mix_points(Pt p1, Pt p2, double w) {
    double x1 = 256.0 * p1.x + p1.xf, y1 = 256.0 * p1.y + p1.yf, x2 = 256.0 * p2.x + p2.xf, y2 = 256.0 * p2.y + p2.yf;
    double x = (w * x1 + (1.0 - w) * x2), y = (w * y1 + (1.0 - w) * y2);
    Pt o;
    o.x = (uint16_t) (x / 256.0); o.xf = (uint8_t) (x - (o.x * 256));
    o.y = (uint16_t) (y / 256.0); o.yf = (uint8_t) (y - (o.y * 256));
    o.flags = 0;
    return o;
}

uint64_t dist256(uint64_t a, uint64_t b) {    // atomorph.h:334-339
    Pt p = unpackp(a), q = unpackp(b);
    int64_t dx = (256LL * p.x + p.xf) - (256LL * q.x + q.xf), dy = (256LL * p.y + p.yf) - (256LL * q.y + q.yf);
    return (uint64_t) (dx * dx) + (uint64_t) (dy * dy);
}

} // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ a-N pure functions
uint32_t amo_rgb_to_hsp(uint32_t c) {          // color.cpp:31-44
    Col v = unpack(c);
    double h, s, p;
    rgb2hsp(v.r / 255.0, v.g / 255.0, v.b / 255.0, &h, &s, &p);
    return pack(u8(round(h * 255.0)), u8(round(s * 255.0)), u8(round(p * 255.0)), v.a);
}
uint32_t amo_hsp_to_rgb(uint32_t c) {          // color.cpp:46-59
    Col v = unpack(c);
    double r = 0, g = 0, b = 0;
    hsp2rgb(v.r / 255.0, v.g / 255.0, v.b / 255.0, &r, &g, &b);
    return pack(u8(std::min(round(r * 255.0), 255.0)), u8(std::min(round(g * 255.0), 255.0)), u8(std::min(round(b * 255.0), 255.0)), v.a);
}
double amo_color_distance(uint32_t c1, uint32_t c2) {   // color.h:17-24
    Col a = unpack(c1), b = unpack(c2);
    int rd = a.r - b.r, gd = a.g - b.g, bd = a.b - b.b, ad = a.a - b.a;
    return sqrt((double) (rd * rd + gd * gd + bd * bd + ad * ad)) / 510.0;
}
uint64_t amo_point_distance(uint64_t a, uint64_t b) { return dist256(a, b); }
double amo_octave_noise(unsigned seed, double x, double y, int octaves) { return Perlin(seed).octave2(x, y, octaves); }
void amo_perlin_table(unsigned seed, int32_t *out512) { Perlin pn(seed); for (int i = 0; i < 512; ++i) out512[i] = pn.p[i]; }
void amo_spline_point(uint64_t n, const double *xs, const double *ys, double t, double *out2) {
    std::vector<double> vx(xs, xs + n), vy(ys, ys + n);
    spline_at(vx, vy, t, out2, out2 + 1);
}
uint32_t amo_interpolate_color(uint32_t c1, uint32_t c2, double lag, double slope, double str, int fading_eased) {
    return mix_colors(c1, c2, fading_eased ? eased(lag, slope, str) : str);
}
uint64_t amo_interpolate_point(uint64_t p1, uint64_t p2, double w) {
    Pt o = mix_points(unpackp(p1), unpackp(p2), w);
    return packp(o.x, o.y, o.xf, o.yf, 0);
}

// ------------------------------------------------------------------------------------------ a-M cost + serial matcher
// thread.cpp:1109-1125, summed over chains of width > 1
double amo_cost(uint64_t nchains, const uint64_t *widths, uint64_t height, const uint64_t *words) {
    double e = 0.0;
    const uint64_t *t = words;
    for (uint64_t c = 0; c < nchains; ++c) {
        uint64_t w = widths[c];
        if (w > 1 && height > 0)
            for (uint64_t x = 0; x < w; ++x)
                for (uint64_t j = 0; j < height; ++j) e += (double) dist256(t[j * w + x], t[((j + 1) % height) * w + x]);
        t += w * height;
    }
    return e;
}

// morph_asynch (thread.cpp:990-1041) for ONE chain, threads = 0: `steps` calls, each with its own
// mt19937 seeded from uniform_int_distribution<unsigned>(e1), e1 = default_random_engine (thread.cpp:1046).
// `e1_state` carries the engine across calls (in/out), so a run can be replayed exactly.
void amo_morph_steps(uint64_t *words, uint64_t width, uint64_t height, uint64_t steps, uint64_t cycle_length, uint64_t *e1_state,
                     double *gain_out) {
    std::default_random_engine e1;
    e1.seed((std::default_random_engine::result_type) *e1_state);         // minstd_rand0: the state is one integer in [1, 2^31-2]
    double gain = 0.0;
    if (width > 1) {
        std::uniform_int_distribution<unsigned> seed_dist(0, std::numeric_limits<unsigned>::max());
        for (uint64_t s = 0; s < steps; ++s) {
            std::mt19937 gen(seed_dist(e1));
            std::uniform_int_distribution<size_t> dx(0, width - 1), dy(0, height - 1);
            size_t y = dy(gen), yn = (y + 1) % height, yp = (y > 0 ? y - 1 : height - 1);
            for (uint64_t i = 0; i < cycle_length; ++i) {
                size_t x1 = dx(gen), x2;
                do { x2 = dx(gen); } while (x1 == x2);
                uint64_t a = words[y * width + x1], b = words[y * width + x2];
                uint64_t ap = words[yp * width + x1], bp = words[yp * width + x2], an = words[yn * width + x1], bn = words[yn * width + x2];
                double c1 = ((double) dist256(ap, a) + (double) dist256(a, an)) + ((double) dist256(bp, b) + (double) dist256(b, bn));
                double c2 = ((double) dist256(bp, a) + (double) dist256(a, bn)) + ((double) dist256(ap, b) + (double) dist256(b, an));
                if (c1 >= c2) { words[y * width + x1] = b; words[y * width + x2] = a; gain += c1 - c2; }
            }
        }
    }
    {   // export the state: x_{n+1} = 16807 x_n mod (2^31-1)  =>  x_n = x_{n+1} * 16807^-1
        std::default_random_engine probe = e1;
        const uint64_t mod = 2147483647ull, inv = 1407677000ull;
        *e1_state = ((uint64_t) probe() * inv) % mod;
    }
    if (gain_out) *gain_out = gain;
}

// ------------------------------------------------------------------------------------------ a-B blobs
// 4-connected components of the presence mask: the partition thread::blobify_frame (thread.cpp:225-412)
// reaches with the default parameters (threshold 1.0, max SIZE_MAX, min 1, number 1) -- SURVEY.md M2.
// labels_out = smallest canvas index of the component (canonical label), -1 where absent;
// stats_out[7*k..] for the k-th component in ascending canonical label: size, x, y, r, g, b, a means.
uint64_t amo_blobify(uint32_t w, uint32_t h, const uint8_t *present, const uint32_t *stored, int64_t *labels_out, double *stats_out, uint64_t stats_cap) {
    size_t n = (size_t) w * h;
    for (size_t i = 0; i < n; ++i) labels_out[i] = -1;
    std::vector<size_t> stack;
    uint64_t count = 0;
    for (size_t s = 0; s < n; ++s) {
        if (!present[s] || labels_out[s] >= 0) continue;
        double sx = 0, sy = 0, sr = 0, sg = 0, sb = 0, sa = 0, cnt = 0;
        stack.push_back(s);
        labels_out[s] = (int64_t) s;
        while (!stack.empty()) {
            size_t i = stack.back(); stack.pop_back();
            uint32_t x = (uint32_t) (i % w), y = (uint32_t) (i / w);
            Col c = unpack(stored[i]);
            sx += x; sy += y; sr += c.r; sg += c.g; sb += c.b; sa += c.a; cnt += 1;
            const long nb[4] = {x + 1 < w ? 1L : 0L, x > 0 ? -1L : 0L, y + 1 < h ? (long) w : 0L, y > 0 ? -(long) w : 0L};
            for (int k = 0; k < 4; ++k) {
                if (!nb[k]) continue;
                size_t j = i + nb[k];
                if (present[j] && labels_out[j] < 0) { labels_out[j] = (int64_t) s; stack.push_back(j); }
            }
        }
        if (count < stats_cap) {
            double *o = stats_out + 7 * count;
            o[0] = cnt; o[1] = sx / cnt; o[2] = sy / cnt; o[3] = sr / cnt / 255.0; o[4] = sg / cnt / 255.0; o[5] = sb / cnt / 255.0; o[6] = sa / cnt / 255.0;
        }
        ++count;
    }
    return count;
}

// thread.cpp:1151-1174; weights as set_blob_weights (1141-1149) produced them; sizes, centroids (double), mean colours (double 0..1)
double amo_blob_distance(double sz1, const double *s1, double sz2, const double *s2, double w_xy, double w_rgba, double w_size, uint32_t bbox_d) {
    double szs = sz1 + sz2, pix = 0.0, col = 0.0, siz = 0.0;
    if (szs > 0) siz = fabs(sz1 - sz2) / szs;
    if (sz1 > 0 && sz2 > 0) {
        uint16_t x1 = (uint16_t) s1[0], y1 = (uint16_t) s1[1], x2 = (uint16_t) s2[0], y2 = (uint16_t) s2[1];
        int32_t xd = x1 - x2, yd = y1 - y2;
        uint32_t pd = xd * xd + yd * yd;
        pix = sqrt(double(pd) / bbox_d);
        col = amo_color_distance(color_from_doubles(s1[2], s1[3], s1[4], s1[5]), color_from_doubles(s2[2], s2[3], s2[4], s2[5]));
    }
    return (w_xy * pix + w_rgba * col + w_size * siz);
}

// ------------------------------------------------------------------------------------------ a-R renderer
// morph::get_pixels(t) for the draw_atoms path (morph.cpp:452-678, 1302-1421, 1431-1465), serial, in the reference's
// own order of double operations.  All inputs are plain arrays:
//   fetch[f]      canvas images (cw*ch) of get_pixel colours (RGB after the store/fetch round trip), 0 where absent
//   has[f]        canvas presence masks
//   blobs         per frame: nb[f] blobs in vector order with group[] and stats[6] (for AVERAGE)
//   chains        nchains chains: key (group), width, words (column-major, height = nframes)
struct amo_scene {
    uint32_t width, height, cw, ch, nframes, nchains;
    uint32_t bbox[4];
    const uint64_t *frame_keys;
    const uint32_t *const *fetch;
    const uint8_t *const *has;
    const uint32_t *nblobs;              // [nframes]
    const uint64_t *const *blob_group;   // [nframes][nblobs]
    const double *const *blob_stats;     // [nframes][nblobs*6]
    const uint64_t *chain_key, *chain_width;
    const uint64_t *const *chain_words;
    uint32_t motion, fading, density, feather, show_blobs, keep_background, blob_delimiter, seed;
};

static uint32_t scene_pixel(const amo_scene *S, uint32_t f, int x, int y) {   // morph::get_pixel, morph.cpp:378-392
    if (x < 0 || y < 0 || (uint32_t) x >= S->cw || (uint32_t) y >= S->ch) return 0;
    size_t i = (size_t) y * S->cw + x;
    return S->has[f][i] ? S->fetch[f][i] : 0;
}

static bool locate(const amo_scene *S, double t, double *time, uint32_t *f, double *tl) {
    double integ, tm = modf(t, &integ);
    if (tm < 0.0) tm += 1.0;
    size_t idx = (size_t) (tm * S->nframes);
    if (S->nframes == 0 || idx >= S->nframes) return false;
    double dt = 1.0 / double(S->nframes);
    *tl = std::max(0.0, (tm - (S->frame_keys[idx] * dt)) / dt);
    *time = tm; *f = (uint32_t) idx;
    return true;
}

static uint32_t background_at(const amo_scene *S, const Perlin &lag_map, const Perlin &slope_map, int x, int y, double t) {
    double time, tl; uint32_t f;
    if (!locate(S, t, &time, &f, &tl)) return 0;
    uint32_t fn = (f + 1) % S->nframes;
    uint32_t c1 = scene_pixel(S, f, x, y), c2 = scene_pixel(S, fn, x, y);
    if (S->fading == 6) {
        double fq = 8.0, bw = double((int) S->bbox[2] - (int) S->bbox[0]) + 1.0, bh = double((int) S->bbox[3] - (int) S->bbox[1]) + 1.0;
        double px = ((x - (int) S->bbox[0]) / double(bw)) * fq, py = ((y - (int) S->bbox[1]) / double(bh)) * fq;
        double lag = lag_map.octave2(px, py, 8) * 0.5 + 0.5, slope = slope_map.octave2(px, py, 8) * 0.5 + 0.5;
        return mix_colors(c1, c2, eased(lag, slope, 1.0 - tl));
    }
    if (S->fading == 5) return mix_colors(c1, c2, eased(0.5, 0.5, 1.0 - tl));
    return mix_colors(c1, c2, 1.0 - tl);
}

// one blob's pixels at time t, in the reference's emission order (morph.cpp:452-678); returns false for "nullptr"
static bool blob_pixels(const amo_scene *S, const Perlin &lag_map, const Perlin &slope_map, uint32_t blob_index, double t_in,
                        std::vector<std::pair<size_t, uint32_t>> *out, uint64_t *group_out) {
    double time, tl; uint32_t f;
    if (!locate(S, t_in, &time, &f, &tl)) return false;
    if (blob_index >= S->nblobs[f]) return false;
    uint64_t group = S->blob_group[f][blob_index];
    if (group_out) *group_out = group;
    int64_t c = -1;
    for (uint32_t k = 0; k < S->nchains; ++k) if (S->chain_key[k] == group) { c = k; break; }
    if (c < 0) return false;
    uint64_t w = S->chain_width[c];
    uint32_t h = S->nframes, y = f, yn = (f + 1) % h;
    const uint64_t *words = S->chain_words[c];
    std::map<size_t, std::vector<uint32_t>> colors;
    std::map<size_t, std::vector<double>> weights;
    int W = (int) S->width, H = (int) S->height, bx1 = S->bbox[0], by1 = S->bbox[1], bx2 = S->bbox[2], by2 = S->bbox[3];
    std::vector<double> sx(h), sy(h);
    for (uint64_t x = 0; x < w; ++x) {
        Pt pt1 = unpackp(words[(size_t) y * w + x]), pt2 = unpackp(words[(size_t) yn * w + x]);
        bool has1 = pt1.flags & 1, has2 = pt2.flags & 1;
        if (!has1 && !has2) continue;
        uint32_t c1, c2;
        if (has1 && !has2) { c1 = scene_pixel(S, y, pt1.x, pt1.y); c2 = c1 & 0x00ffffffu; }
        else if (has2 && !has1) { c2 = scene_pixel(S, yn, pt2.x, pt2.y); c1 = c2 & 0x00ffffffu; }
        else { c1 = scene_pixel(S, y, pt1.x, pt1.y); c2 = scene_pixel(S, yn, pt2.x, pt2.y); }
        Pt pt = pt1;
        if (S->motion == 3) pt = mix_points(pt1, pt2, 1.0 - tl);
        else if (S->motion == 4) {
            for (uint32_t j = 0; j < h; ++j) { Pt q = unpackp(words[(size_t) j * w + x]); sx[j] = q.x + q.xf / 256.0; sy[j] = q.y + q.yf / 256.0; }
            double vx, vy, fract, integ;
            spline_at(sx, sy, time, &vx, &vy);
            fract = modf(vx, &integ); pt.x = (uint16_t) integ; pt.xf = (uint8_t) round(fract * 255);
            fract = modf(vy, &integ); pt.y = (uint16_t) integ; pt.yf = (uint8_t) round(fract * 255);
        }
        uint32_t col;
        if (S->fading == 6) {
            double fq = 8.0, bw = bx2 - bx1 + 1.0, bh = by2 - by1 + 1.0;
            double px = (((pt1.x - bx1) * 256 + pt1.xf) / double(bw * 256)) * fq, py = (((pt1.y - by1) * 256 + pt1.yf) / double(bh * 256)) * fq;
            double lag = lag_map.octave2(px, py, 8) * 0.5 + 0.5, slope = slope_map.octave2(px, py, 8) * 0.5 + 0.5;
            col = mix_colors(c1, c2, eased(lag, slope, 1.0 - tl));
        } else if (S->fading == 5) col = mix_colors(c1, c2, eased(0.5, 0.5, 1.0 - tl));
        else col = mix_colors(c1, c2, 1.0 - tl);
        int X = pt.x, Y = pt.y;
        if (X >= W || Y >= H) { if (X > bx2 || X < bx1 || Y > by2 || Y < by1) continue; }
        double total = 255.0 * 255.0;
        double w11 = ((255 - pt.xf) * (255 - pt.yf)) / total, w21 = (pt.xf * (255 - pt.yf)) / total;
        double w12 = ((255 - pt.xf) * pt.yf) / total, w22 = (pt.xf * pt.yf) / total;
        auto put = [&](int px, int py, double wt) { size_t pos = (size_t) py * 65536 + px; colors[pos].push_back(col); weights[pos].push_back(wt); };
        if (w11 > 0.0) put(X, Y, w11);
        if ((X < bx2 || X + 1 < W) && w21 > 0.0) put(X + 1, Y, w21);
        if ((Y < by2 || Y + 1 < H) && w12 > 0.0) put(X, Y + 1, w12);
        if (w22 > 0.0 && ((Y < by2 && X < bx2) || (Y + 1 < H && X + 1 < W))) put(X + 1, Y + 1, w22);
    }
    std::map<size_t, uint32_t> blob;
    for (auto &kv : colors) {
        const std::vector<uint32_t> &cs = kv.second;
        const std::vector<double> &ws = weights[kv.first];
        double r = 0, g = 0, b = 0, a = 0, wsum = 0;
        for (size_t k = 0; k < cs.size(); ++k) {
            Col cc = unpack(cs[k]);
            wsum += ws[k]; r += cc.r * ws[k]; g += cc.g * ws[k]; b += cc.b * ws[k]; a += cc.a * ws[k];
        }
        double wd = S->density > 0 ? double(cs.size()) / double(S->density) : 0.0;
        if (wd > 1.0) wd = 1.0;
        uint32_t px = pack(u8(round(r / wsum)), u8(round(g / wsum)), u8(round(b / wsum)), u8(round(wd * (a / wsum))));
        if (S->feather > 0) blob[kv.first] = px; else out->push_back(std::make_pair(kv.first, px));
    }
    if (S->feather > 0) {   // morph.cpp:625-674
        std::vector<std::set<size_t>> layers;
        std::map<size_t, uint32_t> peeled = blob;
        while (!peeled.empty() && layers.size() < S->feather) {
            std::set<size_t> border;
            for (auto &kv : peeled) {
                size_t pos = kv.first, x = pos % 65536, yy = pos / 65536;
                if (x == 0 || x == 65535 || yy == 0 || yy == 65535) { border.insert(pos); continue; }
                if (!peeled.count(pos + 1) || !peeled.count(pos - 1) || !peeled.count(pos + 65536) || !peeled.count(pos - 65536)) border.insert(pos);
            }
            for (size_t pos : border) peeled.erase(pos);
            layers.push_back(border);
        }
        for (size_t l = 0; l < layers.size(); ++l)
            for (size_t pos : layers[l]) {
                uint32_t px = blob[pos];
                int a = u8(round(double(px >> 24) * (double(l + 1) / double(S->feather + 1))));
                out->push_back(std::make_pair(pos, (px & 0x00ffffffu) | ((uint32_t) a << 24)));
            }
        for (auto &kv : peeled) out->push_back(kv);
    }
    return true;
}

// whole frame (morph.cpp:1302-1421); out = width*height packed RGBA
void amo_render(const amo_scene *S, double t, uint32_t *out) {
    Perlin lag_map(S->seed), slope_map(S->seed + 1);
    size_t W = S->width, H = S->height;
    for (size_t y = 0; y < H; ++y)
        for (size_t x = 0; x < W; ++x) out[y * W + x] = S->keep_background ? background_at(S, lag_map, slope_map, (int) x, (int) y, t) : 0u;
    double time, tl; uint32_t f;
    if (!locate(S, t, &time, &f, &tl)) return;
    std::map<size_t, std::vector<uint32_t>> colors;
    std::vector<std::pair<size_t, uint32_t>> pixels;
    for (uint32_t b = 0;; ++b) {
        pixels.clear();
        uint64_t group = 0;
        if (!blob_pixels(S, lag_map, slope_map, b, t, &pixels, &group)) break;
        uint32_t avg = color_from_doubles(S->blob_stats[f][6 * b + 2], S->blob_stats[f][6 * b + 3], S->blob_stats[f][6 * b + 4], S->blob_stats[f][6 * b + 5]);
        if (S->blob_delimiter == 1) avg = amo_hsp_to_rgb(avg);
        for (size_t k = pixels.size(); k-- > 0;) {
            size_t px = pixels[k].first % 65536, py = pixels[k].first / 65536, pos = py * W + px;
            if (pos >= W * H) continue;
            uint32_t c = pixels[k].second;
            if (S->show_blobs == 2) {
                std::mt19937 gen((unsigned) group);
                std::uniform_int_distribution<unsigned char> d(0, 255);
                unsigned rr = d(gen), gg = d(gen), bb = d(gen);
                c = pack(rr, gg, bb, 255);
            } else if (S->show_blobs == 1) c = avg;
            if ((c >> 24) == 0) continue;
            colors[pos].push_back(c);
        }
    }
    for (auto &kv : colors) {
        double r = 0, g = 0, b = 0, a = 0;
        for (size_t k = 0; k < kv.second.size(); ++k) {
            Col c = unpack(kv.second[k]);
            double sr = c.r / 255.0, sg = c.g / 255.0, sb = c.b / 255.0, sa = c.a / 255.0;
            if (k == 0) { r = sr; g = sg; b = sb; a = sa; }
            else { r = sa * sr + (1.0 - sa) * r; g = sa * sg + (1.0 - sa) * g; b = sa * sb + (1.0 - sa) * b; a = a + (1.0 - a) * (c.a / 255.0); }
        }
        if (S->keep_background) {
            Col bg = unpack(out[kv.first]);
            double bgr = bg.r / 255.0, bgg = bg.g / 255.0, bgb = bg.b / 255.0, bga = bg.a / 255.0;
            r = a * r + (1.0 - a) * bgr; g = a * g + (1.0 - a) * bgg; b = a * b + (1.0 - a) * bgb; a = bga + (1.0 - bga) * a;
        }
        out[kv.first] = color_from_doubles(r, g, b, a);
    }
}

uint32_t amo_background(const amo_scene *S, int x, int y, double t) {
    Perlin lag_map(S->seed), slope_map(S->seed + 1);
    return background_at(S, lag_map, slope_map, x, y, t);
}

// ------------------------------------------------------------------------------------------ a-F fluid step
// FluidModel::step (fluidmodel.cpp:165-580), serial, on plain arrays.  Particle record = 24 doubles as in
// include/amx.h (AMX_FP_STRIDE); nodes are 13 doubles m d gx gy u v ax ay r g b a weight, zero-initialised.
// Wall clamping (fluidmodel.cpp:553-566) uses libc rand() exactly as the reference does.
void amo_fluid_step(uint32_t gsx, uint32_t gsy, uint32_t n, double *rec, double *nodes, uint64_t steps_left, double freedom_radius) {
    enum { X, Y, U, V, GX, GY, FREE, ACT, MAT, RI, GI, BI, AI, R, G, B, A, STR };
    enum { M, D, NGX, NGY, NU, NV, AX, AY, NR, NG, NB, NA, NW };
    size_t ng = (size_t) gsx * gsy;
    memset(nodes, 0, ng * 13 * sizeof(double));
    struct W9 { unsigned cx, cy; double px[3], py[3], gx[3], gy[3]; };
    std::vector<W9> wts(n);
    auto node = [&](unsigned i, unsigned j) -> double * { return nodes + ((size_t) j * gsx + i) * 13; };
    for (uint32_t i = 0; i < n; ++i) {
        double *p = rec + (size_t) i * 24;
        if (p[ACT] == 0.0) continue;
        W9 &w = wts[i];
        w.cx = (unsigned) (int) (p[X] - 0.5); w.cy = (unsigned) (int) (p[Y] - 0.5);
        double x = w.cx - p[X];
        w.px[0] = (0.5 * x * x + 1.5 * x + 1.125); w.gx[0] = (x + 1.5); x += 1.0;
        w.px[1] = (-x * x + 0.75); w.gx[1] = (-2.0 * x); x += 1.0;
        w.px[2] = (0.5 * x * x - 1.5 * x + 1.125); w.gx[2] = (x - 1.5);
        double y = w.cy - p[Y];
        w.py[0] = (0.5 * y * y + 1.5 * y + 1.125); w.gy[0] = (y + 1.5); y += 1.0;
        w.py[1] = (-y * y + 0.75); w.gy[1] = (-2.0 * y); y += 1.0;
        w.py[2] = (0.5 * y * y - 1.5 * y + 1.125); w.gy[2] = (y - 1.5);
        for (unsigned a = 0; a < 3; ++a)
            for (unsigned b = 0; b < 3; ++b) {
                double *nd = node(w.cx + a, w.cy + b);
                double phi = w.px[a] * w.py[b];
                nd[M] += phi * 1.0; nd[D] += phi; nd[NGX] += w.gx[a] * w.py[b]; nd[NGY] += w.px[a] * w.gy[b];
                if (p[MAT] != 0.0 && p[STR] > 0.0) {
                    double nw = p[STR], ow = nd[NW], sw = nw + ow;
                    nd[NR] = (ow * nd[NR] + nw * p[R]) / sw; nd[NG] = (ow * nd[NG] + nw * p[G]) / sw;
                    nd[NB] = (ow * nd[NB] + nw * p[B]) / sw; nd[NA] = (ow * nd[NA] + nw * p[A]) / sw;
                    nd[NW] = sw;
                }
            }
    }
    for (uint32_t i = 0; i < n; ++i) {
        double *p = rec + (size_t) i * 24;
        if (p[ACT] == 0.0) continue;
        W9 &w = wts[i];
        unsigned cx = (unsigned) (int) p[X], cy = (unsigned) (int) p[Y];
        double *n01 = node(cx, cy), *n02 = node(cx, cy + 1), *n11 = node(cx + 1, cy), *n12 = node(cx + 1, cy + 1);
        double pdx = n11[D] - n01[D], pdy = n02[D] - n01[D];
        double C20 = 3.0 * pdx - n11[NGX] - 2.0 * n01[NGX], C02 = 3.0 * pdy - n02[NGY] - 2.0 * n01[NGY];
        double C30 = -2.0 * pdx + n11[NGX] + n01[NGX], C03 = -2.0 * pdy + n02[NGY] + n01[NGY];
        double csum1 = n01[D] + n01[NGY] + C02 + C03, csum2 = n01[D] + n01[NGX] + C20 + C30;
        double C21 = 3.0 * n12[D] - 2.0 * n02[NGX] - n12[NGX] - 3.0 * csum1 - C20;
        double C31 = -2.0 * n12[D] + n02[NGX] + n12[NGX] + 2.0 * csum1 - C30;
        double C12 = 3.0 * n12[D] - 2.0 * n11[NGY] - n12[NGY] - 3.0 * csum2 - C02;
        double C13 = -2.0 * n12[D] + n11[NGY] + n12[NGY] + 2.0 * csum2 - C03;
        double C11 = n02[NGX] - C13 - C12 - n01[NGX];
        double u = p[X] - cx, u2 = u * u, u3 = u * u2, v = p[Y] - cy, v2 = v * v, v3 = v * v2;
        double density = n01[D] + n01[NGX] * u + n01[NGY] * v + C20 * u2 + C02 * v2 + C30 * u3 + C03 * v3 + C21 * u2 * v + C31 * u3 * v +
                         C12 * u * v2 + C13 * u * v3 + C11 * u * v;
        double pressure = density - 1.0;
        if (pressure > 2.0) pressure = 2.0;
        double fx = 0.0, fy = 0.0;
        if (p[X] < 4.0) fx += 1.0 * (4.0 - p[X]); else if (p[X] > gsx - 5) fx += 1.0 * (gsx - 5 - p[X]);
        if (p[Y] < 4.0) fy += 1.0 * (4.0 - p[Y]); else if (p[Y] > gsy - 5) fy += 1.0 * (gsy - 5 - p[Y]);
        for (unsigned a = 0; a < 3; ++a)
            for (unsigned b = 0; b < 3; ++b) {
                double *nd = node(w.cx + a, w.cy + b);
                double phi = w.px[a] * w.py[b];
                nd[AX] += -((w.gx[a] * w.py[b]) * pressure) + fx * phi;
                nd[AY] += -((w.px[a] * w.gy[b]) * pressure) + fy * phi;
            }
    }
    for (size_t k = 0; k < ng; ++k) { double *nd = nodes + k * 13; if (nd[M] > 0.0) { nd[AX] /= nd[M]; nd[AY] /= nd[M]; } }
    auto pull = [](double x1, double y1, double x2, double y2, double a, double *ox, double *oy) {
        double Ad = fabs(y1 - y2), Bd = fabs(x1 - x2), Cd = sqrt(Ad * Ad + Bd * Bd);
        if (a >= Cd) a = Cd;
        *ox = 0.0; *oy = 0.0;
        if (Bd <= 0.0) { if (y2 <= y1) *oy -= a; else *oy += a; }
        else if (Cd > 0.0) {
            double dx = (a * Bd) / Cd, dy = (Ad * dx) / Bd;
            if (x1 <= x2) *ox += dx; else *ox -= dx;
            if (y1 <= y2) *oy += dy; else *oy -= dy;
        }
    };
    for (uint32_t i = 0; i < n; ++i) {
        double *p = rec + (size_t) i * 24;
        if (p[ACT] == 0.0) continue;
        W9 &w = wts[i];
        for (unsigned a = 0; a < 3; ++a)
            for (unsigned b = 0; b < 3; ++b) {
                double *nd = node(w.cx + a, w.cy + b);
                double phi = w.px[a] * w.py[b], cax, cay;
                pull(p[X], p[Y], p[GX], p[GY], 0.03, &cax, &cay);
                p[U] += phi * (nd[AX] + cax);
                p[V] += phi * (nd[AY] + cay);
            }
        double mu = 1.0 * p[U], mv = 1.0 * p[V];
        if (p[MAT] == 0.0) { mu *= 0.0; mv *= 0.0; }
        for (unsigned a = 0; a < 3; ++a)
            for (unsigned b = 0; b < 3; ++b) {
                double *nd = node(w.cx + a, w.cy + b);
                double phi = w.px[a] * w.py[b];
                nd[NU] += phi * mu; nd[NV] += phi * mv;
            }
    }
    for (size_t k = 0; k < ng; ++k) { double *nd = nodes + k * 13; if (nd[M] > 0.0) { nd[NU] /= nd[M]; nd[NV] /= nd[M]; } }
    for (uint32_t i = 0; i < n; ++i) {
        double *p = rec + (size_t) i * 24;
        if (p[ACT] == 0.0) continue;
        W9 &w = wts[i];
        double gu = 0.0, gv = 0.0, nR = 0, nG = 0, nB = 0, nA = 0, weight = 0.0;
        for (unsigned a = 0; a < 3; ++a)
            for (unsigned b = 0; b < 3; ++b) {
                double *nd = node(w.cx + a, w.cy + b);
                double phi = w.px[a] * w.py[b];
                gu += phi * nd[NU]; gv += phi * nd[NV];
                if (nd[NW] > 0.0) { weight += nd[NW]; nR += nd[NR] * nd[NW]; nG += nd[NG] * nd[NW]; nB += nd[NB] * nd[NW]; nA += nd[NA] * nd[NW]; }
            }
        if (weight > 0.0) {
            nR /= weight; nG /= weight; nB /= weight; nA /= weight;
            double wr = fabs(nR - p[RI]), wg = fabs(nG - p[GI]), wb = fabs(nB - p[BI]), wa = fabs(nA - p[AI]);
            if (p[MAT] == 0.0) { p[R] = nR; p[G] = nG; p[B] = nB; p[A] = nA; }
            else { p[R] = (1.0 - wr) * p[R] + wr * nR; p[G] = (1.0 - wg) * p[G] + wg * nG; p[B] = (1.0 - wb) * p[B] + wb * nB; p[A] = (1.0 - wa) * p[A] + wa * nA; }
        }
        p[X] += gu; p[Y] += gv;
        {
            double Ad = fabs(p[Y] - p[GY]), Bd = fabs(p[X] - p[GX]), Cd = sqrt(Ad * Ad + Bd * Bd);
            double r = freedom_radius * p[FREE];
            if (Cd > r) {
                double mx, my;
                pull(p[X], p[Y], p[GX], p[GY], Cd - r, &mx, &my);
                double ww = 1.0 / (steps_left + 1);
                p[X] += mx * ww; p[Y] += my * ww;
            }
        }
        p[U] += gu - p[U]; p[V] += gv - p[V];
        if (p[X] < 1.0) { p[X] = 1.0 + static_cast<double>(rand()) / RAND_MAX * 0.01; p[U] = 0.0; }
        else if (p[X] > gsx - 2) { p[X] = gsx - 2 - static_cast<double>(rand()) / RAND_MAX * 0.01; p[U] = 0.0; }
        if (p[Y] < 1.0) { p[Y] = 1.0 + static_cast<double>(rand()) / RAND_MAX * 0.01; p[V] = 0.0; }
        else if (p[Y] > gsy - 2) { p[Y] = gsy - 2 - static_cast<double>(rand()) / RAND_MAX * 0.01; p[V] = 0.0; }
        p[22] = w.cx; p[23] = w.cy;
    }
}

const char *amo_version() { return "am_oracle 1 (restates 1Hyena/atomorph morph.cpp thread.cpp fluidmodel.cpp color.cpp perlin.cpp spline.cpp)"; }

} // extern "C"
