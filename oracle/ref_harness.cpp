/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A flat C API around the UNMODIFIED reference (1Hyena/atomorph), compiled from
 * the sources where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libamref.so.  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg can (a) run the reference
 * deterministically stage by stage, (b) dump its private state (frames, blobs,
 * chain tables, fluid particles) and (c) import tables so the reference's own
 * renderer / matcher can be run on given inputs.  Nothing under
 * atomorph_b200/ links, loads or calls this file.
 *
 * Recipe follows SURVEY.md appendix A:  `#define private public` before the
 * reference header (layout is unaffected under the Itanium ABI), threads=0 and
 * iterate(n) only, lost-resume guard (reference thread.cpp:194-196 vs thread.h:48).
 */
#include <thread>
#include <mutex>
#include <atomic>
#include <chrono>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <map>
#include <set>
#include <random>
#include <algorithm>
#include <future>
#include <limits>
#include <iostream>
#include <sstream>
#include <string>

#define private public
#include "atomorph.h"
#undef private

#include "../include/amx_params.h"

namespace {

struct Ref {
    am::morph m;
    bool skipped = false;
};

inline Ref *R(void *h) { return reinterpret_cast<Ref *>(h); }

size_t to_size(double v) {
    if (v >= 1.8e19) return SIZE_MAX;
    if (v < 0) return 0;
    return (size_t) v;
}

// Lost-resume-safe iterate: reference thread.cpp:194-196 may swallow a resume().
void run_steps(am::morph &m, size_t n) {
    if (n == 0) return;
    do {
        m.iterate(n);
        while (m.is_busy()) std::this_thread::sleep_for(std::chrono::microseconds(50));
    } while (m.worker.iterations != 0);
}

am::color unpack(uint32_t rgba) {
    am::color c;
    c.r = rgba & 255; c.g = (rgba >> 8) & 255; c.b = (rgba >> 16) & 255; c.a = (rgba >> 24) & 255;
    return c;
}
uint32_t pack(am::color c) {
    return uint32_t(c.r) | (uint32_t(c.g) << 8) | (uint32_t(c.b) << 16) | (uint32_t(c.a) << 24);
}

} // namespace

extern "C" {

void *amref_create() { return new Ref(); }
void amref_destroy(void *h) { delete R(h); }

void amref_set(void *h, int id, double v) {
    am::morph &m = R(h)->m;
    switch (id) {
        case AMX_P_BLOB_DELIMITER:   m.set_blob_delimiter((unsigned char) v); break;
        case AMX_P_BLOB_THRESHOLD:   m.set_blob_threshold(v); break;
        case AMX_P_BLOB_MAX_SIZE:    m.set_blob_max_size(to_size(v)); break;
        case AMX_P_BLOB_MIN_SIZE:    m.set_blob_min_size(to_size(v)); break;
        case AMX_P_BLOB_BOX_GRIP:    m.set_blob_box_grip((uint16_t) v); break;
        case AMX_P_BLOB_BOX_SAMPLES: m.set_blob_box_samples(to_size(v)); break;
        case AMX_P_BLOB_NUMBER:      m.set_blob_number(to_size(v)); break;
        case AMX_P_BLOB_RGBA_WEIGHT: m.set_blob_rgba_weight((unsigned char) v); break;
        case AMX_P_BLOB_SIZE_WEIGHT: m.set_blob_size_weight((unsigned char) v); break;
        case AMX_P_BLOB_XY_WEIGHT:   m.set_blob_xy_weight((unsigned char) v); break;
        case AMX_P_DEGENERATION:     m.set_degeneration(to_size(v)); break;
        case AMX_P_DENSITY:          m.set_density((uint16_t) v); break;
        case AMX_P_MOTION:           m.set_motion((unsigned char) v); break;
        case AMX_P_FADING:           m.set_fading((unsigned char) v); break;
        case AMX_P_THREADS:          m.set_threads(to_size(v)); break;
        case AMX_P_CYCLE_LENGTH:     m.set_cycle_length(to_size(v)); break;
        case AMX_P_FEATHER:          m.set_feather(to_size(v)); break;
        case AMX_P_KEEP_BACKGROUND:  m.set_keep_background(v != 0.0); break;
        case AMX_P_FINITE:           m.set_finite(v != 0.0); break;
        case AMX_P_SHOW_BLOBS:       m.set_show_blobs((unsigned) v); break;
        case AMX_P_FLUID:            m.set_fluid((unsigned) v); break;
        case AMX_P_SEED:             m.set_seed((unsigned) v); break;
        default: break;
    }
}

void amref_add_pixels(void *h, uint64_t frame, uint64_t n, const uint16_t *x, const uint16_t *y, const uint32_t *rgba) {
    am::morph &m = R(h)->m;
    for (uint64_t i = 0; i < n; ++i) m.add_pixel(frame, am::create_pixel(x[i], y[i], unpack(rgba[i])));
}
int amref_add_frame(void *h, uint64_t frame) { return R(h)->m.add_frame(frame) ? 1 : 0; }
void amref_set_resolution(void *h, uint16_t w, uint16_t hh) { R(h)->m.set_resolution(w, hh); }

void amref_sync(void *h) { R(h)->m.suspend(); R(h)->m.synchronize(); }
unsigned amref_state(void *h) { return R(h)->m.get_state(); }
void amref_iterate(void *h, uint64_t n) { run_steps(R(h)->m, n); }
void amref_next_state(void *h) { R(h)->m.next_state(); }
double amref_energy(void *h) { return R(h)->m.get_energy(); }
double amref_best_blob_energy(void *h) { return R(h)->m.worker.best_blob_map_e; }
// worker internals for diagnostics: blob_map_e, best_e, best_blob_map_e, bbox_d, blob_map_w, blob_map_h, counter, weights(3)
void amref_worker_values(void *h, double *out10) {
    am::thread &w = R(h)->m.worker;
    out10[0] = w.blob_map_e; out10[1] = w.best_e; out10[2] = w.best_blob_map_e; out10[3] = w.bbox_d;
    out10[4] = (double) w.blob_map_w; out10[5] = (double) w.blob_map_h; out10[6] = (double) w.counter;
    out10[7] = w.blob_rgba_weight; out10[8] = w.blob_size_weight; out10[9] = w.blob_xy_weight;
}

// state of the worker's std::default_random_engine (minstd_rand0: one integer), for exact replays
uint64_t amref_e1_state(void *h) {
    // x_{n+1} = 16807 x_n mod (2^31-1): read the next output of a COPY and step it back (no iostreams: this
    // library carries a static libstdc++ whose stream locale is not initialised inside a dlopen'ed object)
    std::default_random_engine probe = R(h)->m.worker.e1;
    const uint64_t mod = 2147483647ull, inv = 1407677000ull;     // 16807 * inv = 1 (mod 2^31-1)
    return ((uint64_t) probe() * inv) % mod;
}

// Drive the pre-stages deterministically until the state reaches `target`.
// In STATE_BLOB_MATCHING the reference never leaves by itself (unless best_e==0,
// thread.cpp:734), so after `match_steps` matching steps next_state() is issued ONCE.
// Returns the state reached.
unsigned amref_run_until(void *h, unsigned target, uint64_t chunk, uint64_t match_steps) {
    Ref *r = R(h);
    am::morph &m = r->m;
    uint64_t matched = 0;
    // The first synchronize() after ingest resets the worker AFTER pushing the bounding box
    // (morph.cpp:119-136: set_bbox, then worker.clear()), so bbox_d is 0 until the next one.  The demo
    // synchronizes every 100 ms; do the second synchronize here so blob matching never sees bbox_d == 0.
    m.suspend();
    m.synchronize();
    for (;;) {
        m.suspend();
        m.synchronize();
        unsigned st = m.get_state();
        if (st >= target || st == am::STATE_DONE) return st;
        if (st == am::STATE_BLOB_MATCHING) {
            if (!m.worker.blob_map) { run_steps(m, 1); continue; } // first match() call builds blob_map (thread.cpp:599-665)
            if (target == am::STATE_BLOB_MATCHING) return st;
            if (matched >= match_steps) {
                if (!r->skipped) { m.next_state(); r->skipped = true; m.synchronize(); }
                run_steps(m, 1);
                continue;
            }
            uint64_t n = std::min<uint64_t>(chunk, match_steps - matched);
            run_steps(m, n);
            matched += n;
            continue;
        }
        run_steps(m, chunk);
    }
}

double amref_true_cost(void *h) {
    // thread.cpp:1109-1125 restated on the morph-side mirror (never get_energy()'s absolute value).
    am::morph &m = R(h)->m;
    double e = 0;
    for (auto &kv : m.chains) {
        am::chain &c = kv.second;
        if (c.width <= 1 || c.height == 0) continue;
        for (size_t x = 0; x < c.width; ++x)
            for (size_t j = 0; j < c.height; ++j)
                e += (double) am::point_distance(c.points[x][j], c.points[x][(j + 1) % c.height]);
    }
    return e;
}

// ---------------------------------------------------------------- dumps
uint64_t amref_frame_count(void *h) { return R(h)->m.frames.size(); }
void amref_frame_keys(void *h, uint64_t *out) {
    size_t i = 0;
    for (auto &kv : R(h)->m.frames) out[i++] = kv.first;
}
uint64_t amref_pixel_count(void *h, uint64_t frame) { return R(h)->m.get_pixel_count(frame); }
uint32_t amref_get_pixel(void *h, uint64_t frame, uint64_t pos) { return pack(R(h)->m.get_pixel(frame, pos).c); }
// stored (possibly HSP) colour, as the worker sees it
uint32_t amref_stored_pixel(void *h, uint64_t frame, uint64_t pos, int *present) {
    am::morph &m = R(h)->m;
    *present = m.has_pixel(frame, pos) ? 1 : 0;
    if (!*present) return 0;
    return pack(m.frames[frame].pixels[pos].c);
}
void amref_average_pixel(void *h, uint64_t frame, uint16_t *xy, uint32_t *rgba) {
    am::pixel p = R(h)->m.get_average_pixel(frame);
    xy[0] = p.x; xy[1] = p.y; *rgba = pack(p.c);
}
void amref_frame_means(void *h, uint64_t frame, double *out6) {
    am::frame &f = R(h)->m.frames[frame];
    out6[0] = f.x; out6[1] = f.y; out6[2] = f.r; out6[3] = f.g; out6[4] = f.b; out6[5] = f.a;
}
void amref_bbox(void *h, uint16_t *out4) {
    am::morph &m = R(h)->m;
    out4[0] = m.bbox_x1; out4[1] = m.bbox_y1; out4[2] = m.bbox_x2; out4[3] = m.bbox_y2;
}
void amref_perlin(void *h, int which, int32_t *p512) {
    am::morph &m = R(h)->m;
    const PerlinNoise &pn = which ? m.slope_map : m.lag_map;
    for (int i = 0; i < 512; ++i) p512[i] = pn.p[i];
}

uint64_t amref_blob_count(void *h, uint64_t frame) { return R(h)->m.get_blob_count(frame); }
// stats: x,y,r,g,b,a ; meta: group, surface size
int amref_blob_info(void *h, uint64_t frame, uint64_t b, double *stats6, uint64_t *meta2) {
    const am::blob *bl = R(h)->m.get_blob(frame, b);
    if (!bl) return 0;
    stats6[0] = bl->x; stats6[1] = bl->y; stats6[2] = bl->r; stats6[3] = bl->g; stats6[4] = bl->b; stats6[5] = bl->a;
    meta2[0] = bl->group; meta2[1] = bl->surface.size();
    return 1;
}
void amref_blob_surface(void *h, uint64_t frame, uint64_t b, uint64_t *pos_out) {
    const am::blob *bl = R(h)->m.get_blob(frame, b);
    if (!bl) return;
    size_t i = 0;
    for (size_t p : bl->surface) pos_out[i++] = p;
}
// label image (w*h, row-major): blob index in the frame's blob vector, or -1 where no pixel
void amref_blob_labels(void *h, uint64_t frame, uint32_t w, uint32_t hh, int32_t *out) {
    am::morph &m = R(h)->m;
    for (size_t i = 0; i < size_t(w) * hh; ++i) out[i] = -1;
    if (!m.has_frame(frame)) return;
    std::vector<am::blob *> &bs = m.frames[frame].blobs;
    for (size_t b = 0; b < bs.size(); ++b) {
        if (!bs[b]) continue;
        for (size_t p : bs[b]->surface) {
            uint32_t x = p % 65536, y = p / 65536;
            if (x < w && y < hh) out[size_t(y) * w + x] = (int32_t) b;
        }
    }
}

uint64_t amref_chain_count(void *h) { return R(h)->m.chains.size(); }
// info: key, width, height, max_surface
void amref_chain_info(void *h, uint64_t idx, uint64_t *info4) {
    auto it = R(h)->m.chains.begin();
    std::advance(it, idx);
    info4[0] = it->first; info4[1] = it->second.width; info4[2] = it->second.height; info4[3] = it->second.max_surface;
}
// column-major words: out[j*width + x] = points[x][j].word  (flags byte masked to the 7 defined bytes)
void amref_chain_points(void *h, uint64_t idx, uint64_t *out) {
    auto it = R(h)->m.chains.begin();
    std::advance(it, idx);
    am::chain &c = it->second;
    for (size_t j = 0; j < c.height; ++j)
        for (size_t x = 0; x < c.width; ++x) {
            am::point p = c.points[x][j];
            uint64_t w = uint64_t(p.s.x) | (uint64_t(p.s.y) << 16) | (uint64_t(p.s.x_fract) << 32) |
                         (uint64_t(p.s.y_fract) << 40) | (uint64_t(p.s.flags) << 48);
            out[j * c.width + x] = w;
        }
}

// ---------------------------------------------------------------- imports
// Install blobs on the WORKER side for one frame (so synchronize() mirrors them):
// group[b], stats[6*b..], surface positions concatenated with offsets[nblobs+1].
// The worker's frame pixels are (re)sent first if needed.
static void ensure_worker_frames(am::morph &m) {
    if (m.worker.identifier == m.identifier && !m.worker.frames.empty()) return;
    m.worker.clear();
    m.refresh_frames();
    for (auto &kv : m.frames) m.worker.set_frame(kv.first, &kv.second);
    m.worker.set_identifier(m.identifier);
    m.worker.set_bbox(m.bbox_x1, m.bbox_y1, m.bbox_x2, m.bbox_y2);
}

void amref_import_blobs(void *h, uint64_t frame, uint64_t nblobs, const uint64_t *group, const double *stats,
                        const uint64_t *offsets, const uint64_t *positions) {
    am::morph &m = R(h)->m;
    ensure_worker_frames(m);
    am::frame &f = m.worker.frames[frame];
    while (!f.blobs.empty()) { delete f.blobs.back(); f.blobs.pop_back(); }
    for (uint64_t b = 0; b < nblobs; ++b) {
        am::blob *bl = new am::blob;
        bl->index = b;
        bl->group = group[b];
        bl->unified = true;
        bl->x = stats[6 * b + 0]; bl->y = stats[6 * b + 1];
        bl->r = stats[6 * b + 2]; bl->g = stats[6 * b + 3]; bl->b = stats[6 * b + 4]; bl->a = stats[6 * b + 5];
        for (uint64_t i = offsets[b]; i < offsets[b + 1]; ++i) bl->surface.insert(bl->surface.end(), (size_t) positions[i]);
        f.blobs.push_back(bl);
    }
}

// Install a chain table on the WORKER side; words column-major [j*width+x] as in amref_chain_points.
int amref_import_chain(void *h, uint64_t key, uint64_t width, uint64_t height, uint64_t max_surface, const uint64_t *words) {
    am::morph &m = R(h)->m;
    ensure_worker_frames(m);
    am::chain &c = m.worker.chains[key];
    if (!am::renew_chain(&c, width, height)) { m.worker.chains.erase(key); return 0; }
    c.max_surface = max_surface;
    for (size_t j = 0; j < height; ++j)
        for (size_t x = 0; x < width; ++x) {
            uint64_t w = words[j * width + x];
            am::point p; p.word = 0;
            p.s.x = w & 0xffff; p.s.y = (w >> 16) & 0xffff; p.s.x_fract = (w >> 32) & 0xff; p.s.y_fract = (w >> 40) & 0xff;
            p.s.flags = (w >> 48) & 0xff;
            c.points[x][j] = p;
        }
    return 1;
}

// After imports: put the worker into STATE_ATOM_MORPHING with consistent bookkeeping, then mirror.
void amref_finish_import(void *h) {
    am::morph &m = R(h)->m;
    ensure_worker_frames(m);
    m.worker.state = am::STATE_ATOM_MORPHING;
    m.worker.blob_map_w = m.worker.chains.size();
    m.worker.blob_map_h = m.worker.frames.size();
    m.worker.chain_map_e = 0.0;
    m.worker.counter = 0;
    R(h)->skipped = true;
    m.synchronize();
}

// ---------------------------------------------------------------- render
void amref_render(void *h, double t, uint32_t *rgba_out) {
    am::morph &m = R(h)->m;
    std::vector<am::pixel> v;
    m.get_pixels(t, &v);
    size_t w = m.get_width();
    for (auto &px : v) rgba_out[size_t(px.y) * w + px.x] = pack(px.c);
}
// per-blob fetch (appends, reference order); returns count or -1 when blob is null. xy_out: 2 u16 per pixel
int64_t amref_render_blob(void *h, uint64_t b, double t, uint64_t cap, uint16_t *xy_out, uint32_t *rgba_out, uint64_t *group) {
    am::morph &m = R(h)->m;
    std::vector<am::pixel> v;
    const am::blob *bl = m.get_pixels((size_t) b, t, &v);
    if (!bl) return -1;
    *group = bl->group;
    for (size_t i = 0; i < v.size() && i < cap; ++i) { xy_out[2 * i] = v[i].x; xy_out[2 * i + 1] = v[i].y; rgba_out[i] = pack(v[i].c); }
    return (int64_t) v.size();
}
double amref_get_time(void *h, uint64_t f, uint64_t total) { return R(h)->m.get_time(f, total); }
uint64_t amref_get_frame_key(void *h, double t) { return R(h)->m.get_frame_key(t); }
uint32_t amref_get_background(void *h, uint16_t x, uint16_t y, double t) { return pack(R(h)->m.get_background(x, y, t)); }

// interpolate helpers of the facade (morph.cpp:1467-1515)
uint64_t amref_interpolate_point(void *h, uint64_t w1, uint64_t w2, double weight) {
    am::point p1, p2; p1.word = 0; p2.word = 0;
    p1.s.x = w1 & 0xffff; p1.s.y = (w1 >> 16) & 0xffff; p1.s.x_fract = (w1 >> 32) & 0xff; p1.s.y_fract = (w1 >> 40) & 0xff;
    p2.s.x = w2 & 0xffff; p2.s.y = (w2 >> 16) & 0xffff; p2.s.x_fract = (w2 >> 32) & 0xff; p2.s.y_fract = (w2 >> 40) & 0xff;
    am::point p = R(h)->m.interpolate(p1, p2, weight);
    return uint64_t(p.s.x) | (uint64_t(p.s.y) << 16) | (uint64_t(p.s.x_fract) << 32) | (uint64_t(p.s.y_fract) << 40);
}
uint32_t amref_interpolate_color(void *h, uint32_t c1, uint32_t c2, double lag, double slope, double str, int plain) {
    if (plain) return pack(R(h)->m.interpolate(unpack(c1), unpack(c2), str));
    return pack(R(h)->m.interpolate(unpack(c1), unpack(c2), lag, slope, str));
}

// ---------------------------------------------------------------- pure functions (a-N)
uint32_t amref_rgb_to_hsp(uint32_t c) { return pack(am::rgb_to_hsp(unpack(c))); }
uint32_t amref_hsp_to_rgb(uint32_t c) { return pack(am::hsp_to_rgb(unpack(c))); }
double amref_color_distance(uint32_t a, uint32_t b) { return am::color_distance(unpack(a), unpack(b)); }
double amref_octave_noise(unsigned seed, double x, double y, int octaves) { PerlinNoise pn(seed); return pn.octaveNoise(x, y, octaves); }
void amref_spline_point(uint64_t n, const double *xs, const double *ys, double t, double *out2) {
    glnemo::CRSpline s;
    for (uint64_t i = 0; i < n; ++i) s.AddSplinePoint(glnemo::Vec3D(xs[i], ys[i], 0.0));
    glnemo::Vec3D v = s.GetInterpolatedSplinePoint(t);
    out2[0] = v.x; out2[1] = v.y;
}
uint64_t amref_point_distance(uint64_t w1, uint64_t w2) {
    am::point p1, p2; p1.word = 0; p2.word = 0;
    p1.s.x = w1 & 0xffff; p1.s.y = (w1 >> 16) & 0xffff; p1.s.x_fract = (w1 >> 32) & 0xff; p1.s.y_fract = (w1 >> 40) & 0xff;
    p2.s.x = w2 & 0xffff; p2.s.y = (w2 >> 16) & 0xffff; p2.s.x_fract = (w2 >> 32) & 0xff; p2.s.y_fract = (w2 >> 40) & 0xff;
    return am::point_distance(p1, p2);
}

// ---------------------------------------------------------------- timing legs (bench.py cpu_baseline / --impl reference)
// n steps of thread::morph() (thread.cpp:1043-1068) = n*max(1,threads)*cycle_length proposals on a 1-chain scene.
double amref_time_morph_steps(void *h, uint64_t nsteps) {
    am::morph &m = R(h)->m;
    auto t0 = std::chrono::steady_clock::now();
    run_steps(m, nsteps);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
double amref_time_render(void *h, double t, uint32_t *rgba_out) {
    auto t0 = std::chrono::steady_clock::now();
    amref_render(h, t, rgba_out);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
unsigned amref_hardware_concurrency() { return std::thread::hardware_concurrency(); }

// ---------------------------------------------------------------- standalone FluidModel (fluidmodel.cpp:165-580)
// particle record = 24 doubles:
// 0 x 1 y 2 u 3 v 4 gravity_x 5 gravity_y 6 freedom_r 7 active 8 mature 9 R 10 G 11 B 12 A 13 r 14 g 15 b 16 a
// 17 strength 18 source_owner 19 frame_key 20 source_pos 21 destination_pos 22 cx 23 cy
enum { FP_STRIDE = 24 };
// Node::Node() leaves r,g,b,a,weight uninitialised (fluidmodel.cpp:58-69) and Node::clear() only runs for
// nodes already on the active list, so a node's FIRST activation reads whatever new[] returned.  Fresh
// large allocations are zero pages, small ones are recycled heap: give every node the cleared state the
// algorithm assumes, so single-step comparisons are deterministic.
static void sanitize_nodes(FluidModel *fm) {
    for (unsigned i = 0; i < fm->gsizeX; ++i)
        for (unsigned j = 0; j < fm->gsizeY; ++j) if (!fm->grid[i][j].active) fm->grid[i][j].clear();
}
void *amref_fluid_create(unsigned gx, unsigned gy, unsigned n) {
    FluidModel *fm = new FluidModel(gx, gy, n);
    sanitize_nodes(fm);
    return fm;
}
void amref_morph_fluid_sanitize(void *h) { am::morph &m = R(h)->m; if (m.fluid) sanitize_nodes(m.fluid); }
void amref_fluid_destroy(void *f) { delete reinterpret_cast<FluidModel *>(f); }
static void put_particles(FluidModel *f, uint64_t n, const double *rec) {
    Particle *ps = f->getParticles();
    for (uint64_t i = 0; i < n; ++i) {
        Particle *p = ps + i;
        const double *r = rec + i * FP_STRIDE;
        p->clear();
        p->x = r[0]; p->y = r[1]; p->u = r[2]; p->v = r[3];
        p->gravity_x = r[4]; p->gravity_y = r[5]; p->freedom_r = r[6];
        p->active = r[7] != 0.0; p->mature = r[8] != 0.0;
        p->R = r[9]; p->G = r[10]; p->B = r[11]; p->A = r[12];
        p->r = r[13]; p->g = r[14]; p->b = r[15]; p->a = r[16];
        p->strength = r[17]; p->source_owner = r[18] != 0.0;
        p->frame_key = (size_t) r[19]; p->source_pos = (size_t) r[20]; p->destination_pos = (size_t) r[21];
        p->hack = false;
    }
}
static void get_particles(FluidModel *f, uint64_t n, double *rec) {
    Particle *ps = f->getParticles();
    for (uint64_t i = 0; i < n; ++i) {
        Particle *p = ps + i;
        double *r = rec + i * FP_STRIDE;
        r[0] = p->x; r[1] = p->y; r[2] = p->u; r[3] = p->v;
        r[4] = p->gravity_x; r[5] = p->gravity_y; r[6] = p->freedom_r;
        r[7] = p->active ? 1.0 : 0.0; r[8] = p->mature ? 1.0 : 0.0;
        r[9] = p->R; r[10] = p->G; r[11] = p->B; r[12] = p->A;
        r[13] = p->r; r[14] = p->g; r[15] = p->b; r[16] = p->a;
        r[17] = p->strength; r[18] = p->source_owner ? 1.0 : 0.0;
        r[19] = (double) p->frame_key; r[20] = (double) p->source_pos; r[21] = (double) p->destination_pos;
        r[22] = p->cx; r[23] = p->cy;
    }
}
void amref_fluid_set_particles(void *f, uint64_t n, const double *rec) { put_particles(reinterpret_cast<FluidModel *>(f), n, rec); }
void amref_fluid_get_particles(void *f, uint64_t n, double *rec) { get_particles(reinterpret_cast<FluidModel *>(f), n, rec); }
void amref_fluid_step(void *f, uint64_t steps_left, double freedom_radius, double t) {
    reinterpret_cast<FluidModel *>(f)->step(steps_left, freedom_radius, t);
}
// node record = 13 doubles: m d gx gy u v ax ay r g b a weight ; layout out[(j*gx+i)*13+k] (pos = j*gsizeX+i, fluidmodel.cpp:121)
void amref_fluid_get_nodes(void *f, double *out) {
    FluidModel *fm = reinterpret_cast<FluidModel *>(f);
    for (unsigned j = 0; j < fm->gsizeY; ++j)
        for (unsigned i = 0; i < fm->gsizeX; ++i) {
            Node &n = fm->grid[i][j];
            double *o = out + (size_t(j) * fm->gsizeX + i) * 13;
            if (!n.active) { for (int k = 0; k < 13; ++k) o[k] = 0.0; continue; }
            o[0] = n.m; o[1] = n.d; o[2] = n.gx; o[3] = n.gy; o[4] = n.u; o[5] = n.v; o[6] = n.ax; o[7] = n.ay;
            o[8] = n.r; o[9] = n.g; o[10] = n.b; o[11] = n.a; o[12] = n.weight;
        }
}
// the morph-owned fluid model (after a get_pixels with fluid>0)
uint64_t amref_morph_fluid_count(void *h) { am::morph &m = R(h)->m; return m.fluid ? m.fluid->get_particle_count() : 0; }
void amref_morph_fluid_get(void *h, uint64_t n, double *rec) { am::morph &m = R(h)->m; if (m.fluid) get_particles(m.fluid, n, rec); }
void amref_morph_fluid_dims(void *h, uint32_t *out2) { am::morph &m = R(h)->m; out2[0] = m.fluid ? m.fluid->gsizeX : 0; out2[1] = m.fluid ? m.fluid->gsizeY : 0; }

const char *amref_version() { return am::get_version(); }

} // extern "C"
