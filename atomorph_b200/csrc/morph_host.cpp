/*
 * morph_host.cpp -- the am::morph facade (include/atomorph/morph.h) on top of the device engine
 * (include/amx.h).  Host side only: key-frame ingest, parameter mirror, run control with a
 * worker thread, lazy blob mirrors, single-value helpers.  Every pipeline stage and every
 * rendered frame comes from the CUDA library -- no CPU fallback.
 *
 * Reference behaviour followed (file:line under the reference tree):
 *   ingest            morph.cpp:287-337   add_frame / add_pixel (HSP store, bbox, running means)
 *   run control       morph.cpp:52-100    compute / iterate / suspend; thread.cpp:173-206 run loop
 *   synchronize       morph.cpp:102-269   parameter push, restart on identifier change, mirrors
 *   time mapping      morph.cpp:404-450, 1522-1538
 *   fetch             morph.cpp:339-392, 1405-1465
 */
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>

#include "../../include/atomorph/atomorph.h"
#include "../../include/amx.h"
#include "amx_math.h"

namespace am {

// ------------------------------------------------------------------ namespace-level helpers
static inline uint32_t packc(color c) { return amx::c_make(c.r, c.g, c.b, c.a); }
static inline color unpackc(uint32_t v) {
    color c;
    c.r = amx::c_r(v); c.g = amx::c_g(v); c.b = amx::c_b(v); c.a = amx::c_a(v);
    return c;
}

color create_color(unsigned char r, unsigned char g, unsigned char b, unsigned char a) {
    color c; c.r = r; c.g = g; c.b = b; c.a = a;
    return c;
}
color create_color(double r, double g, double b, double a) { return unpackc(amx::create_color_d(r, g, b, a)); }
color rgb_to_hsp(color c) { return unpackc(amx::rgb_to_hsp(packc(c))); }
color hsp_to_rgb(color c) { return unpackc(amx::hsp_to_rgb(packc(c))); }
void RGBtoHSP(double R, double G, double B, double *H, double *S, double *P) { amx::rgb_to_hsp_d(R, G, B, H, S, P); }
void HSPtoRGB(double H, double S, double P, double *R, double *G, double *B) { amx::hsp_to_rgb_d(H, S, P, R, G, B); }

pixel create_pixel(uint16_t x, uint16_t y, unsigned char r, unsigned char g, unsigned char b, unsigned char a) {
    pixel px; px.x = x; px.y = y; px.c = create_color(r, g, b, a);
    return px;
}
pixel create_pixel(uint16_t x, uint16_t y, color c) { return create_pixel(x, y, c.r, c.g, c.b, c.a); }

const char *get_version() { return "1.0"; }
size_t get_warning() {
    size_t flags = 0;
    if (sizeof(void *) < 8) flags = flags | WARN_POINTER_SIZE;
    if (sizeof(pixel) > 8) flags = flags | WARN_PIXEL_SIZE;
    if (sizeof(point) > 8) flags = flags | WARN_POINT_SIZE;
    return flags;
}
bool uses_opencv() { return false; }

// ------------------------------------------------------------------ host key frame
namespace {

struct HostFrame {
    uint32_t cap_w = 0, cap_h = 0;          // dense storage, grows by powers of two
    std::vector<uint32_t> rgba;             // RAW colours as passed to add_pixel
    std::vector<uint8_t>  present;
    size_t count = 0;
    double x = 0, y = 0, r = 0, g = 0, b = 0, a = 0;   // running means in the STORED colour space (morph.cpp:318-334)
    size_t index = 0;
    std::vector<blob *> blobs;              // mirror, rebuilt lazily after synchronize()
    bool blobs_stale = true;
    uint32_t blob_count = 0;

    bool has(uint32_t px, uint32_t py) const { return px < cap_w && py < cap_h && present[(size_t) py * cap_w + px]; }
    void grow(uint32_t need_w, uint32_t need_h) {
        uint32_t nw = cap_w ? cap_w : 64, nh = cap_h ? cap_h : 64;
        while (nw < need_w) nw *= 2;
        while (nh < need_h) nh *= 2;
        if (nw == cap_w && nh == cap_h) return;
        std::vector<uint32_t> nr((size_t) nw * nh, 0);
        std::vector<uint8_t> np((size_t) nw * nh, 0);
        for (uint32_t yy = 0; yy < cap_h; ++yy) {
            std::memcpy(&nr[(size_t) yy * nw], &rgba[(size_t) yy * cap_w], (size_t) cap_w * 4);
            std::memcpy(&np[(size_t) yy * nw], &present[(size_t) yy * cap_w], cap_w);
        }
        rgba.swap(nr); present.swap(np);
        cap_w = nw; cap_h = nh;
    }
    void free_blobs() {
        for (blob *b : blobs) delete b;
        blobs.clear();
    }
};

void perlin_table(unsigned seed, int *p) {          // perlin.cpp:11-17
    if (seed == 0) seed = std::mt19937::default_seed;
    std::iota(p, p + 256, 0);
    std::shuffle(p, p + 256, std::mt19937(seed));
    for (int i = 0; i < 256; ++i) p[256 + i] = p[i];
}

struct LibmCos { double operator()(double v) const { return ::cos(v); } };

} // namespace

struct morph::impl {
    amx_ctx *ctx = nullptr;
    std::string err;
    bool warned = false;
    std::mutex dev;                          // serialises device calls between the worker and the user thread

    std::map<size_t, HostFrame> frames;

    // parameter mirror, reference defaults morph.h:97-118
    unsigned char fading = PERLIN, motion = SPLINE, blob_delimiter = HSP;
    double        blob_threshold = 1.0;
    size_t        blob_number = 1, blob_max_size = SIZE_MAX, blob_min_size = 1, blob_box_samples = 10;
    uint16_t      blob_box_grip = UINT16_MAX;
    unsigned char blob_rgba_weight = 1, blob_size_weight = 1, blob_xy_weight = 1;
    size_t        degeneration = 0;
    uint16_t      density = 1;
    size_t        threads = 0, cycle_length = 1000, feather = 0;
    bool          keep_background = false, finite = false;
    unsigned      fluidsteps = 0, show_blobs = TEXTURE, seed = 0;

    uint16_t bbox_x1 = UINT16_MAX, bbox_y1 = UINT16_MAX, bbox_x2 = 0, bbox_y2 = 0;
    uint16_t width = 0, height = 0;
    size_t   identifier = 0, device_identifier = SIZE_MAX;
    unsigned state = STATE_BLOB_DETECTION;
    bool     skip_state = false;
    double   energy = 0.0;
    int      lag_p[512], slope_p[512];
    uint32_t canvas_w = 0, canvas_h = 0;     // of the last upload
    std::vector<int32_t> labels_scratch;

    // worker (thread.h:42-53, thread.cpp:173-206)
    std::thread worker;
    std::atomic<bool> running{false}, paused{true}, signal_pause{false}, signal_stop{false};
    std::atomic<size_t> iterations{0};
    std::atomic<double> seconds{0.0};

    impl() { perlin_table(0, lag_p); perlin_table(1, slope_p); }

    bool ensure_ctx() {
        if (ctx) return true;
        int rc = amx_create(&ctx, 0);
        if (rc != AMX_OK) {
            // Fail LOUDLY and make polling callers (demo/main.cpp:230-291 loops until STATE_DONE) terminate.
            ctx = nullptr;
            err = "no CUDA device: the morph pipeline has no CPU fallback";
            if (!warned) { fprintf(stderr, "atomorph-b200: %s\n", err.c_str()); warned = true; }
            state = STATE_DONE;
            return false;
        }
        return true;
    }
    bool ck(int rc, const char *what) {
        if (rc == AMX_OK) return true;
        err = std::string(what) + ": " + (ctx ? amx_last_error(ctx) : "no context");
        return false;
    }
    bool has_frame(size_t f) const { return frames.find(f) != frames.end(); }
    void refresh_frames() {                   // morph.cpp:271-285
        size_t index = 0;
        for (auto &kv : frames) {
            kv.second.index = index++;
            if (kv.second.count == 0) {
                kv.second.x = (bbox_x1 <= bbox_x2 ? (bbox_x1 + bbox_x2) / 2.0 : 0.0);
                kv.second.y = (bbox_y1 <= bbox_y2 ? (bbox_y1 + bbox_y2) / 2.0 : 0.0);
                kv.second.r = kv.second.g = kv.second.b = kv.second.a = 0.0;
            }
        }
    }
    bool busy() const { return running.load() && !paused.load(); }

    void run() {                               // thread.cpp:173-206
        auto start = std::chrono::steady_clock::now();
        while (!signal_stop.load()) {
            if (!signal_pause.load() && !paused.load()) {
                size_t it = iterations.load();
                size_t chunk = it > 0 ? std::min<size_t>(it, 64) : 16;
                {
                    std::lock_guard<std::mutex> lk(dev);
                    if (ctx) amx_step(ctx, chunk);
                }
                if (it > 0) {
                    iterations -= chunk;
                    if (iterations.load() == 0) signal_pause = true;
                }
                double s = seconds.load();
                if (s > 0.0) {
                    auto end = std::chrono::steady_clock::now();
                    if (std::chrono::duration<double>(end - start).count() > s) { signal_pause = true; seconds = 0.0; }
                }
            } else {
                paused = true;
                signal_pause = false;
                std::this_thread::sleep_for(std::chrono::microseconds(50));
                start = std::chrono::steady_clock::now();
            }
        }
        running = false; signal_stop = false; signal_pause = false; paused = true;
    }
    void start_or_resume() {
        if (!running.load()) {
            if (worker.joinable()) worker.join();
            running = true;
            paused = true;
            worker = std::thread(&impl::run, this);
        }
        signal_pause = false;
        paused = false;
    }
    void stop_worker() {
        if (worker.joinable()) {
            signal_stop = true;
            worker.join();
        }
        running = false; paused = true; signal_stop = false; signal_pause = false;
    }

    // blob mirror of one frame from the device (morph.cpp:143-172)
    bool load_blobs(HostFrame &f) {
        uint32_t n = 0;
        if (!ck(amx_blob_count(ctx, (uint32_t) f.index, &n), "blob_count")) return false;
        std::vector<double> stats((size_t) n * 6 + 6);
        std::vector<uint64_t> meta((size_t) n * 2 + 2);
        labels_scratch.assign((size_t) canvas_w * canvas_h, -1);
        if (!ck(amx_export_blobs(ctx, (uint32_t) f.index, labels_scratch.data(), stats.data(), meta.data()), "export_blobs")) return false;
        f.free_blobs();
        for (uint32_t b = 0; b < n; ++b) {
            blob *bl = new blob;
            bl->index = b;
            bl->group = (size_t) meta[2 * b];
            bl->unified = true;
            bl->x = stats[6 * b + 0]; bl->y = stats[6 * b + 1];
            bl->r = stats[6 * b + 2]; bl->g = stats[6 * b + 3]; bl->b = stats[6 * b + 4]; bl->a = stats[6 * b + 5];
            f.blobs.push_back(bl);
        }
        for (uint32_t yy = 0; yy < canvas_h; ++yy)
            for (uint32_t xx = 0; xx < canvas_w; ++xx) {
                int32_t b = labels_scratch[(size_t) yy * canvas_w + xx];
                if (b >= 0 && (uint32_t) b < n) f.blobs[b]->surface.insert(f.blobs[b]->surface.end(), xy2pos(xx, yy));
            }
        f.blob_count = n;
        return true;
    }

    size_t frame_key_at(double t) {            // morph.cpp:404-423
        double integ;
        t = std::modf(t, &integ);
        if (t < 0.0) t += 1.0;
        size_t f = t * frames.size();
        size_t i = 0;
        for (auto &kv : frames) if (i++ == f) return kv.first;
        return SIZE_MAX;
    }
    color raw_get_pixel(size_t f, size_t pos) {   // morph.cpp:378-392
        auto it = frames.find(f);
        uint32_t px = pos % 65536, py = pos / 65536;
        if (it == frames.end() || !it->second.has(px, py)) return create_color((unsigned char) 0, 0, 0, 0);
        uint32_t raw = it->second.rgba[(size_t) py * it->second.cap_w + px];
        if (blob_delimiter == HSP) return unpackc(amx::hsp_to_rgb(amx::rgb_to_hsp(raw)));
        return unpackc(raw);
    }
    double ease(double lag, double slope, double str) {
        if (fading == COSINE || fading == PERLIN) return amx::ease_strength(lag, slope, str, LibmCos());
        return str;
    }
};

// ------------------------------------------------------------------ lifetime
morph::morph() : p(new impl) {}

morph::~morph() {
    clear();
    if (p->ctx) amx_destroy(p->ctx);
    delete p;
}

void morph::clear() {                       // morph.cpp:18-50 (without the reference's double free of `fluid`)
    p->stop_worker();
    for (auto &kv : p->frames) kv.second.free_blobs();
    p->frames.clear();
    p->identifier = 0;
    p->device_identifier = SIZE_MAX;
    p->energy = 0.0;
    p->state = STATE_BLOB_DETECTION;
    p->bbox_x1 = UINT16_MAX; p->bbox_y1 = UINT16_MAX; p->bbox_x2 = 0; p->bbox_y2 = 0;
    if (p->ctx) { std::lock_guard<std::mutex> lk(p->dev); amx_reset(p->ctx); }
}

const char *morph::last_error() const { return p->err.c_str(); }
void *morph::device_context() { p->ensure_ctx(); return p->ctx; }

// ------------------------------------------------------------------ ingest
bool morph::add_frame(size_t frame_key) {   // morph.cpp:287-296
    if (p->frames.find(frame_key) != p->frames.end()) return false;
    HostFrame &f = p->frames[frame_key];
    f.x = (p->bbox_x1 <= p->bbox_x2 ? (p->bbox_x1 + p->bbox_x2) / 2.0 : 0.0);
    f.y = (p->bbox_y1 <= p->bbox_y2 ? (p->bbox_y1 + p->bbox_y2) / 2.0 : 0.0);
    p->refresh_frames();
    return true;
}

bool morph::add_pixel(size_t frame_key, pixel px) {   // morph.cpp:298-337
    color stored = px.c;
    if (p->blob_delimiter == HSP) stored = rgb_to_hsp(px.c);
    add_frame(frame_key);
    HostFrame &f = p->frames[frame_key];
    if (px.x >= f.cap_w || px.y >= f.cap_h) f.grow((uint32_t) px.x + 1, (uint32_t) px.y + 1);
    size_t i = (size_t) px.y * f.cap_w + px.x;
    bool fresh = !f.present[i];
    f.rgba[i] = packc(px.c);
    f.present[i] = 1;
    if (fresh) f.count++;
    p->identifier++;
    if (px.x < p->bbox_x1) p->bbox_x1 = px.x;
    if (px.y < p->bbox_y1) p->bbox_y1 = px.y;
    if (px.x > p->bbox_x2) p->bbox_x2 = px.x;
    if (px.y > p->bbox_y2) p->bbox_y2 = px.y;
    if (f.count == 1) {
        f.x = px.x; f.y = px.y;
        f.r = stored.r / 255.0; f.g = stored.g / 255.0; f.b = stored.b / 255.0; f.a = stored.a / 255.0;
        return true;
    }
    double weight = 1.0 / double(f.count);
    f.x = (1.0 - weight) * f.x + weight * px.x;
    f.y = (1.0 - weight) * f.y + weight * px.y;
    f.r = (1.0 - weight) * f.r + weight * (stored.r / 255.0);
    f.g = (1.0 - weight) * f.g + weight * (stored.g / 255.0);
    f.b = (1.0 - weight) * f.b + weight * (stored.b / 255.0);
    f.a = (1.0 - weight) * f.a + weight * (stored.a / 255.0);
    return true;
}

void morph::set_resolution(uint16_t w, uint16_t h) { p->width = w; p->height = h; }
uint16_t morph::get_width() { return p->width; }
uint16_t morph::get_height() { return p->height; }
size_t morph::get_frame_count() { return p->frames.size(); }
size_t morph::get_pixel_count(size_t f) { return p->has_frame(f) ? p->frames[f].count : 0; }

// ------------------------------------------------------------------ setters (stored; pushed by synchronize)
void morph::set_blob_delimiter(unsigned char d) { p->blob_delimiter = d; }
void morph::set_blob_threshold(double t) { p->blob_threshold = t; }
void morph::set_blob_max_size(size_t s) { p->blob_max_size = s; }
void morph::set_blob_min_size(size_t s) { p->blob_min_size = s; }
void morph::set_blob_box_grip(uint16_t g) { p->blob_box_grip = g; }
void morph::set_blob_box_samples(size_t s) { p->blob_box_samples = s; }
void morph::set_blob_number(size_t n) { p->blob_number = n; }
void morph::set_blob_rgba_weight(unsigned char w) { p->blob_rgba_weight = w; }
void morph::set_blob_size_weight(unsigned char w) { p->blob_size_weight = w; }
void morph::set_blob_xy_weight(unsigned char w) { p->blob_xy_weight = w; }
void morph::set_degeneration(size_t d) { p->degeneration = d; }
void morph::set_motion(unsigned char m) { p->motion = m; }
void morph::set_fading(unsigned char f) { p->fading = f; }
void morph::set_threads(size_t t) { p->threads = t; }
void morph::set_cycle_length(size_t c) { p->cycle_length = c; }
void morph::set_feather(size_t f) { p->feather = f; }
void morph::set_keep_background(bool k) { p->keep_background = k; }
void morph::set_finite(bool f) { p->finite = f; }
void morph::set_show_blobs(unsigned b) { p->show_blobs = b; }

void morph::set_fluid(unsigned f) {         // morph.cpp:1553-1560
    if (p->fluidsteps != f) {
        p->fluidsteps = f;
        if (p->fluidsteps > 0 && p->density > 1) set_density(1);
    }
}
void morph::set_density(uint16_t d) {       // morph.cpp:1562-1570
    if (p->density != d) {
        p->density = d;
        p->identifier++;
        if (p->fluidsteps != 0 && p->density > 1) set_fluid(0);
    }
}
void morph::set_seed(unsigned seed) {       // morph.cpp:436-444
    perlin_table(seed, p->lag_p);
    perlin_table(seed + 1, p->slope_p);
    p->seed = seed;
}

// ------------------------------------------------------------------ run control
bool morph::is_busy() const { return p->busy(); }
unsigned morph::get_state() { return p->state; }
void morph::next_state() { p->skip_state = true; }
double morph::get_energy() { return p->energy; }

void morph::compute() {
    if (is_busy()) return;
    p->iterations = 0; p->seconds = 0.0;
    p->start_or_resume();
}
void morph::iterate(size_t iterations) {
    if (is_busy()) return;
    p->iterations = iterations; p->seconds = 0.0;
    if (iterations == 0) return;
    p->start_or_resume();
}
void morph::compute(double seconds) {
    if (is_busy()) return;
    p->iterations = 0; p->seconds = seconds;
    p->start_or_resume();
}
void morph::suspend() {
    while (is_busy()) {
        p->signal_pause = true;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
}
bool morph::suspend(double timeout) {
    if (!is_busy()) return true;
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        p->signal_pause = true;
        if (p->paused.load()) break;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout) return false;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    return true;
}

bool morph::synchronize() {                 // morph.cpp:102-269
    if (p->busy()) return false;
    if (!p->ensure_ctx()) return false;
    std::lock_guard<std::mutex> lk(p->dev);
    amx_ctx *c = p->ctx;
    impl &m = *p;
    bool ok = true;
    ok &= m.ck(amx_set_param(c, AMX_P_SEED, m.seed), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_DELIMITER, m.blob_delimiter), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_THRESHOLD, m.blob_threshold), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_MAX_SIZE, m.blob_max_size == SIZE_MAX ? 1.9e19 : (double) m.blob_max_size), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_MIN_SIZE, (double) m.blob_min_size), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_BOX_GRIP, m.blob_box_grip), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_BOX_SAMPLES, (double) m.blob_box_samples), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_NUMBER, (double) m.blob_number), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_RGBA_WEIGHT, m.blob_rgba_weight), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_SIZE_WEIGHT, m.blob_size_weight), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_BLOB_XY_WEIGHT, m.blob_xy_weight), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_DEGENERATION, (double) m.degeneration), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_DENSITY, m.density), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_THREADS, (double) m.threads), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_CYCLE_LENGTH, (double) m.cycle_length), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_MOTION, m.motion), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_FADING, m.fading), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_FEATHER, (double) m.feather), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_KEEP_BACKGROUND, m.keep_background), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_FINITE, m.finite), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_SHOW_BLOBS, m.show_blobs), "param");
    ok &= m.ck(amx_set_param(c, AMX_P_FLUID, m.fluidsteps), "param");
    if (!ok) return false;

    uint16_t bbox[4] = {m.bbox_x1, m.bbox_y1, m.bbox_x2, m.bbox_y2};
    if (m.identifier != m.device_identifier) {
        // key frames changed: restart everything on the device (morph.cpp:122-136)
        m.refresh_frames();
        uint32_t cw = m.width, ch = m.height;
        if (m.bbox_x1 <= m.bbox_x2) { cw = std::max<uint32_t>(cw, (uint32_t) m.bbox_x2 + 1); ch = std::max<uint32_t>(ch, (uint32_t) m.bbox_y2 + 1); }
        if (cw == 0 || ch == 0) { cw = std::max<uint32_t>(cw, 1); ch = std::max<uint32_t>(ch, 1); }
        if (!m.ck(amx_reset(c), "reset") || !m.ck(amx_set_canvas(c, m.width, m.height, cw, ch, bbox), "set_canvas")) return false;
        m.canvas_w = cw; m.canvas_h = ch;
        std::vector<uint64_t> keys;
        for (auto &kv : m.frames) keys.push_back(kv.first);
        if (!m.ck(amx_set_frame_count(c, (uint32_t) keys.size(), keys.data()), "set_frame_count")) return false;
        std::vector<uint32_t> rgba((size_t) cw * ch);
        std::vector<uint8_t> pres((size_t) cw * ch);
        for (auto &kv : m.frames) {
            HostFrame &f = kv.second;
            std::fill(rgba.begin(), rgba.end(), 0u);
            std::fill(pres.begin(), pres.end(), (uint8_t) 0);
            uint32_t copy_w = std::min(cw, f.cap_w), copy_h = std::min(ch, f.cap_h);
            for (uint32_t yy = 0; yy < copy_h; ++yy) {
                std::memcpy(&rgba[(size_t) yy * cw], &f.rgba[(size_t) yy * f.cap_w], (size_t) copy_w * 4);
                std::memcpy(&pres[(size_t) yy * cw], &f.present[(size_t) yy * f.cap_w], copy_w);
            }
            double means[6] = {f.x, f.y, f.r, f.g, f.b, f.a};
            if (!m.ck(amx_upload_frame(c, (uint32_t) f.index, rgba.data(), pres.data(), means), "upload_frame")) return false;
            f.free_blobs(); f.blobs_stale = true; f.blob_count = 0;
        }
        m.device_identifier = m.identifier;
    } else {
        if (!m.ck(amx_set_canvas(c, m.width, m.height, m.canvas_w, m.canvas_h, bbox), "set_canvas")) return false;
    }

    // load results (morph.cpp:138-232)
    m.state = amx_get_state(c);
    m.energy = amx_get_energy(c);
    for (auto &kv : m.frames) {
        kv.second.blobs_stale = true;
        uint32_t n = 0;
        amx_blob_count(c, (uint32_t) kv.second.index, &n);
        kv.second.blob_count = n;
    }
    if (m.skip_state) { amx_next_state(c); m.skip_state = false; }
    return true;
}

// ------------------------------------------------------------------ time mapping
size_t morph::get_frame_key(double t) { return p->frame_key_at(t); }
double morph::normalize_time(double t) {    // morph.cpp:446-450
    double integ, time = std::modf(t, &integ);
    if (time < 0.0) time += 1.0;
    return time;
}
double morph::get_time(size_t frame_index, size_t total_frames_out) {   // morph.cpp:1522-1538
    if (total_frames_out == 0) return 0.0;
    if (p->finite) {
        if (total_frames_out == 1) {
            if (get_frame_count() == 0) return 0.0;
            return (1.0 - (1.0 / double(get_frame_count()))) / 2.0;
        }
        double t = double(frame_index) / double(total_frames_out - 1);
        t *= 1.0 - (1.0 / double(get_frame_count()));
        return t;
    }
    return double(frame_index) / double(total_frames_out);
}

// ------------------------------------------------------------------ single-value fetches
pixel morph::get_pixel(size_t f, size_t pos) {
    pixel px = create_pixel(0, 0, 0, 0, 0, 0);
    auto it = p->frames.find(f);
    uint32_t x = pos % 65536, y = pos / 65536;
    if (it == p->frames.end() || !it->second.has(x, y)) return px;
    px.x = x; px.y = y;
    px.c = p->raw_get_pixel(f, pos);
    return px;
}

pixel morph::get_average_pixel(size_t f) {  // morph.cpp:339-354
    pixel px = create_pixel(0, 0, 0, 0, 0, 0);
    if (!p->has_frame(f)) return px;
    HostFrame &fr = p->frames[f];
    px.x = round(fr.x); px.y = round(fr.y);
    px.c.r = round(255.0 * fr.r); px.c.g = round(255.0 * fr.g); px.c.b = round(255.0 * fr.b); px.c.a = round(255.0 * fr.a);
    if (p->blob_delimiter == HSP) px.c = hsp_to_rgb(px.c);
    return px;
}

pixel morph::get_average_pixel(size_t f, size_t b) {   // morph.cpp:356-376
    pixel px = create_pixel(0, 0, 0, 0, 0, 0);
    const blob *bl = get_blob(f, b);
    if (!bl) return px;
    px.x = round(bl->x); px.y = round(bl->y);
    px.c.r = round(255.0 * bl->r); px.c.g = round(255.0 * bl->g); px.c.b = round(255.0 * bl->b); px.c.a = round(255.0 * bl->a);
    if (p->blob_delimiter == HSP) px.c = hsp_to_rgb(px.c);
    return px;
}

size_t morph::get_blob_count(size_t f) { return p->has_frame(f) ? p->frames[f].blob_count : 0; }
size_t morph::get_blob_count() {
    size_t n = 0;
    for (auto &kv : p->frames) n += kv.second.blob_count;
    return n;
}

const blob *morph::get_blob(size_t f, size_t b) {
    auto it = p->frames.find(f);
    if (it == p->frames.end() || !p->ctx || p->busy()) return nullptr;
    HostFrame &fr = it->second;
    if (b >= fr.blob_count) return nullptr;
    if (fr.blobs_stale || fr.blobs.size() != fr.blob_count) {
        std::lock_guard<std::mutex> lk(p->dev);
        if (!p->load_blobs(fr)) return nullptr;
        fr.blobs_stale = false;
    }
    return b < fr.blobs.size() ? fr.blobs[b] : nullptr;
}

pixel morph::blob2pixel(const blob *bl) {   // morph.cpp:1423-1429
    pixel px = create_pixel(bl->x, bl->y, create_color(bl->r, bl->g, bl->b, bl->a));
    if (p->blob_delimiter == HSP) px.c = hsp_to_rgb(px.c);
    return px;
}

// ------------------------------------------------------------------ interpolation helpers (morph.cpp:1467-1515)
color morph::interpolate(color c1, color c2, double c1_weight) { return unpackc(amx::lerp_color(packc(c1), packc(c2), c1_weight)); }
color morph::interpolate(color c1, color c2, double lag, double slope, double str) { return interpolate(c1, c2, p->ease(lag, slope, str)); }
pixel morph::interpolate(pixel px1, pixel px2, double w) {
    pixel px;
    px.x = std::round(w * double(px1.x) + (1.0 - w) * double(px2.x));
    px.y = std::round(w * double(px1.y) + (1.0 - w) * double(px2.y));
    px.c = interpolate(px1.c, px2.c, w);
    return px;
}
point morph::interpolate(point pt1, point pt2, double w) {
    point pt;
    uint32_t x, y, xf, yf;
    amx::lerp_point(amx::pw_make(pt1.s.x, pt1.s.y, pt1.s.x_fract, pt1.s.y_fract, 0), amx::pw_make(pt2.s.x, pt2.s.y, pt2.s.x_fract, pt2.s.y_fract, 0),
                    w, &x, &y, &xf, &yf);
    pt.word = 0;
    pt.s.x = x; pt.s.y = y; pt.s.x_fract = xf; pt.s.y_fract = yf;
    return pt;
}

color morph::get_background(uint16_t x, uint16_t y, double time) {   // morph.cpp:1431-1465
    if (p->frames.empty()) return create_color(0.0, 0.0, 0.0, 0.0);
    size_t position = xy2pos(x, y);
    time = normalize_time(time);
    double t = time;
    size_t frame_key = get_frame_key(t);
    auto it = p->frames.find(frame_key);
    if (it == p->frames.end()) return create_color(0.0, 0.0, 0.0, 0.0);
    ++it;
    if (it == p->frames.end()) it = p->frames.begin();
    size_t next_frame_key = it->first;
    double dt = 1.0 / double(p->frames.size());
    t = std::max(0.0, (t - (frame_key * dt)) / dt);
    color c1 = p->raw_get_pixel(frame_key, position), c2 = p->raw_get_pixel(next_frame_key, position);
    if (p->fading == PERLIN) {
        double f = 8.0;
        double bbox_w = p->bbox_x2 - p->bbox_x1 + 1.0, bbox_h = p->bbox_y2 - p->bbox_y1 + 1.0;
        double perlin_x = ((x - p->bbox_x1) / double(bbox_w)) * f, perlin_y = ((y - p->bbox_y1) / double(bbox_h)) * f;
        double lag = amx::pn_octave2(p->lag_p, perlin_x, perlin_y, 8) * 0.5 + 0.5;
        double slope = amx::pn_octave2(p->slope_p, perlin_x, perlin_y, 8) * 0.5 + 0.5;
        return interpolate(c1, c2, lag, slope, 1.0 - t);
    } else if (p->fading == COSINE) return interpolate(c1, c2, 0.5, 0.5, 1.0 - t);
    return interpolate(c1, c2, 1.0 - t);
}

// ------------------------------------------------------------------ frame fetch (morph.cpp:1405-1421)
void morph::get_pixels(double t, std::vector<pixel> *image) {
    image->clear();
    size_t w = p->width, h = p->height;
    image->resize(w * h);
    static_assert(sizeof(pixel) == 8, "am::pixel must stay 8 bytes (atomorph.cpp:1017-1023)");
    bool ok = false;
    if (w && h && p->ctx && p->device_identifier == p->identifier) {
        // the device writes finished am::pixel records: one copy straight into the caller's vector
        std::lock_guard<std::mutex> lk(p->dev);
        ok = p->ck(amx_render_pixels(p->ctx, t, reinterpret_cast<uint64_t *>(image->data())), "render");
    }
    if (!ok) {
        pixel *dst = image->data();
        for (size_t y = 0; y < h; ++y)
            for (size_t x = 0; x < w; ++x) {
                pixel &px = dst[y * w + x];
                px.x = (uint16_t) x; px.y = (uint16_t) y;
                px.c = unpackc(0u);
            }
    }
}

const blob *morph::get_pixels(size_t blob_index, double time, std::vector<pixel> *to) {   // morph.cpp:452-678
    if (p->frames.empty() || !p->ctx) return nullptr;
    double t = normalize_time(time);
    size_t frame_key = get_frame_key(t);
    if (frame_key == SIZE_MAX || blob_index >= get_blob_count(frame_key)) return nullptr;
    const blob *bl = get_blob(frame_key, blob_index);
    if (!bl) return nullptr;
    uint64_t cap = (uint64_t) p->canvas_w * p->canvas_h + 16;
    std::vector<uint16_t> xy(cap * 2);
    std::vector<uint32_t> col(cap);
    int64_t n = -1;
    uint64_t group = 0;
    {
        std::lock_guard<std::mutex> lk(p->dev);
        if (!p->ck(amx_render_blob(p->ctx, (uint32_t) blob_index, time, cap, xy.data(), col.data(), &n, &group), "render_blob")) return nullptr;
    }
    if (n < 0) return nullptr;
    for (int64_t i = 0; i < n; ++i) to->push_back(create_pixel(xy[2 * i], xy[2 * i + 1], unpackc(col[i])));
    return bl;
}

}
