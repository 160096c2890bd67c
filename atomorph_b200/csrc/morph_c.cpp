/*
 * morph_c.cpp -- flat C wrappers of am::morph (include/amx_morph.h).
 */
#include "../../include/atomorph/atomorph.h"
#include "../../include/amx_morph.h"

struct amx_morph { am::morph m; };

static inline uint32_t packc(am::color c) { return (uint32_t) c.r | ((uint32_t) c.g << 8) | ((uint32_t) c.b << 16) | ((uint32_t) c.a << 24); }
static inline am::color unpackc(uint32_t v) { return am::create_color((unsigned char) (v & 255), (unsigned char) ((v >> 8) & 255), (unsigned char) ((v >> 16) & 255), (unsigned char) (v >> 24)); }
static inline am::point unpackp(uint64_t w) {
    am::point p; p.word = 0;
    p.s.x = w & 0xffff; p.s.y = (w >> 16) & 0xffff; p.s.x_fract = (w >> 32) & 0xff; p.s.y_fract = (w >> 40) & 0xff; p.s.flags = (w >> 48) & 0xff;
    return p;
}
static inline uint64_t packp(am::point p) {
    return (uint64_t) p.s.x | ((uint64_t) p.s.y << 16) | ((uint64_t) p.s.x_fract << 32) | ((uint64_t) p.s.y_fract << 40) | ((uint64_t) p.s.flags << 48);
}
static size_t to_size(double v) { return v >= 1.8e19 ? SIZE_MAX : (v < 0 ? 0 : (size_t) v); }

extern "C" {

amx_morph *amx_morph_create(void) { return new amx_morph(); }
void amx_morph_destroy(amx_morph *m) { delete m; }
void amx_morph_clear(amx_morph *m) { m->m.clear(); }
const char *amx_morph_last_error(amx_morph *m) { return m->m.last_error(); }
void *amx_morph_device_context(amx_morph *m) { return m->m.device_context(); }

void amx_morph_set(amx_morph *h, int id, double v) {
    am::morph &m = h->m;
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

int amx_morph_add_pixel(amx_morph *m, uint64_t frame, uint16_t x, uint16_t y, uint32_t rgba) {
    return m->m.add_pixel((size_t) frame, am::create_pixel(x, y, unpackc(rgba))) ? 1 : 0;
}
int amx_morph_add_pixels(amx_morph *m, uint64_t frame, uint64_t n, const uint16_t *x, const uint16_t *y, const uint32_t *rgba) {
    for (uint64_t i = 0; i < n; ++i) m->m.add_pixel((size_t) frame, am::create_pixel(x[i], y[i], unpackc(rgba[i])));
    return 1;
}
int amx_morph_add_frame(amx_morph *m, uint64_t frame) { return m->m.add_frame((size_t) frame) ? 1 : 0; }
void amx_morph_set_resolution(amx_morph *m, uint16_t w, uint16_t h) { m->m.set_resolution(w, h); }
uint16_t amx_morph_get_width(amx_morph *m) { return m->m.get_width(); }
uint16_t amx_morph_get_height(amx_morph *m) { return m->m.get_height(); }
uint64_t amx_morph_get_frame_count(amx_morph *m) { return m->m.get_frame_count(); }
uint64_t amx_morph_get_pixel_count(amx_morph *m, uint64_t frame) { return m->m.get_pixel_count((size_t) frame); }

void amx_morph_compute(amx_morph *m) { m->m.compute(); }
void amx_morph_compute_seconds(amx_morph *m, double s) { m->m.compute(s); }
void amx_morph_iterate(amx_morph *m, uint64_t n) { m->m.iterate((size_t) n); }
void amx_morph_suspend(amx_morph *m) { m->m.suspend(); }
int amx_morph_suspend_timeout(amx_morph *m, double t) { return m->m.suspend(t) ? 1 : 0; }
int amx_morph_is_busy(amx_morph *m) { return m->m.is_busy() ? 1 : 0; }
int amx_morph_synchronize(amx_morph *m) { return m->m.synchronize() ? 1 : 0; }
void amx_morph_next_state(amx_morph *m) { m->m.next_state(); }
unsigned amx_morph_get_state(amx_morph *m) { return m->m.get_state(); }
double amx_morph_get_energy(amx_morph *m) { return m->m.get_energy(); }

uint64_t amx_morph_get_frame_key(amx_morph *m, double t) { return (uint64_t) m->m.get_frame_key(t); }
double amx_morph_get_time(amx_morph *m, uint64_t f, uint64_t total) { return m->m.get_time((size_t) f, (size_t) total); }
double amx_morph_normalize_time(amx_morph *m, double t) { return m->m.normalize_time(t); }

int amx_morph_get_pixels(amx_morph *m, double t, uint32_t *out) {
    std::vector<am::pixel> v;
    m->m.get_pixels(t, &v);
    size_t w = m->m.get_width();
    for (const am::pixel &px : v) out[(size_t) px.y * w + px.x] = packc(px.c);
    return (int) (m->m.last_error()[0] == 0);
}
int64_t amx_morph_get_pixels_blob(amx_morph *m, uint64_t blob, double t, uint64_t cap, uint16_t *xy, uint32_t *rgba, uint64_t *group) {
    std::vector<am::pixel> v;
    const am::blob *bl = m->m.get_pixels((size_t) blob, t, &v);
    if (!bl) return -1;
    if (group) *group = bl->group;
    for (size_t i = 0; i < v.size() && i < cap; ++i) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; rgba[i] = packc(v[i].c); }
    return (int64_t) v.size();
}
uint32_t amx_morph_get_pixel(amx_morph *m, uint64_t frame, uint64_t pos) { return packc(m->m.get_pixel((size_t) frame, (size_t) pos).c); }
void amx_morph_get_average_pixel(amx_morph *m, uint64_t frame, uint16_t xy[2], uint32_t *rgba) {
    am::pixel p = m->m.get_average_pixel((size_t) frame);
    xy[0] = p.x; xy[1] = p.y; *rgba = packc(p.c);
}
void amx_morph_get_average_pixel_blob(amx_morph *m, uint64_t frame, uint64_t blob, uint16_t xy[2], uint32_t *rgba) {
    am::pixel p = m->m.get_average_pixel((size_t) frame, (size_t) blob);
    xy[0] = p.x; xy[1] = p.y; *rgba = packc(p.c);
}
uint32_t amx_morph_get_background(amx_morph *m, uint16_t x, uint16_t y, double t) { return packc(m->m.get_background(x, y, t)); }
uint64_t amx_morph_get_blob_count(amx_morph *m, uint64_t frame) { return m->m.get_blob_count((size_t) frame); }
uint64_t amx_morph_get_blob_count_all(amx_morph *m) { return m->m.get_blob_count(); }
int amx_morph_get_blob(amx_morph *m, uint64_t frame, uint64_t blob, double s[6], uint64_t meta[2]) {
    const am::blob *bl = m->m.get_blob((size_t) frame, (size_t) blob);
    if (!bl) return 0;
    s[0] = bl->x; s[1] = bl->y; s[2] = bl->r; s[3] = bl->g; s[4] = bl->b; s[5] = bl->a;
    meta[0] = bl->group; meta[1] = bl->surface.size();
    return 1;
}
int amx_morph_get_blob_surface(amx_morph *m, uint64_t frame, uint64_t blob, uint64_t *out) {
    const am::blob *bl = m->m.get_blob((size_t) frame, (size_t) blob);
    if (!bl) return 0;
    size_t i = 0;
    for (size_t pos : bl->surface) out[i++] = pos;
    return 1;
}
uint32_t amx_morph_blob2pixel(amx_morph *m, uint64_t frame, uint64_t blob, uint16_t xy[2]) {
    const am::blob *bl = m->m.get_blob((size_t) frame, (size_t) blob);
    if (!bl) { xy[0] = xy[1] = 0; return 0; }
    am::pixel p = m->m.blob2pixel(bl);
    xy[0] = p.x; xy[1] = p.y;
    return packc(p.c);
}

uint64_t amx_morph_interpolate_point(amx_morph *m, uint64_t p1, uint64_t p2, double w) { return packp(m->m.interpolate(unpackp(p1), unpackp(p2), w)); }
uint32_t amx_morph_interpolate_color(amx_morph *m, uint32_t c1, uint32_t c2, double w) { return packc(m->m.interpolate(unpackc(c1), unpackc(c2), w)); }
uint32_t amx_morph_interpolate_color_eased(amx_morph *m, uint32_t c1, uint32_t c2, double lag, double slope, double w) {
    return packc(m->m.interpolate(unpackc(c1), unpackc(c2), lag, slope, w));
}

}
