/*
 * amx_morph.h -- flat C view of the am::morph facade (include/atomorph/morph.h), one function per
 * public member of the reference class (reference morph.h:11-76), for FFI bindings (ctypes, cgo,
 * JNI ...) that cannot bind a C++ class.  Same argument meaning and error behaviour as the C++
 * members; colours packed r | g<<8 | b<<16 | a<<24, key points packed as in amx.h.
 */
#ifndef AMX_MORPH_H
#define AMX_MORPH_H

#include <stdint.h>
#include <stddef.h>
#include "amx_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amx_morph amx_morph;

amx_morph *amx_morph_create(void);                                   /* morph::morph()          morph.cpp:9   */
void     amx_morph_destroy(amx_morph *m);                            /* morph::~morph()         morph.cpp:14  */
void     amx_morph_clear(amx_morph *m);                              /* clear()                 morph.cpp:18  */
const char *amx_morph_last_error(amx_morph *m);
void    *amx_morph_device_context(amx_morph *m);                     /* amx_ctx* for the batch API of amx.h    */

/* setters: one id per member (morph.h:52-76), value as double; AMX_P_SEED = set_seed */
void     amx_morph_set(amx_morph *m, int param_id, double value);

/* ingest */
int      amx_morph_add_pixel(amx_morph *m, uint64_t frame, uint16_t x, uint16_t y, uint32_t rgba);   /* morph.cpp:298 */
int      amx_morph_add_pixels(amx_morph *m, uint64_t frame, uint64_t n, const uint16_t *x, const uint16_t *y, const uint32_t *rgba);
int      amx_morph_add_frame(amx_morph *m, uint64_t frame);          /* morph.cpp:287 */
void     amx_morph_set_resolution(amx_morph *m, uint16_t w, uint16_t h);
uint16_t amx_morph_get_width(amx_morph *m);
uint16_t amx_morph_get_height(amx_morph *m);
uint64_t amx_morph_get_frame_count(amx_morph *m);
uint64_t amx_morph_get_pixel_count(amx_morph *m, uint64_t frame);

/* run control (morph.cpp:52-100) */
void     amx_morph_compute(amx_morph *m);
void     amx_morph_compute_seconds(amx_morph *m, double seconds);
void     amx_morph_iterate(amx_morph *m, uint64_t iterations);
void     amx_morph_suspend(amx_morph *m);
int      amx_morph_suspend_timeout(amx_morph *m, double timeout);
int      amx_morph_is_busy(amx_morph *m);
int      amx_morph_synchronize(amx_morph *m);                        /* morph.cpp:102 */
void     amx_morph_next_state(amx_morph *m);
unsigned amx_morph_get_state(amx_morph *m);
double   amx_morph_get_energy(amx_morph *m);

/* time mapping */
uint64_t amx_morph_get_frame_key(amx_morph *m, double t);            /* SIZE_MAX when there are no frames */
double   amx_morph_get_time(amx_morph *m, uint64_t frame, uint64_t total);
double   amx_morph_normalize_time(amx_morph *m, double t);

/* fetch */
int      amx_morph_get_pixels(amx_morph *m, double t, uint32_t *rgba_out /* width*height */);      /* morph.cpp:1405 */
/* per-blob fetch: returns the pixel count (appended order), -1 when the reference would return nullptr */
int64_t  amx_morph_get_pixels_blob(amx_morph *m, uint64_t blob, double t, uint64_t cap, uint16_t *xy_out, uint32_t *rgba_out, uint64_t *group);
uint32_t amx_morph_get_pixel(amx_morph *m, uint64_t frame, uint64_t position);
void     amx_morph_get_average_pixel(amx_morph *m, uint64_t frame, uint16_t xy[2], uint32_t *rgba);
void     amx_morph_get_average_pixel_blob(amx_morph *m, uint64_t frame, uint64_t blob, uint16_t xy[2], uint32_t *rgba);
uint32_t amx_morph_get_background(amx_morph *m, uint16_t x, uint16_t y, double t);
uint64_t amx_morph_get_blob_count(amx_morph *m, uint64_t frame);
uint64_t amx_morph_get_blob_count_all(amx_morph *m);
/* get_blob: stats6 = x,y,r,g,b,a ; meta2 = group, surface size ; returns 0 when nullptr */
int      amx_morph_get_blob(amx_morph *m, uint64_t frame, uint64_t blob, double stats6[6], uint64_t meta2[2]);
int      amx_morph_get_blob_surface(amx_morph *m, uint64_t frame, uint64_t blob, uint64_t *positions_out);
uint32_t amx_morph_blob2pixel(amx_morph *m, uint64_t frame, uint64_t blob, uint16_t xy[2]);

/* interpolate overloads (morph.cpp:1467-1515) */
uint64_t amx_morph_interpolate_point(amx_morph *m, uint64_t p1, uint64_t p2, double w);
uint32_t amx_morph_interpolate_color(amx_morph *m, uint32_t c1, uint32_t c2, double w);
uint32_t amx_morph_interpolate_color_eased(amx_morph *m, uint32_t c1, uint32_t c2, double lag, double slope, double w);

#ifdef __cplusplus
}
#endif
#endif
