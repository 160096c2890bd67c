/*
 * atomorph/atomorph.h -- drop-in for the reference's atomorph.h (live part, atomorph.h:233-360):
 * namespace am constants, value types and inline metrics.  Same names, layouts and meanings, so
 * that code written against the reference (demo/main.cpp) compiles unchanged; the pipeline
 * behind am::morph runs on the GPU through the C-ABI of include/amx.h.
 *
 * Not provided: the deprecated v0.51 API (atomorph.h:28-231) and the OpenCV variant (out of scope,
 * SURVEY.md section 2).
 */
#ifndef ATOMORPH_B200_ATOMORPH_H
#define ATOMORPH_B200_ATOMORPH_H

#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <iostream>     // the reference header pulls these in transitively (vec3d.h, <future>);
#include <future>       // callers such as demo/main.cpp rely on that
#include <thread>
#include <algorithm>
#include <vector>
#include <map>
#include <set>
#include <limits>
#include <random>
#include <string>
#include <assert.h>

#include "color.h"

namespace am {

// colour spaces / interpolation modes (atomorph.h:235-241)
const unsigned RGB    = 0;
const unsigned HSP    = 1;
const unsigned NONE   = 2;
const unsigned LINEAR = 3;
const unsigned SPLINE = 4;
const unsigned COSINE = 5;
const unsigned PERLIN = 6;
// pipeline states (atomorph.h:242-246)
const unsigned STATE_BLOB_DETECTION   = 0;
const unsigned STATE_BLOB_UNIFICATION = 1;
const unsigned STATE_BLOB_MATCHING    = 2;
const unsigned STATE_ATOM_MORPHING    = 3;
const unsigned STATE_DONE             = 4;
// key point flags (atomorph.h:247-248)
const unsigned char HAS_PIXEL = 1;
const unsigned char HAS_FLUID = 2;

typedef struct pixel {
    uint16_t x;
    uint16_t y;
    color    c;
} pixel;

typedef struct blob {
    size_t index;              // position in the frame's blob vector
    std::set<size_t> surface;  // pixel positions (xy2pos)
    std::set<size_t> border;   // kept for source compatibility; always empty here
    size_t group   = 0;
    bool   unified = false;
    double x, y, r, g, b, a;
} blob;

typedef union key_point {
    struct {
        uint16_t x;
        uint16_t y;
        uint8_t  x_fract;
        uint8_t  y_fract;
        uint8_t  flags;
    } s;
    uint64_t word;
} point;

const size_t WARN_POINTER_SIZE = 1;
const size_t WARN_PIXEL_SIZE   = 2;
const size_t WARN_POINT_SIZE   = 3;

const unsigned TEXTURE  = 0;
const unsigned AVERAGE  = 1;
const unsigned DISTINCT = 2;

pixel create_pixel(uint16_t x, uint16_t y, unsigned char r, unsigned char g, unsigned char b, unsigned char a);
pixel create_pixel(uint16_t x, uint16_t y, color c);

inline size_t xy2pos(uint16_t x, uint16_t y) { return (size_t) y * 65536u + x; }

inline void point2xy(point pt, float *x, float *y) {
    *x = pt.s.x + pt.s.x_fract / 256.0f;
    *y = pt.s.y + pt.s.y_fract / 256.0f;
}

inline uint32_t pixel_distance(pixel p1, pixel p2) {
    int32_t xd = p1.x - p2.x, yd = p1.y - p2.y;
    return xd * xd + yd * yd;
}

inline uint32_t approx_point_distance(point p1, point p2) {
    int32_t xd = p1.s.x - p2.s.x, yd = p1.s.y - p2.s.y;
    return xd * xd + yd * yd;
}

// squared travel in 1/256 px units: the atom matcher's cost (atomorph.h:334-339)
inline uint64_t point_distance(point p1, point p2) {
    int64_t xd = (256 * (int64_t) p1.s.x + p1.s.x_fract) - (256 * (int64_t) p2.s.x + p2.s.x_fract);
    int64_t yd = (256 * (int64_t) p1.s.y + p1.s.y_fract) - (256 * (int64_t) p2.s.y + p2.s.y_fract);
    return (uint64_t) (xd * xd) + (uint64_t) (yd * yd);
}

inline double distance(double x1, double y1, double x2, double y2) {
    double xd = x1 - x2, yd = y1 - y2;
    return xd * xd + yd * yd;
}

inline bool point_has_pixel(point p) { return (p.s.flags & HAS_PIXEL); }
inline bool point_has_fluid(point p) { return (p.s.flags & HAS_FLUID); }

const char *get_version();
size_t get_warning();
bool uses_opencv();

}

#include "morph.h"

#endif
