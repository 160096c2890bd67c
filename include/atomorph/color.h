/*
 * atomorph/color.h -- drop-in for the reference's color.h (color.h:7-30): the am::color value
 * type and the colour helpers that are part of the public namespace.  Implemented in
 * atomorph_b200/csrc/morph_host.cpp on top of the shared host/device math (amx_math.h).
 */
#ifndef ATOMORPH_B200_COLOR_H
#define ATOMORPH_B200_COLOR_H

#include <stdint.h>
#include <math.h>

namespace am {

typedef struct color {
    uint8_t r;
    uint8_t g;
    uint8_t b;
    uint8_t a;
} color;

color create_color(unsigned char r, unsigned char g, unsigned char b, unsigned char a);
color create_color(double r, double g, double b, double a);   // round(v * 255), color.cpp:22-29

inline double color_distance(color c1, color c2) {            // color.h:17-24
    int rd = c1.r - c2.r, gd = c1.g - c2.g, bd = c1.b - c2.b, ad = c1.a - c2.a;
    return sqrt((double) (rd * rd + gd * gd + bd * bd + ad * ad)) / 510.0;
}

color rgb_to_hsp(color c);                                    // color.cpp:31-44
color hsp_to_rgb(color c);                                    // color.cpp:46-59
void RGBtoHSP(double R, double G, double B, double *H, double *S, double *P);
void HSPtoRGB(double H, double S, double P, double *R, double *G, double *B);

}

#endif
