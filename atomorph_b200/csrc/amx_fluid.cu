/*
 * amx_fluid.cu -- K5: the Material-Point-Method liquid step (SURVEY.md row a-F, M3).
 *
 * Reference: FluidModel::step (fluidmodel.cpp:165-580; Grant Kot's MPM with atomorph's
 * customisations: inactive particles, colour diffusion through nodes, attractor, freedom
 * radius).  The reference walks particles serially and scatters into AoS nodes.  Here the particles
 * are ordered by base cell once per step (positions only move in the last pass) and every
 * particle -> node transfer is a GATHER: one thread per node walks the particles of the 3 x 3
 * base cells whose quadratic B-spline stencil covers the node -- three contiguous runs of the
 * ordered list -- and sums their contributions in registers.  No floating-point atomics: the node
 * sums differ from the reference's only by summation order (<< 1e-5 relative; the only run-to-run
 * freedom left is the order of the few particles inside one cell).  The node colour, a strength-weighted running mean in
 * the reference (fluidmodel.cpp:234-244), is kept as sum(w*c) and sum(w): the G2P pass only ever
 * uses mean*weight (fluidmodel.cpp:463-469), which is that sum.
 *
 * Passes per step (algorithmic bytes: Np*(120 read + 64 written) + Ng*2*104, section 8d):
 *   order by cell (count, scan, place) | nodes <- mass/gradients, cells <- colour | colour box sum |
 *   particles: pressure + wall force | nodes <- acceleration / m | particles: velocity update |
 *   nodes <- momentum / m | colour box sum | G2P gather + move
 *
 * Wall clamping uses the counter-based RNG where the reference calls rand() (fluidmodel.cpp:553-566).
 */
#include <cub/cub.cuh>
#include "amx_engine.h"
#include "amx_fluid.h"

namespace amx {

struct PW { int cx, cy; double px[3], py[3], gx[3], gy[3]; };

__device__ __forceinline__ void particle_weights(double x, double y, PW &w) {
    // fluidmodel.cpp:195-218
    w.cx = (int) (x - 0.5);
    w.cy = (int) (y - 0.5);
    double t = (double) (unsigned) w.cx - x;
    w.px[0] = (0.5 * t * t + 1.5 * t + 1.125); w.gx[0] = (t + 1.5);
    t += 1.0;
    w.px[1] = (-t * t + 0.75); w.gx[1] = (-2.0 * t);
    t += 1.0;
    w.px[2] = (0.5 * t * t - 1.5 * t + 1.125); w.gx[2] = (t - 1.5);
    t = (double) (unsigned) w.cy - y;
    w.py[0] = (0.5 * t * t + 1.5 * t + 1.125); w.gy[0] = (t + 1.5);
    t += 1.0;
    w.py[1] = (-t * t + 0.75); w.gy[1] = (-2.0 * t);
    t += 1.0;
    w.py[2] = (0.5 * t * t - 1.5 * t + 1.125); w.gy[2] = (t - 1.5);
}

#define PFI(k) pf[(size_t) (k) * n + i]
#define NODE(k, idx) nf[(size_t) (k) * ng + (idx)]

// ---- ordering by base cell: a counting sort -----------------------------------------------------------------------------
// key = base cell (cy * gsx + cx) of an active particle, gsx * gsy for an inactive one (ordered behind everything).  One 32-bit
// atomic per particle counts the cell's particles and hands out the particle's place inside its cell (arrival order: the
// order of the few particles of ONE cell is the only thing that varies from run to run, i.e. node sums to ~1e-16).
__global__ void __launch_bounds__(256)
k_fluid_count(const double *__restrict__ pf, const uint8_t *__restrict__ active, uint32_t n, uint32_t gsx, uint32_t gsy,
              uint32_t *__restrict__ key, uint32_t *__restrict__ within, uint32_t *__restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = gsx * gsy;
    if (active[i]) {
        const int cx = (int) (PFI(PF_X) - 0.5), cy = (int) (PFI(PF_Y) - 0.5);
        if (cx >= 0 && cy >= 0 && (uint32_t) cx < gsx && (uint32_t) cy < gsy) k = (uint32_t) cy * gsx + (uint32_t) cx;
    }
    key[i] = k;
    within[i] = atomicAdd(&cnt[k], 1u);
}

// cs = exclusive scan of the counts: first ordered particle of every cell.  perm: ordered -> particle, rank: its inverse;
// ordered copies of the positions.
__global__ void __launch_bounds__(256)
k_fluid_place(const uint32_t *__restrict__ key, const uint32_t *__restrict__ within, const uint32_t *__restrict__ cs, const double *__restrict__ pf,
              uint32_t n, uint32_t *__restrict__ perm, uint32_t *__restrict__ rank, double *__restrict__ sx, double *__restrict__ sy) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = cs[key[i]] + within[i];
    perm[s] = i;
    rank[i] = s;
    sx[s] = PFI(PF_X);
    sy[s] = PFI(PF_Y);
}

// quadratic B-spline weight and gradient of stencil offset o (0, 1, 2) at t = cell - position, exactly the expressions of
// particle_weights: ((a t) t + b t) + c and g t + h with a = -1, b = 0 for the middle one
__device__ __forceinline__ void bspline(double t0, int o, double *p, double *g) {
    double t = t0;
    if (o >= 1) t += 1.0;
    if (o >= 2) t += 1.0;
    if (o == 1) { *p = (-t * t + 0.75); *g = (-2.0 * t); }
    else {
        const double lin = o == 0 ? 1.5 : -1.5;
        *p = (0.5 * t * t + lin * t + 1.125);
        *g = (t + lin);
    }
}

// One thread per NODE (X, Y): the particles whose 3 x 3 stencil covers it sit in base cells [X-2, X] x [Y-2, Y], i.e. in three
// contiguous runs of the ordered list.  MODE 0: mass, mass gradients (fluidmodel.cpp:222-228) and, for the node's own cell,
// the colour sums strength * (r, g, b, a, 1) of its mature particles (229-246: the same five values go to all nine nodes
// of a stencil, so node (X, Y) receives the 3 x 3 box sum of the cell sums -- k_fluid_colour_box).  MODE 1: acceleration
// from pressure and wall forces, divided by the mass (349-369); q0 = pressure, q1 = fx, q2 = fy.  MODE 2: momentum divided
// by the mass (423-443); q0 = u * mass, q1 = v * mass (zero for immature particles).
template <int MODE>
__global__ void __launch_bounds__(256)
k_fluid_nodes(const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ q0, const double *__restrict__ q1,
              const double *__restrict__ q2, const uint32_t *__restrict__ cs, const uint32_t *__restrict__ perm, const double *__restrict__ pf,
              const uint8_t *__restrict__ mature, uint32_t n, double *__restrict__ nf, double *__restrict__ cell, uint32_t gsx, uint32_t gsy) {
    const uint32_t X = blockIdx.x * 32u + (threadIdx.x & 31u), Y = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (X >= gsx || Y >= gsy) return;
    const size_t ng = (size_t) gsx * gsy, idx = (size_t) Y * gsx + X;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const uint32_t xa = X >= 2u ? X - 2u : 0u;
    for (uint32_t jj = 0; jj < 3u && jj <= Y; ++jj) {
        const uint32_t cy = Y - jj;
        const uint32_t s0 = cs[cy * gsx + xa], s1 = cs[cy * gsx + X + 1u];
        for (uint32_t s = s0; s < s1; ++s) {
            const double x = sx[s], y = sy[s];
            const int cx = (int) (x - 0.5);
            double px, gx, py, gy;
            bspline((double) (unsigned) cx - x, (int) X - cx, &px, &gx);
            bspline((double) cy - y, (int) jj, &py, &gy);
            if (MODE == 0) {
                a0 += (px * py) * 1.0;
                a1 += gx * py;
                a2 += px * gy;
            } else if (MODE == 1) {
                const double phi = px * py, pressure = q0[s];
                a0 += -((gx * py) * pressure) + q1[s] * phi;
                a1 += -((px * gy) * pressure) + q2[s] * phi;
            } else {
                const double phi = px * py;
                a0 += phi * q0[s];
                a1 += phi * q1[s];
            }
        }
    }
    if (MODE == 0) {
        NODE(NF_M, idx) = a0; NODE(NF_GX, idx) = a1; NODE(NF_GY, idx) = a2;
        // the node's own base cell: colour sums of its mature particles
        double cr = 0.0, cg = 0.0, cb = 0.0, ca = 0.0, cw = 0.0;
        for (uint32_t s = cs[idx], s1 = cs[idx + 1u]; s < s1; ++s) {
            const uint32_t i = perm[s];
            const double strength = PFI(PF_STRENGTH);
            if (mature[i] && strength > 0.0) {
                cr += strength * PFI(PF_R); cg += strength * PFI(PF_G); cb += strength * PFI(PF_B); ca += strength * PFI(PF_A); cw += strength;
            }
        }
        cell[0 * ng + idx] = cr; cell[1 * ng + idx] = cg; cell[2 * ng + idx] = cb; cell[3 * ng + idx] = ca; cell[4 * ng + idx] = cw;
    } else {
        const double m = NODE(NF_M, idx);
        if (m > 0.0) { a0 /= m; a1 /= m; }
        NODE(MODE == 1 ? NF_AX : NF_U, idx) = a0;
        NODE(MODE == 1 ? NF_AY : NF_V, idx) = a1;
    }
}

// node colour fields = 3x3 box sum of the cell sums: node (X, Y) <- cells [X-2, X] x [Y-2, Y].  One CTA = a 32x8 block of
// nodes; the 34x10 cells it needs are staged in shared memory per field.
__global__ void __launch_bounds__(256)
k_fluid_colour_box(const double *__restrict__ cell, double *__restrict__ nf, uint32_t gsx, uint32_t gsy) {
    __shared__ double s[10][34];
    const size_t ng = (size_t) gsx * gsy;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const int x0 = (int) (blockIdx.x * 32u) - 2, y0 = (int) (blockIdx.y * 8u) - 2;
    const uint32_t X = blockIdx.x * 32u + tx, Y = blockIdx.y * 8u + ty;
    for (int f = 0; f < 5; ++f) {
        for (uint32_t l = threadIdx.x; l < 340u; l += 256u) {
            const int cx = x0 + (int) (l % 34u), cy = y0 + (int) (l / 34u);
            s[l / 34u][l % 34u] = (cx >= 0 && cy >= 0 && (uint32_t) cx < gsx && (uint32_t) cy < gsy) ? cell[(size_t) f * ng + (size_t) cy * gsx + cx] : 0.0;
        }
        __syncthreads();
        if (X < gsx && Y < gsy) {
            double sum = 0.0;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) sum += s[ty + dy][tx + dx];
            NODE(NF_R + f, (size_t) Y * gsx + X) = sum;
        }
        __syncthreads();
    }
}

// The colour part of the gather (fluidmodel.cpp:462-475) sums the colour fields of the nine nodes of a particle's stencil
// UNWEIGHTED: a 3x3 box sum again, computed once per base cell (same operand order as the per-particle loop: x offset
// outer, y offset inner, so the sums are bit-identical) into `cell`, which k_fluid_colour_box has consumed by then.  A
// particle reads 5 values instead of 45.
__global__ void __launch_bounds__(256)
k_fluid_colour_gather(const double *__restrict__ nf, double *__restrict__ cell, uint32_t gsx, uint32_t gsy) {
    __shared__ double s[10][34];
    const size_t ng = (size_t) gsx * gsy;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const uint32_t X = blockIdx.x * 32u + tx, Y = blockIdx.y * 8u + ty;
    for (int f = 0; f < 5; ++f) {
        for (uint32_t l = threadIdx.x; l < 340u; l += 256u) {
            const uint32_t nx = blockIdx.x * 32u + l % 34u, ny = blockIdx.y * 8u + l / 34u;
            s[l / 34u][l % 34u] = (nx < gsx && ny < gsy) ? NODE(NF_R + f, (size_t) ny * gsx + nx) : 0.0;
        }
        __syncthreads();
        if (X < gsx && Y < gsy) {
            double sum = 0.0;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    sum += s[ty + jj][tx + ii];
                }
            cell[(size_t) f * ng + (size_t) Y * gsx + X] = sum;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128)
k_fluid_forces(const double *__restrict__ pf, const uint8_t *__restrict__ active, uint32_t n, const double *__restrict__ nf, uint32_t gsx, uint32_t gsy,
               const uint32_t *__restrict__ rank, double *__restrict__ q0, double *__restrict__ q1, double *__restrict__ q2) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
    size_t ng = (size_t) gsx * gsy;
    double x = PFI(PF_X), y = PFI(PF_Y);
    // fluidmodel.cpp:275-317 cubic density interpolant from the 4 corner nodes
    uint32_t cx = (uint32_t) (int) x, cy = (uint32_t) (int) y;
    uint32_t cxi = cx + 1, cyi = cy + 1;
    auto ld = [&](int k, uint32_t ix, uint32_t iy) -> double { return (ix < gsx && iy < gsy) ? NODE(k, (size_t) iy * gsx + ix) : 0.0; };
    double n01d = ld(NF_M, cx, cy), n01gx = ld(NF_GX, cx, cy), n01gy = ld(NF_GY, cx, cy);
    double n02d = ld(NF_M, cx, cyi), n02gx = ld(NF_GX, cx, cyi), n02gy = ld(NF_GY, cx, cyi);
    double n11d = ld(NF_M, cxi, cy), n11gx = ld(NF_GX, cxi, cy), n11gy = ld(NF_GY, cxi, cy);
    double n12d = ld(NF_M, cxi, cyi), n12gx = ld(NF_GX, cxi, cyi), n12gy = ld(NF_GY, cxi, cyi);
    double pdx = n11d - n01d, pdy = n02d - n01d;
    double C20 = 3.0 * pdx - n11gx - 2.0 * n01gx;
    double C02 = 3.0 * pdy - n02gy - 2.0 * n01gy;
    double C30 = -2.0 * pdx + n11gx + n01gx;
    double C03 = -2.0 * pdy + n02gy + n01gy;
    double csum1 = n01d + n01gy + C02 + C03;
    double csum2 = n01d + n01gx + C20 + C30;
    double C21 = 3.0 * n12d - 2.0 * n02gx - n12gx - 3.0 * csum1 - C20;
    double C31 = -2.0 * n12d + n02gx + n12gx + 2.0 * csum1 - C30;
    double C12 = 3.0 * n12d - 2.0 * n11gy - n12gy - 3.0 * csum2 - C02;
    double C13 = -2.0 * n12d + n11gy + n12gy + 2.0 * csum2 - C03;
    double C11 = n02gx - C13 - C12 - n01gx;
    double u = x - (double) cx, u2 = u * u, u3 = u * u2;
    double v = y - (double) cy, v2 = v * v, v3 = v * v2;
    double density = n01d + n01gx * u + n01gy * v + C20 * u2 + C02 * v2 + C30 * u3 + C03 * v3 + C21 * u2 * v + C31 * u3 * v +
                     C12 * u * v2 + C13 * u * v3 + C11 * u * v;
    double pressure = density - 1.0;
    if (pressure > 2.0) pressure = 2.0;
    double fx = 0.0, fy = 0.0;
    if (x < 4.0) fx += 1.0 * (4.0 - x);
    else if (x > (double) (gsx - 5)) fx += 1.0 * ((double) (gsx - 5) - x);
    if (y < 4.0) fy += 1.0 * (4.0 - y);
    else if (y > (double) (gsy - 5)) fy += 1.0 * ((double) (gsy - 5) - y);
    // the scatter of -(gradient * pressure) + force * phi into the nine nodes is done by k_fluid_nodes<1> as a gather
    const uint32_t s = rank[i];
    q0[s] = pressure; q1[s] = fx; q2[s] = fy;
}

// attractor pull, shared by the velocity and the move pass (fluidmodel.cpp:385-413 / 505-548)
__device__ __forceinline__ void pull_towards(double x1, double y1, double x2, double y2, double a, double *ox, double *oy) {
    double A = fabs(y1 - y2), B = fabs(x1 - x2);
    double C = sqrt(A * A + B * B);
    if (a >= C) a = C;
    *ox = 0.0; *oy = 0.0;
    if (B <= 0.0) {
        if (y2 <= y1) *oy -= a; else *oy += a;
    } else if (C > 0.0) {
        double dx = (a * B) / C;
        double dy = (A * dx) / B;
        if (x1 <= x2) *ox += dx; else *ox -= dx;
        if (y1 <= y2) *oy += dy; else *oy -= dy;
    }
}

__global__ void __launch_bounds__(128)
k_fluid_velocity(double *__restrict__ pf, const uint8_t *__restrict__ active, const uint8_t *__restrict__ mature, uint32_t n,
                 const double *__restrict__ nf, uint32_t gsx, uint32_t gsy, const uint32_t *__restrict__ rank, double *__restrict__ q0, double *__restrict__ q1) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
    size_t ng = (size_t) gsx * gsy;
    double x = PFI(PF_X), y = PFI(PF_Y);
    PW w;
    particle_weights(x, y, w);
    double cax, cay;
    pull_towards(x, y, PFI(PF_GX), PFI(PF_GY), 0.03, &cax, &cay);
    double u = PFI(PF_U), v = PFI(PF_V);
    for (int ii = 0; ii < 3; ++ii)
        for (int jj = 0; jj < 3; ++jj) {
            uint32_t nx = (uint32_t) (w.cx + ii), ny = (uint32_t) (w.cy + jj);
            double ax = 0.0, ay = 0.0;
            if (nx < gsx && ny < gsy) { size_t idx = (size_t) ny * gsx + nx; ax = NODE(NF_AX, idx); ay = NODE(NF_AY, idx); }
            double phi = w.px[ii] * w.py[jj];
            u += phi * (ax + cax);
            v += phi * (ay + cay);
        }
    PFI(PF_U) = u;
    PFI(PF_V) = v;
    double mu = 1.0 * u, mv = 1.0 * v;
    if (!mature[i]) { mu *= 0.0; mv *= 0.0; }
    // the momentum scatter phi * (mu, mv) into the nine nodes is done by k_fluid_nodes<2> as a gather
    const uint32_t s = rank[i];
    q0[s] = mu; q1[s] = mv;
}

__global__ void __launch_bounds__(128)
k_fluid_g2p(double *__restrict__ pf, const uint8_t *__restrict__ active, const uint8_t *__restrict__ mature, uint32_t n,
            const double *__restrict__ nf, const double *__restrict__ cell, uint32_t gsx, uint32_t gsy, double steps_left, double freedom_radius, uint64_t seed, uint64_t stepno) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
    size_t ng = (size_t) gsx * gsy;
    double x = PFI(PF_X), y = PFI(PF_Y);
    PW w;
    particle_weights(x, y, w);
    double gu = 0.0, gv = 0.0, nR = 0.0, nG = 0.0, nB = 0.0, nA = 0.0, weight = 0.0;
    // the whole stencil inside the grid (always, away from the walls): the colour sums of the nine nodes come precomputed
    // per base cell (k_fluid_colour_gather); a node without colour weight holds zeros, so the reference's `if (weight > 0)`
    // per node changes nothing
    const bool boxed = w.cx >= 0 && w.cy >= 0 && (uint32_t) w.cx + 2u < gsx && (uint32_t) w.cy + 2u < gsy;
    if (boxed) {
        const size_t cidx = (size_t) w.cy * gsx + (size_t) w.cx;
        nR = cell[0 * ng + cidx]; nG = cell[1 * ng + cidx]; nB = cell[2 * ng + cidx]; nA = cell[3 * ng + cidx]; weight = cell[4 * ng + cidx];
    }
    for (int ii = 0; ii < 3; ++ii)
        for (int jj = 0; jj < 3; ++jj) {
            uint32_t nx = (uint32_t) (w.cx + ii), ny = (uint32_t) (w.cy + jj);
            if (nx >= gsx || ny >= gsy) continue;
            size_t idx = (size_t) ny * gsx + nx;
            double phi = w.px[ii] * w.py[jj];
            gu += phi * NODE(NF_U, idx);
            gv += phi * NODE(NF_V, idx);
            if (!boxed) {
                double nw = NODE(NF_W, idx);
                if (nw > 0.0) { weight += nw; nR += NODE(NF_R, idx); nG += NODE(NF_G, idx); nB += NODE(NF_B, idx); nA += NODE(NF_A, idx); }
            }
        }
    if (weight > 0.0) {          // fluidmodel.cpp:476-498
        nR /= weight; nG /= weight; nB /= weight; nA /= weight;
        if (!mature[i]) { PFI(PF_R) = nR; PFI(PF_G) = nG; PFI(PF_B) = nB; PFI(PF_A) = nA; }
        else {
            double wr = fabs(nR - PFI(PF_RI)), wg = fabs(nG - PFI(PF_GI)), wb = fabs(nB - PFI(PF_BI)), wa = fabs(nA - PFI(PF_AI));
            PFI(PF_R) = (1.0 - wr) * PFI(PF_R) + wr * nR;
            PFI(PF_G) = (1.0 - wg) * PFI(PF_G) + wg * nG;
            PFI(PF_B) = (1.0 - wb) * PFI(PF_B) + wb * nB;
            PFI(PF_A) = (1.0 - wa) * PFI(PF_A) + wa * nA;
        }
    }
    x += gu; y += gv;
    {   // freedom radius pull (fluidmodel.cpp:505-548)
        double x2 = PFI(PF_GX), y2 = PFI(PF_GY);
        double A = fabs(y - y2), B = fabs(x - x2);
        double C = sqrt(A * A + B * B);
        double r = freedom_radius * PFI(PF_FREE);
        if (C > r) {
            double mx, my;
            pull_towards(x, y, x2, y2, C - r, &mx, &my);
            double ww = 1.0 / (steps_left + 1.0);
            x += mx * ww; y += my * ww;
        }
    }
    double u = PFI(PF_U), v = PFI(PF_V);
    u += gu - u; v += gv - v;
    uint64_t rr = rng64(seed, 0xf1u + stepno, i);
    double j1 = (double) (rr & 0x7fffffffu) / 2147483647.0 * 0.01, j2 = (double) ((rr >> 32) & 0x7fffffffu) / 2147483647.0 * 0.01;
    if (x < 1.0) { x = 1.0 + j1; u = 0.0; }
    else if (x > (double) (gsx - 2)) { x = (double) (gsx - 2) - j1; u = 0.0; }
    if (y < 1.0) { y = 1.0 + j2; v = 0.0; }
    else if (y > (double) (gsy - 2)) { y = (double) (gsy - 2) - j2; v = 0.0; }
    PFI(PF_X) = x; PFI(PF_Y) = y; PFI(PF_U) = u; PFI(PF_V) = v;
}

void engine_fluid_free(Engine *E) {
    if (!E->fluid) return;
    Fluid *F = E->fluid;
    fluid_draw_free(F);
    dev_free(F->pf); dev_free(F->active); dev_free(F->mature); dev_free(F->owner); dev_free(F->aux); dev_free(F->nf); dev_free(F->cell);
    dev_free(F->sortbuf); dev_free(F->cs); dev_free(F->cnt); dev_free(F->sq); dev_free(F->sort_tmp);
    delete F;
    E->fluid = nullptr;
}

int fluid_step(Engine *E, uint64_t steps_left, double freedom_radius) {
    Fluid *F = E->fluid;
    const size_t ng = (size_t) F->gx * F->gy;
    const uint32_t n = F->n;
    if (n == 0) { cudaMemsetAsync(F->nf, 0, ng * NF_COUNT * 8, E->stream); return E->check("fluid step") ? AMX_ERR_CUDA : AMX_OK; }
    uint32_t *key = F->sortbuf, *within = key + n, *perm = within + n, *rank = perm + n;
    double *sx = F->sq, *sy = sx + n, *q0 = sy + n, *q1 = q0 + n, *q2 = q1 + n;
    const dim3 ngrid(div_up(F->gx, 32), div_up(F->gy, 8));
    // order the particles by base cell: their positions do not change before the last pass of the step
    cudaMemsetAsync(F->cnt, 0, (ng + 1) * 4, E->stream);
    k_fluid_count<<<div_up(n, 256), 256, 0, E->stream>>>(F->pf, F->active, n, F->gx, F->gy, key, within, F->cnt);
    size_t tmp = F->sort_tmp_bytes;
    cub::DeviceScan::ExclusiveSum(F->sort_tmp, tmp, F->cnt, F->cs, (int) (ng + 1), E->stream);
    k_fluid_place<<<div_up(n, 256), 256, 0, E->stream>>>(key, within, F->cs, F->pf, n, perm, rank, sx, sy);
    // P2G: mass, gradients; colour through the cell sums and their 3 x 3 box sum
    k_fluid_nodes<0><<<ngrid, 256, 0, E->stream>>>(sx, sy, q0, q1, q2, F->cs, perm, F->pf, F->mature, n, F->nf, F->cell, F->gx, F->gy);
    k_fluid_colour_box<<<ngrid, 256, 0, E->stream>>>(F->cell, F->nf, F->gx, F->gy);
    // forces -> node acceleration
    k_fluid_forces<<<div_up(n, 128), 128, 0, E->stream>>>(F->pf, F->active, n, F->nf, F->gx, F->gy, rank, q0, q1, q2);
    k_fluid_nodes<1><<<ngrid, 256, 0, E->stream>>>(sx, sy, q0, q1, q2, F->cs, perm, F->pf, F->mature, n, F->nf, F->cell, F->gx, F->gy);
    // particle velocities -> node velocity
    k_fluid_velocity<<<div_up(n, 128), 128, 0, E->stream>>>(F->pf, F->active, F->mature, n, F->nf, F->gx, F->gy, rank, q0, q1);
    k_fluid_nodes<2><<<ngrid, 256, 0, E->stream>>>(sx, sy, q0, q1, q2, F->cs, perm, F->pf, F->mature, n, F->nf, F->cell, F->gx, F->gy);
    // G2P
    k_fluid_colour_gather<<<ngrid, 256, 0, E->stream>>>(F->nf, F->cell, F->gx, F->gy);
    k_fluid_g2p<<<div_up(n, 128), 128, 0, E->stream>>>(F->pf, F->active, F->mature, n, F->nf, F->cell, F->gx, F->gy, (double) steps_left, freedom_radius,
                                                      E->p.seed, F->step_counter++);
    E->launches += 11;
    return E->check("fluid step") ? AMX_ERR_CUDA : AMX_OK;
}

int fluid_alloc(Engine *E, uint32_t gsize_x, uint32_t gsize_y, uint32_t particle_count) {
    engine_fluid_free(E);
    Fluid *F = new Fluid();
    F->gx = gsize_x; F->gy = gsize_y; F->n = particle_count;
    size_t ng = (size_t) gsize_x * gsize_y, n = particle_count;
    E->fluid = F;
    if (!dev_alloc(E, (void **) &F->pf, n * PF_COUNT * 8, "fluid particles") || !dev_alloc(E, (void **) &F->active, n, "fluid active") ||
        !dev_alloc(E, (void **) &F->mature, n, "fluid mature") || !dev_alloc(E, (void **) &F->owner, n, "fluid owner") ||
        !dev_alloc(E, (void **) &F->aux, n * 3 * 8, "fluid aux") || !dev_alloc(E, (void **) &F->nf, ng * NF_COUNT * 8, "fluid nodes") ||
        !dev_alloc(E, (void **) &F->cell, ng * 5 * 8, "fluid cell sums") ||
        !dev_alloc(E, (void **) &F->sortbuf, (n ? n : 1) * 5 * 4, "fluid order") || !dev_alloc(E, (void **) &F->cs, (ng + 2) * 4, "fluid cell starts") || !dev_alloc(E, (void **) &F->cnt, (ng + 2) * 4, "fluid cell counts") ||
        !dev_alloc(E, (void **) &F->sq, (n ? n : 1) * 5 * 8, "fluid ordered factors")) {
        engine_fluid_free(E);
        return AMX_ERR_NOMEM;
    }
    F->sort_tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, F->sort_tmp_bytes, (uint32_t *) nullptr, (uint32_t *) nullptr, (int) (ng + 1), E->stream);
    if (!dev_alloc(E, &F->sort_tmp, F->sort_tmp_bytes, "fluid sort workspace")) { engine_fluid_free(E); return AMX_ERR_NOMEM; }
    cudaMemsetAsync(F->pf, 0, n * PF_COUNT * 8 + 0, E->stream);
    cudaMemsetAsync(F->active, 0, n, E->stream);
    cudaMemsetAsync(F->mature, 0, n, E->stream);
    cudaMemsetAsync(F->owner, 0, n, E->stream);
    cudaMemsetAsync(F->aux, 0, n * 3 * 8, E->stream);
    cudaMemsetAsync(F->nf, 0, ng * NF_COUNT * 8, E->stream);
    return E->fail(cudaStreamSynchronize(E->stream), "fluid create") ? AMX_ERR_CUDA : AMX_OK;
}

} // namespace amx

using namespace amx;
extern "C" {

int amx_fluid_create(amx_ctx *ctx, uint32_t gsize_x, uint32_t gsize_y, uint32_t particle_count) {
    if (!ctx || gsize_x < 8 || gsize_y < 8) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return fluid_alloc(&ctx->e, gsize_x, gsize_y, particle_count);
}

// record field -> SoA slot
static const int rec2pf[17] = {PF_X, PF_Y, PF_U, PF_V, PF_GX, PF_GY, PF_FREE, -1, -1, PF_RI, PF_GI, PF_BI, PF_AI, PF_R, PF_G, PF_B, PF_A};

int amx_fluid_set_particles(amx_ctx *ctx, uint32_t n, const double *rec) {
    if (!ctx || !ctx->e.fluid || !rec || n != ctx->e.fluid->n) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    Fluid *F = E->fluid;
    std::vector<double> soa((size_t) n * PF_COUNT), aux((size_t) n * 3);
    std::vector<uint8_t> act(n), mat(n), own(n);
    for (uint32_t i = 0; i < n; ++i) {
        const double *r = rec + (size_t) i * AMX_FP_STRIDE;
        for (int k = 0; k < 17; ++k) if (rec2pf[k] >= 0) soa[(size_t) rec2pf[k] * n + i] = r[k];
        soa[(size_t) PF_STRENGTH * n + i] = r[17];
        act[i] = r[7] != 0.0; mat[i] = r[8] != 0.0; own[i] = r[18] != 0.0;
        aux[i] = r[19]; aux[(size_t) n + i] = r[20]; aux[(size_t) 2 * n + i] = r[21];
    }
    cudaMemcpyAsync(F->pf, soa.data(), soa.size() * 8, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(F->active, act.data(), n, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(F->mature, mat.data(), n, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(F->owner, own.data(), n, cudaMemcpyHostToDevice, E->stream);
    cudaMemcpyAsync(F->aux, aux.data(), aux.size() * 8, cudaMemcpyHostToDevice, E->stream);
    return E->fail(cudaStreamSynchronize(E->stream), "fluid set") ? AMX_ERR_CUDA : AMX_OK;
}

int amx_fluid_get_particles(amx_ctx *ctx, uint32_t n, double *rec) {
    if (!ctx || !ctx->e.fluid || !rec || n != ctx->e.fluid->n) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    Fluid *F = E->fluid;
    std::vector<double> soa((size_t) n * PF_COUNT), aux((size_t) n * 3);
    std::vector<uint8_t> act(n), mat(n), own(n);
    cudaMemcpyAsync(soa.data(), F->pf, soa.size() * 8, cudaMemcpyDeviceToHost, E->stream);
    cudaMemcpyAsync(act.data(), F->active, n, cudaMemcpyDeviceToHost, E->stream);
    cudaMemcpyAsync(mat.data(), F->mature, n, cudaMemcpyDeviceToHost, E->stream);
    cudaMemcpyAsync(own.data(), F->owner, n, cudaMemcpyDeviceToHost, E->stream);
    cudaMemcpyAsync(aux.data(), F->aux, aux.size() * 8, cudaMemcpyDeviceToHost, E->stream);
    if (E->fail(cudaStreamSynchronize(E->stream), "fluid get")) return AMX_ERR_CUDA;
    for (uint32_t i = 0; i < n; ++i) {
        double *r = rec + (size_t) i * AMX_FP_STRIDE;
        for (int k = 0; k < 17; ++k) if (rec2pf[k] >= 0) r[k] = soa[(size_t) rec2pf[k] * n + i];
        r[17] = soa[(size_t) PF_STRENGTH * n + i];
        r[7] = act[i]; r[8] = mat[i]; r[18] = own[i];
        r[19] = aux[i]; r[20] = aux[(size_t) n + i]; r[21] = aux[(size_t) 2 * n + i];
        double x = r[0], y = r[1];
        r[22] = (double) (unsigned) (int) (x - 0.5); r[23] = (double) (unsigned) (int) (y - 0.5);
    }
    return AMX_OK;
}

int amx_fluid_step(amx_ctx *ctx, uint64_t steps_left, double freedom_radius, double t) {
    (void) t;   // morph_time is unused by the reference step as well (fluidmodel.cpp:165)
    if (!ctx || !ctx->e.fluid) return AMX_ERR_ARG;
    cudaSetDevice(ctx->e.device);
    return fluid_step(&ctx->e, steps_left, freedom_radius);
}

int amx_fluid_get_nodes(amx_ctx *ctx, double *out) {
    if (!ctx || !ctx->e.fluid || !out) return AMX_ERR_ARG;
    Engine *E = &ctx->e;
    cudaSetDevice(E->device);
    Fluid *F = E->fluid;
    size_t ng = (size_t) F->gx * F->gy;
    std::vector<double> soa(ng * NF_COUNT);
    if (E->fail(cudaMemcpyAsync(soa.data(), F->nf, soa.size() * 8, cudaMemcpyDeviceToHost, E->stream), "nodes D2H") ||
        E->fail(cudaStreamSynchronize(E->stream), "nodes"))
        return AMX_ERR_CUDA;
    for (size_t idx = 0; idx < ng; ++idx) {
        double *o = out + idx * 13;
        for (int k = 0; k < 13; ++k) o[k] = soa[(size_t) k * ng + idx];
        o[NF_D] = o[NF_M];                                            // one sum on the device (k_fluid_p2g)
        double w = o[12];
        if (w > 0.0) { o[8] /= w; o[9] /= w; o[10] /= w; o[11] /= w; }   // report the running mean like the reference node
    }
    return AMX_OK;
}

}
