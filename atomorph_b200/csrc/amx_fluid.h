/*
 * amx_fluid.h -- device state of the MPM liquid (amx_fluid.cu: FluidModel::step) and of its morph driver
 * (amx_fluiddraw.cu: morph::step_fluid / update_particle / draw_fluid).  Internal, not part of the ABI.
 */
#ifndef AMX_FLUID_H
#define AMX_FLUID_H
#include "amx_engine.h"

namespace amx {

// particle fields (SoA doubles [PF_COUNT][n]): position, velocity, attractor ("gravity"), freedom radius multiplier,
// ideal colour R G B A (PF_RI..PF_AI), current colour r g b a (PF_R..PF_A), strength   (fluidmodel.h:33-79)
enum { PF_X, PF_Y, PF_U, PF_V, PF_GX, PF_GY, PF_FREE, PF_RI, PF_GI, PF_BI, PF_AI, PF_R, PF_G, PF_B, PF_A, PF_STRENGTH, PF_COUNT };
enum { NF_M, NF_D, NF_GX, NF_GY, NF_U, NF_V, NF_AX, NF_AY, NF_R, NF_G, NF_B, NF_A, NF_W, NF_COUNT };

struct FluidDraw;    // amx_fluiddraw.cu

struct Fluid {
    uint32_t gx = 0, gy = 0, n = 0;
    double *pf = nullptr;        // [PF_COUNT][n]
    uint8_t *active = nullptr, *mature = nullptr, *owner = nullptr;
    double *aux = nullptr;       // [3][n] frame_key, source_pos, destination_pos (as handed in / out by the C-ABI)
    double *nf = nullptr;        // [NF_COUNT][gx*gy]
    double *cell = nullptr;      // [5][gx*gy] colour sums per base cell of a step (k_fluid_p2g -> k_fluid_colour_box)
    // particles ordered by base cell, rebuilt every step (amx_fluid.cu: the node passes gather instead of scattering)
    uint32_t *sortbuf = nullptr; // cell key, place inside the cell, perm (ordered -> particle), rank (particle -> ordered): [5][n]
    uint32_t *cs = nullptr;      // [gx*gy + 1] first ordered particle of every base cell (monotone; [gx*gy] = active particles)
    uint32_t *cnt = nullptr;     // [gx*gy + 1] particles per base cell
    double   *sq = nullptr;      // [5][n] in sorted order: x, y and up to three per-particle factors of the pass at hand
    void     *sort_tmp = nullptr;
    size_t    sort_tmp_bytes = 0;
    uint64_t step_counter = 0;
    FluidDraw *draw = nullptr;   // driver state, allocated on the first fluid frame
};

int  fluid_step(Engine *E, uint64_t steps_left, double freedom_radius);     // amx_fluid.cu
int  fluid_alloc(Engine *E, uint32_t gsize_x, uint32_t gsize_y, uint32_t particle_count);
void fluid_draw_free(Fluid *F);                                            // amx_fluiddraw.cu

} // namespace amx
#endif
