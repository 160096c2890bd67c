/*
 * amx_swap.h -- pieces of the pair-swap matcher (amx_swap.cu) that the multi-GPU layer (amx_dist.cu) drives.
 */
#ifndef AMX_SWAP_H
#define AMX_SWAP_H
#include "amx_engine.h"

namespace amx {

#define TILE_BITS 10                       // default: 1024 atoms per tile: 16 KB (h = 2) / 24 KB (h >= 3) of shared memory
#define TILE_ATOMS (1u << TILE_BITS)
#define TILE_THREADS 256
#define TILE_MAX_ROUNDS 64                 // rounds per epoch (one load / store of the tile)
#define AMX_MAX_PEERS 7                    // other GPUs of one NVSwitch box

// Per-epoch bijection u -> atom on k-bit indices: two multiply / xorshift rounds (each step is invertible mod 2^k).
// Tile t owns u in [t * 2^TB, (t + 1) * 2^TB): a pseudo-random subset of the chain.
// Multi-GPU parts (amx_dist.cu): the OUTER bijection of a step splits the slots into contiguous per-rank parts; every
// sub-epoch re-tiles a part through an INNER bijection of the low `ik` bits (imask != 0), so a rank refines its part for
// several epochs without touching another rank's atoms.
// Locality epochs (default off): `perm` lists the atoms in Morton order of their column-y position
// (padding slots hold 0xffffffff) and slot u is atom perm[(u + shift) & mask], so a tile holds 1024 spatial neighbours.
struct TileMap {
    uint32_t a1, a2, c, s1, s2, mask;
    uint32_t ia1, ia2, ic, is1, is2, imask;
    const uint32_t *perm; uint32_t shift;
};
// Where a tile's refined key points go besides the local column (peer-mapped device memory of the other GPUs of the box,
// amx_dist.cu): the exchange of the multi-GPU matcher happens inside the swap kernel, tile by tile over NVLink, while the
// next tiles are still being refined -- not in a collective after it.
//   staged == 0: dst[p] is the same column of peer p's table; a key point is stored at its atom index (scattered 8-byte
//                stores; used where a rank refines whole columns and the copy is done by k_push_column instead)
//   staged == 1: dst[p] is peer p's staging buffer; a tile's key points are stored in SLOT order, i.e. as one contiguous
//                8 KB run per tile (full NVLink write packets); the peer scatters them into its column after the barrier
struct PeerCols { pword *dst[AMX_MAX_PEERS]; uint32_t n, staged; };


unsigned ceil_log2(uint64_t w);
bool tiled_ok(Engine *E, uint32_t chain);
TileMap make_tilemap(uint64_t seed, uint64_t stream, uint64_t epoch, unsigned k);
void tilemap_set_inner(TileMap &tm, uint64_t seed, uint64_t stream, uint64_t sub, unsigned ik);
void launch_swap_tiled(Engine *E, bool h2, int tb, pword *col, const pword *prev, const pword *next, uint64_t off, uint32_t w, const TileMap &tm,
                       uint32_t t0, uint32_t ntl, uint32_t rounds, uint64_t round_base, const PeerCols &peers);
void launch_pack_tiled(Engine *E, const pword *col, uint64_t off, uint32_t w, const TileMap &tm, uint32_t u0, uint32_t n, pword *out);
void launch_unpack_tiled(Engine *E, pword *col, uint64_t off, uint32_t w, const TileMap &tm, uint32_t u0, uint32_t n, uint32_t skip0, uint32_t skip1,
                         const pword *in);

} // namespace amx
#endif
