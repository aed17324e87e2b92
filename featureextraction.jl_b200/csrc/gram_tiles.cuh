// gram_tiles.cuh -- tile ownership and sweep order of the Gram kernel (gram.cu), kept in a header
// of __host__ __device__ templates so that the very same code is exercised on the CPU
// (tests/host/gram_tiles_host.cu, driven by tests/test_device_math_host.py).  P is any struct
// with the fields tiles_x, tiles_y, sym, rank, world, diag_shift (GramParams in gram.cu).
#pragma once

constexpr int FX_GRAM_BM = 128, FX_GRAM_BN = 64;

// What this rank does with tile (ty, tx): 0 = skip (someone's mirror store produces it),
// 1 = compute, 2 = compute and mirror into block `peer` (== own block inside the diagonal block).
template <class P>
__host__ __device__ __forceinline__ int tile_role(const P& p, int ty, int tx, int& peer) {
    constexpr int BN = FX_GRAM_BN;
    peer = p.rank;
    if (!p.sym) return 1;
    const int bw = 2 * p.tiles_y;                                 // tile-columns per rank block
    const int c = (p.sym == 2) ? tx / bw : ((tx >= (int)(p.diag_shift / BN) && tx < (int)(p.diag_shift / BN) + bw) ? p.rank : -1);
    if (c == p.rank) {
        const int rel = tx - (int)(p.diag_shift / BN);
        if (rel < 2 * ty) return 0;                               // left of the diagonal
        return rel >= 2 * ty + 2 ? 2 : 1;
    }
    if (p.sym != 2) return 1;
    const int delta = (c - p.rank + p.world) % p.world;
    bool own;
    if (2 * delta < p.world) own = true;
    else if (2 * delta > p.world) own = false;
    else own = (((ty + (tx - c * bw) / 2) & 1) == (p.rank < c ? 0 : 1));   // opposite blocks: checkerboard
    peer = c;
    return own ? 2 : 0;
}

// Tile order: column panels of PANEL tile-columns swept over all row blocks, so that the panel's
// slab of the column operand (PANEL x 131 KB at d=256) stays L2-resident while the row operand
// streams; a plain row-major sweep re-reads all of X from HBM once per row block.
// sym == 2: the sweep goes rank block by rank block and starts at the rank's OWN block
// (c = rank, rank+1, ... mod world), so that at any moment the ranks mirror into pairwise distinct
// peers (a permutation over NVSwitch) instead of all into the same one.
constexpr int FX_GRAM_PANEL = 128;
template <class P>
__host__ __device__ __forceinline__ void tile_coords(const P& p, long tile, long& i0, long& j0, int& tx) {
    constexpr int PANEL = FX_GRAM_PANEL;
    const int nblk = (p.sym == 2) ? p.world : 1;
    const int bw = p.tiles_x / nblk;                           // tile-columns per block
    const long per_block = (long)bw * p.tiles_y;
    const int kb = (int)(tile / per_block);
    const long inblk = tile - (long)kb * per_block;
    const long per_panel = (long)PANEL * p.tiles_y;
    const int panel = (int)(inblk / per_panel);
    const long within = inblk - (long)panel * per_panel;
    const int rest = bw - panel * PANEL;
    const int pw = rest < PANEL ? rest : PANEL;                // last panel of a block may be narrower
    const int ty = (int)(within / pw);
    const int blk = (p.sym == 2) ? (p.rank + kb) % p.world : 0;
    tx = blk * bw + panel * PANEL + (int)(within - (long)ty * pw);
    i0 = (long)ty * FX_GRAM_BM;
    j0 = (long)tx * FX_GRAM_BN;
}
