// Host harness for the Gram kernel's tile ownership and sweep order: includes the very header the
// kernel uses (csrc/gram_tiles.cuh, __host__ __device__ templates) and checks, for every rank of a
// world, that (1) tile_coords is a bijection of the tile index onto the tile grid, (2) over all ranks
// every 64x64 cell of the global matrix is produced exactly once (computed, or mirrored by exactly
// one tile), (3) in sym = 2 the k-th block of every rank's sweep targets pairwise distinct peers.
// Prints one line per (world, rank) with a checksum that tests/test_device_math_host.py compares
// with the Python restatement (sharding.gram_tile_roles).  Test infrastructure, not product.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define __forceinline__ inline
#ifndef __CUDACC__
#define __host__
#define __device__
#endif
#include "../../featureextraction.jl_b200/csrc/gram_tiles.cuh"

struct P { int tiles_x, tiles_y, sym, rank, world; long diag_shift; };

int main(int argc, char** argv) {
    const int tiles_y = argc > 1 ? atoi(argv[1]) : 3;
    int failures = 0;
    for (int sym = 1; sym <= 2; ++sym)
        for (int world = 1; world <= 8; ++world) {
            const int bw = 2 * tiles_y, tx_n = bw * world;
            std::vector<int> cells((size_t)tx_n * tx_n, 0);          // 64x64 cells of the global matrix
            std::vector<int> first_target((size_t)world * world, -1);
            for (int rank = 0; rank < world; ++rank) {
                P p{tx_n, tiles_y, sym, rank, world, (long)rank * tiles_y * FX_GRAM_BM};
                const long total = (long)tx_n * tiles_y;
                std::vector<int> seen((size_t)total, 0);
                long checksum = 0, executed = 0;
                for (long t = 0; t < total; ++t) {
                    long i0, j0; int tx;
                    tile_coords(p, t, i0, j0, tx);
                    const int ty = (int)(i0 / FX_GRAM_BM);
                    if (ty < 0 || ty >= tiles_y || tx < 0 || tx >= tx_n || j0 != (long)tx * FX_GRAM_BN) { ++failures; continue; }
                    seen[(size_t)ty * tx_n + tx]++;
                    int peer;
                    const int role = tile_role(p, ty, tx, peer);
                    checksum += (long)role * (ty * 131 + tx * 7 + 1) + (role == 2 ? 1000003L * peer : 0);
                    if (role == 0) continue;
                    ++executed;
                    const int gi = 2 * (rank * tiles_y + ty);
                    cells[(size_t)gi * tx_n + tx]++; cells[(size_t)(gi + 1) * tx_n + tx]++;
                    if (role == 2) {
                        if (sym == 1 && peer != rank) ++failures;
                        if (tx / bw != peer) ++failures;                 // mirrored tile lands in the block it came from
                        cells[(size_t)tx * tx_n + gi]++; cells[(size_t)tx * tx_n + gi + 1]++;
                        // block order index of this tile in the rank's sweep
                        const int kb = (int)(t / ((long)bw * tiles_y));
                        if (sym == 2 && peer != rank) {
                            int& ft = first_target[(size_t)kb * world + rank];
                            if (ft == -1) ft = peer;
                            else if (ft != peer) ++failures;             // one target per block of the sweep
                        }
                    }
                }
                for (long t = 0; t < total; ++t) if (seen[(size_t)t] != 1) ++failures;
                printf("sym=%d world=%d rank=%d executed=%ld checksum=%ld\n", sym, world, rank, executed, checksum);
            }
            if (sym == 2) {
                for (size_t c = 0; c < cells.size(); ++c) if (cells[c] != 1) { ++failures; break; }
                for (int kb = 0; kb < world; ++kb)                       // targets of sweep block kb: distinct
                    for (int a = 0; a < world; ++a)
                        for (int b = a + 1; b < world; ++b) {
                            const int ta = first_target[(size_t)kb * world + a], tb = first_target[(size_t)kb * world + b];
                            if (ta >= 0 && ta == tb) ++failures;
                        }
            } else {
                // sym = 1 (no peer access): every rank fills exactly its own rows
                for (int gi = 0; gi < tx_n; ++gi)
                    for (int tx = 0; tx < tx_n; ++tx)
                        if (cells[(size_t)gi * tx_n + tx] != 1) { ++failures; gi = tx_n; break; }
            }
        }
    printf("failures=%d\n", failures);
    return failures ? 1 : 0;
}
