// gram.cu -- kernel-matrix construction: K = map(A B') with the kernel map, zero padding and
// per-tile row sums fused behind the contraction.  Replaces MLKernels.kernelmatrix as called at
// /root/reference/src/kpca.jl:51 (symmetric) and kpca.jl:65 (cross).
//
// Data path (persistent CTAs, 2 per SM, each looping over 128x64 output tiles):
//   producer warpgroup (one active warp; gives its registers to the MMA warps via setmaxnreg):
//     HBM --cp.async.bulk (TMA engine, UBLKCP) + mbarrier full/empty--> 3-stage smem ring of
//     16-feature slabs; it runs ahead across tile boundaries.
//   4 MMA warps: LDS.64 conflict-free (row stride = 4 mod 16 doubles) --> DMMA.8x8x4
//     (mma.sync f64), 64x32 register tile per warp, no CTA-wide barrier in the main loop.
//   epilogue, SOFTWARE-PIPELINED THROUGH TENSOR MEMORY: when a tile's contraction is done, each MMA
//     warp parks its 64x32 block of raw dot products in TMEM (four tcgen05.st.32x32b.x32 per warp:
//     lane L's 128 accumulator registers go to TMEM lane 32 (warp & 3) + L, 128 columns) and starts
//     the next tile at once.  The kernel map of the parked tile (branch-free exp / templated powers /
//     tanh), its row sums and its stores are then done two elements at a time (tcgen05.ld.x4) BETWEEN
//     THE DMMAs OF THE NEXT TILE'S MAIN LOOP, a fixed number of units per k-step in the same basic
//     block, so the map's FP64 operations (DMMA and DFMA share one datapath) fill issue gaps of the
//     DMMA stream instead of stopping it.  fp64 has no tcgen05.mma, so on this path TMEM is otherwise
//     idle: it serves as 128 KB/SM of register spill space with a one-instruction spill.
//     What this replaces (measured, profiles/r02_gram_design_experiments.md): the map as a burst in
//     the MMA warps after each tile (round 1) leaves the FP64 datapath idle whenever both warps of a
//     scheduler are in their bursts: 77-80 % of the DMMA peak; the map in separate epilogue warps
//     behind a shared-memory staging tile is starved by the DMMA stream and the 128 KB staging tile
//     costs the operand ring its depth: 73-78 %.
//   EDGE tiles (ragged borders, tiles touching the diagonal: O(n/128) of them) keep the burst
//     epilogue with validity masks and the same-point test.
//
// Symmetry.  K is symmetric when A is a row block of B.  sym = 1 (one rank, or no peer access):
// inside this rank's square diagonal block only the tiles on / right of the diagonal are computed
// and the strictly-right ones are mirrored by transposed stores from the fragments.  sym = 2
// (several ranks, peer arenas mapped through CUDA IPC): the off-diagonal blocks are shared out
// too -- of each pair of blocks (r,c) / (c,r) exactly one rank computes every 128x128 super-tile
// and stores the transposed values, and the column sums that are the peer's row-sum partials,
// STRAIGHT INTO THE PEER'S K over NVLink (512-byte cp.async.bulk stores from a per-warp transposition
// buffer to the IPC-mapped address), so every rank executes ~half of its row block and no separate
// exchange step exists.
#include <stdlib.h>
#include "fx_kernels.cuh"
#include "gram_tiles.cuh"

int fx_launch_gram_v3(fx_ctx* ctx, const GramArgs& a, int variant);      // gram_v3.cu (A/B only)

namespace {

constexpr int BM = FX_GRAM_BM, BN = FX_GRAM_BN, BK = 16, STAGES = 3;
constexpr int AS = BM + 4;                      // 132: (t*AS + g) mod 16 distinct for t,g in 0..3
constexpr int BS = BN + 4;                      // 68
constexpr int STAGE_D = BK * AS + BK * BS;      // doubles per stage (3200 -> 25600 B)
// mirror staging: per MMA warp two buffers of 8 transposed rows x 64 (+2 pad: conflict-free
// 8-byte stores for lanes (g,t) -> word (2t+q)*TS + g) -- sent off with 512-byte bulk stores
constexpr int TS = 66, TBUF_D = 8 * TS;
constexpr int UNITS = 32;                       // (mt, nt) accumulator pairs per lane per tile
// smem (doubles): stages | rs [2 bufs][2 halves][BM] | 16 (barriers + TMEM base) | tstage
// (114,816 bytes: two CTAs per SM exactly fit the 228 KB of an SM)
constexpr int OFF_RS = STAGES * STAGE_D;
constexpr int OFF_BAR = OFF_RS + 4 * BM;
constexpr int OFF_TST = OFF_BAR + 16;
constexpr int SMEM_BYTES = (OFF_TST + 4 * 2 * TBUF_D) * 8;

// 1-D bulk store shared -> global (TMA engine); the destination may be a peer GPU's memory
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     __cvta_generic_to_global(dst_gmem)),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}

struct GramParams {
    const double* A; long lda;
    const double* B; long ldb;
    const double* na; const double* nb;
    double* K; long ldk;
    double* rowpart; long rows_pad;
    long rows_valid, cols_valid, diag_shift;
    int kt;
    int tiles_x, tiles_y; long tiles_total;
    int sym;                          // 0 / 1 / 2, see the header comment
    int rank, world;                  // sym == 2: this rank's block index and the number of blocks
    double* peerK[16];                // sym == 2: K of rank c (own entry = K)
    double* peerRowpart[16];          //           rowpart of rank c (or null)
    KernelMap km;
};

// tile_role() / tile_coords(): gram_tiles.cuh (host-testable)

__device__ __forceinline__ void bar_sync_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---- tensor memory as accumulator spill space -------------------------------------------------
// One instruction parks a lane's 64 accumulators (128 x 32-bit registers) in TMEM: lane L of warp w
// owns TMEM lane 32 (w & 3) + L; accumulator pair (mt, nt) sits at columns 4 (4 mt + nt) .. + 3.
// (Issued as four .x32 pieces: ptxas allocates the operand block of a TMEM store against the
// kernel's LAUNCH register target -- 128 with two CTAs per SM -- not against the setmaxnreg budget.)
__device__ __forceinline__ void tmem_park32(uint32_t taddr, const uint32_t* r) {
#define R8(i) "r"(r[i]), "r"(r[i + 1]), "r"(r[i + 2]), "r"(r[i + 3]), "r"(r[i + 4]), "r"(r[i + 5]), "r"(r[i + 6]), "r"(r[i + 7])
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {"
        "%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), R8(0), R8(8), R8(16), R8(24)
        : "memory");
#undef R8
}
__device__ __forceinline__ void tmem_park(uint32_t taddr, const double (&acc)[8][4][2]) {
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = part * 16 + i;                        // element (mt, nt, q) in storage order
            const double v = acc[e >> 3][(e >> 1) & 3][e & 1];
            r[2 * i] = (uint32_t)__double2loint(v);
            r[2 * i + 1] = (uint32_t)__double2hiint(v);
        }
        tmem_park32(taddr + 32 * part, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ double2 tmem_fetch2(uint32_t taddr) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return make_double2(__hiloint2double((int)b, (int)a), __hiloint2double((int)d, (int)c));
}

// Map of one element; for the Gaussian kernel (KID 0) the norms arrive pre-scaled by -gamma.
template <int KID>
__device__ __forceinline__ double gram_map(const KernelMap& km, double dot, double ni, double nj, bool same) {
    if (KID == 0) return gaussian_map_prescaled(2.0 * km.p0, dot, ni, nj, same);
    return kernel_map_t<KID>(km, dot, ni, nj, same);
}

// ---- the parked tile ------------------------------------------------------------------------------
// Scalars only (a run-time index into a member array would put the struct in local memory).
struct Parked {
    long i0, j0;            // tile coordinates
    int tx;                 // tile-column index (row-sum slot)
    int q;                  // next unit (UNITS = done); unit q is accumulator pair (mt = q & 7, nt = q >> 3)
    bool valid, mirror;
    double* Kmir;           // mirror target: K / rowpart of the owner of the transposed tile, its first row
    double* rpmir;
    long mir_row0;
    double cs0, cs1;        // running column sums of the current 8-column group (mirror)
};

// One unit of the parked tile: fetch the raw dot products of accumulator pair (mt, nt) from TMEM,
// map, store, add to the row-sum slot; MIRROR tiles also stage the transposed values and, at the end
// of each 8-column group, send 8 transposed rows (512-byte bulk stores) and the group's column sums.
// Interior tiles only: no masks, no same-point test.
template <int KID, bool MIRROR>
__device__ __forceinline__ void parked_unit(const GramParams& p, Parked& pk, int q, uint32_t taddr, int wm, int wn, int g,
                                            int t, int lane, double* rsb, double* tbuf) {
    const int mt = q & 7, nt = q >> 3;
    const double2 raw = tmem_fetch2(taddr + 4 * (4 * mt + nt));
    const long i = pk.i0 + wm * 64 + mt * 8 + g;
    const long jb = pk.j0 + wn * 32 + nt * 8 + 2 * t;
    double ni = 0.0, nb0 = 0.0, nb1 = 0.0;
    if (KID == 0) {
        ni = -p.km.p0 * __ldg(p.na + i);
        const double2 v = __ldg(reinterpret_cast<const double2*>(p.nb + jb));
        nb0 = -p.km.p0 * v.x; nb1 = -p.km.p0 * v.y;
    }
    const double v0 = gram_map<KID>(p.km, raw.x, ni, nb0, false);
    const double v1 = gram_map<KID>(p.km, raw.y, ni, nb1, false);
    *reinterpret_cast<double2*>(p.K + i * p.ldk + jb) = make_double2(v0, v1);
    if (p.rowpart) {
        // row-sum slot of (column half wn, row): folded over the four t lanes, then accumulated by the
        // t == 0 lane alone, in unit order -- deterministic (no atomics)
        double rsum = v0 + v1;
        rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
        rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
        if (t == 0) rsb[wn * BM + wm * 64 + mt * 8 + g] += rsum;
    }
    if (MIRROR) {
        double* tb = tbuf + (nt & 1) * TBUF_D;
        if (mt == 0 && nt >= 2) {
            // the bulk stores issued from this buffer two groups ago must have read it
            if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        tb[(2 * t) * TS + mt * 8 + g] = v0;                     // transposed: row = column of the tile
        tb[(2 * t + 1) * TS + mt * 8 + g] = v1;
        pk.cs0 += v0;
        pk.cs1 += v1;
        if (mt == 7) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < 8) {
                // transposed row j = j0 + wn*32 + nt*8 + lane, columns [i0 + wm*64, +64) of the owner's block
                double* dst = pk.Kmir + (pk.j0 + wn * 32 + nt * 8 + lane - pk.mir_row0) * p.ldk + pk.i0 + wm * 64 + p.diag_shift;
                bulk_s2g(dst, tb + lane * TS, 64 * 8);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (pk.rpmir) {
                // column sums over this warp's 64 rows = row-sum partial of mirrored row j for the
                // 64-column slot [i0 + wm*64, +64) of that row: slot index = 2*(i0/BM) + wm
                const long slot = p.diag_shift / BN + 2 * (pk.i0 / BM) + wm;
                double cs0 = pk.cs0, cs1 = pk.cs1;
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
                if (g == 0) {
                    double* r = pk.rpmir + slot * p.rows_pad + (jb - pk.mir_row0);
                    *reinterpret_cast<double2*>(r) = make_double2(cs0, cs1);
                }
            }
            pk.cs0 = pk.cs1 = 0.0;
        }
    }
}

// All units of the parked tile are done: drain the mirror stores and -- after the 4 MMA warps met --
// write the tile's row-sum partials (and clear the slots for the tile after the next).
__device__ __forceinline__ void parked_finish(const GramParams& p, Parked& pk, int lane, int tid, double* rsb) {
    if (pk.mirror) {
        if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
    if (p.rowpart) {
        bar_sync_compute();                  // both column halves of every row are in smem
        p.rowpart[(long)pk.tx * p.rows_pad + pk.i0 + tid] = rsb[tid] + rsb[BM + tid];
        rsb[tid] = 0.0;                      // this buffer is used again by the tile after the next
        rsb[BM + tid] = 0.0;
    }
    pk.valid = false;
}

// Kernel map + row sums + stores for one warp's 64x32 accumulator tile.  EDGE tiles (ragged
// borders or touching the diagonal) carry the validity masks and the same-point test; interior
// tiles run the bare arithmetic.
template <int KID, bool EDGE>
__device__ __forceinline__ void gram_epilogue_plain(const GramParams& p, double (&acc)[8][4][2],
                                                    const double (&ni)[8], const double (&nbv)[4][2], long i0,
                                                    long j0, int wm, int wn, int g, int t, double* rsb) {
    const long jb = j0 + wn * 32 + 2 * t;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
        const long i = i0 + wm * 64 + mt * 8 + g;
        double* krow = p.K + i * p.ldk + jb;
        const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
        const bool iv = i < p.rows_valid;
        double sacc = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            double v0, v1;
            if (EDGE) {
                const long j = jb + nt * 8;
                v0 = gram_map<KID>(p.km, acc[mt][nt][0], ni[mt], nbv[nt][0], ig == j);
                v1 = gram_map<KID>(p.km, acc[mt][nt][1], ni[mt], nbv[nt][1], ig == j + 1);
                v0 = (iv && j < p.cols_valid) ? v0 : 0.0;
                v1 = (iv && j + 1 < p.cols_valid) ? v1 : 0.0;
            } else {
                v0 = gram_map<KID>(p.km, acc[mt][nt][0], ni[mt], nbv[nt][0], false);
                v1 = gram_map<KID>(p.km, acc[mt][nt][1], ni[mt], nbv[nt][1], false);
            }
            *reinterpret_cast<double2*>(krow + nt * 8) = make_double2(v0, v1);
            sacc += v0 + v1;
        }
        if (p.rowpart) {
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (t == 0) rsb[wn * BM + wm * 64 + mt * 8 + g] = sacc;
        }
    }
}

// The same for MIRROR tiles (symmetric mode): also sends the transposed tile to its owner.
template <int KID, bool EDGE>
__device__ __forceinline__ void gram_epilogue_mirror(const GramParams& p, double (&acc)[8][4][2],
                                              const double (&ni)[8], long i0,
                                              long j0, int wm, int wn, int g, int t, double* rsb,
                                              double* Kmir, double* rpmir, long mir_row0, double* tbuf) {
    // Kmir / rpmir: K and rowpart of the rank that owns rows [mir_row0, +rows_pad) -- this rank
    // itself inside the diagonal block, a peer (NVLink) for the shared-out blocks.  The mirrored
    // values go through a small transposed staging buffer (tbuf: 2 x [8][TS] per warp) and leave
    // as 512-byte bulk stores, 8 per 8-column group: direct transposed stores from the fragments
    // are 64-byte pieces, which NVLink carries at a fraction of its bandwidth.
    const long jb = j0 + wn * 32 + 2 * t;
    const int lane = g * 4 + t;
    double sacc[8];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) sacc[mt] = 0.0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        double* tb = tbuf + (nt & 1) * TBUF_D;
        double cs0 = 0.0, cs1 = 0.0;
        // column norms re-read per group (L1 hit): keeps 12 registers free for the row sums
        double nb0 = 0.0, nb1 = 0.0;
        if (KID == 0) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(p.nb + jb + nt * 8));
            nb0 = -p.km.p0 * v.x; nb1 = -p.km.p0 * v.y;
        }
        if (nt >= 2) {
            // the bulk stores issued from this buffer two groups ago must have read it
            if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            const long i = i0 + wm * 64 + mt * 8 + g;
            double v0, v1;
            if (EDGE) {
                const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
                const bool iv = i < p.rows_valid;
                const long j = jb + nt * 8;
                v0 = gram_map<KID>(p.km, acc[mt][nt][0], ni[mt], nb0, ig == j);
                v1 = gram_map<KID>(p.km, acc[mt][nt][1], ni[mt], nb1, ig == j + 1);
                v0 = (iv && j < p.cols_valid) ? v0 : 0.0;
                v1 = (iv && j + 1 < p.cols_valid) ? v1 : 0.0;
            } else {
                v0 = gram_map<KID>(p.km, acc[mt][nt][0], ni[mt], nb0, false);
                v1 = gram_map<KID>(p.km, acc[mt][nt][1], ni[mt], nb1, false);
            }
            *reinterpret_cast<double2*>(p.K + i * p.ldk + jb + nt * 8) = make_double2(v0, v1);
            tb[(2 * t) * TS + mt * 8 + g] = v0;              // transposed: row = column of the tile
            tb[(2 * t + 1) * TS + mt * 8 + g] = v1;
            cs0 += v0;
            cs1 += v1;
            sacc[mt] += v0 + v1;
        }
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < 8) {
                // transposed row j = j0 + wn*32 + nt*8 + lane, columns [i0 + wm*64, +64) of the owner's block
                double* dst = Kmir + (j0 + wn * 32 + nt * 8 + lane - mir_row0) * p.ldk + i0 + wm * 64 + p.diag_shift;
                bulk_s2g(dst, tb + lane * TS, 64 * 8);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (rpmir) {
                // column sums over this warp's 64 rows = row-sum partial of mirrored row j for the
                // 64-column slot [i0 + wm*64, +64) of that row: slot index = 2*(i0/BM) + wm
                const long slot = p.diag_shift / BN + 2 * (i0 / BM) + wm;
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
                cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16);
                cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
                if (g == 0) {
                    double* r = rpmir + slot * p.rows_pad + (jb - mir_row0) + nt * 8;
                    *reinterpret_cast<double2*>(r) = make_double2(cs0, cs1);
                }
            }
        }
    }
    {
        // both staging buffers free again before the next tile touches them
        if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
    if (p.rowpart) {
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            double v = sacc[mt];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (t == 0) rsb[wn * BM + wm * 64 + mt * 8 + g] = v;
        }
    }
}

// UPS = parked units handled per k-step (32 / UPS k-steps finish the parked tile; the launcher picks
// the smallest UPS in {1, 2, 4, 8} with UPS * kt >= 32, left-overs are handled after the main loop).
template <int KID, int UPS>
__global__ void __launch_bounds__(256, 2) gram_dmma_kernel(const GramParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* rs = smem + OFF_RS;                                    // [2 bufs][2 halves][BM]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + STAGES;
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(empty + STAGES);
    double* tstage = smem + OFF_TST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int KT = p.kt;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        fence_mbar_init();
    }
    if (warp == 0) {
        // 128 TMEM columns for the 4 MMA warps of this CTA (two CTAs per SM: 256 of 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 4 * BM; i += 256) rs[i] = 0.0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp >= 4) {
        // ---------------- producer warpgroup: hands its registers to the MMA warps ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp != 4) return;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            long i0, j0; int tx;
            tile_coords(p, tile, i0, j0, tx);
            int peer;
            if (tile_role(p, (int)(i0 / BM), tx, peer) == 0) continue;      // produced by a mirror store
            for (int kt = 0; kt < KT; ++kt, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);    // passes at once on first use
                double* As = smem + (size_t)s * STAGE_D;
                double* Bs = As + BK * AS;
                if (lane == 0) mbar_arrive_expect_tx(&full[s], BK * (BM + BN) * 8);
                __syncwarp();
                if (lane < BK)
                    bulk_g2s(As + lane * AS, p.A + (long)(kt * BK + lane) * p.lda + i0, BM * 8, &full[s]);
                else
                    bulk_g2s(Bs + (lane - BK) * BS, p.B + (long)(kt * BK + lane - BK) * p.ldb + j0, BN * 8,
                             &full[s]);
            }
        }
        return;
    }

    // -------------------------------- MMA warps --------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int wm = warp >> 1, wn = warp & 1;
    const uint32_t taddr = *tmem_base + ((uint32_t)(32 * warp) << 16);      // this warp's TMEM lane quarter
    double* tbuf = tstage + warp * 2 * TBUF_D;
    Parked pk;
    pk.valid = false; pk.mirror = false; pk.q = UNITS; pk.i0 = pk.j0 = 0; pk.tx = 0;
    pk.Kmir = nullptr; pk.rpmir = nullptr; pk.mir_row0 = 0; pk.cs0 = pk.cs1 = 0.0;
    int it = 0, fin = 0;

#define FX_GRAM_KSTEP()                                                                                          \
    {                                                                                                            \
        const int s = it % STAGES;                                                                               \
        mbar_wait(&full[s], (it / STAGES) & 1);                                                                  \
        const double* As = smem + (size_t)s * STAGE_D + wm * 64 + g;                                             \
        const double* Bs = smem + (size_t)s * STAGE_D + BK * AS + wn * 32 + g;                                   \
        _Pragma("unroll")                                                                                        \
        for (int kk = 0; kk < BK / 4; ++kk) {                                                                    \
            double a[8], b[4];                                                                                   \
            const int kr = kk * 4 + t;                                                                           \
            _Pragma("unroll")                                                                                    \
            for (int mt = 0; mt < 8; ++mt) a[mt] = As[kr * AS + mt * 8];                                         \
            _Pragma("unroll")                                                                                    \
            for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[kr * BS + nt * 8];                                         \
            _Pragma("unroll")                                                                                    \
            for (int mt = 0; mt < 8; ++mt)                                                                       \
                _Pragma("unroll")                                                                                \
                for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);               \
        }                                                                                                        \
    }
#define FX_GRAM_KSTEP_END()                                                                                      \
    {                                                                                                            \
        __syncwarp();                                                                                            \
        if (lane == 0) mbar_arrive(&empty[(it - 1) % STAGES]);          /* this warp is done with the stage */   \
    }

    for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        long i0, j0; int tx;
        tile_coords(p, tile, i0, j0, tx);
        int peer;
        const int role = tile_role(p, (int)(i0 / BM), tx, peer);
        if (role == 0) continue;
        const bool mirror = role == 2;
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        int kt = 0;
        if (pk.valid) {
            // k-steps that also take UPS units of the parked tile (same basic block as the DMMAs)
            if (pk.mirror) {
                for (; kt < KT && pk.q < UNITS; ++kt) {
                    FX_GRAM_KSTEP();
#pragma unroll
                    for (int j = 0; j < UPS; ++j) parked_unit<KID, true>(p, pk, pk.q + j, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
                    pk.q += UPS;
                    ++it;
                    FX_GRAM_KSTEP_END();
                }
            } else {
                for (; kt < KT && pk.q < UNITS; ++kt) {
                    FX_GRAM_KSTEP();
#pragma unroll
                    for (int j = 0; j < UPS; ++j) parked_unit<KID, false>(p, pk, pk.q + j, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
                    pk.q += UPS;
                    ++it;
                    FX_GRAM_KSTEP_END();
                }
            }
        }
        for (; kt < KT; ++kt) {
            FX_GRAM_KSTEP();
            ++it;
            FX_GRAM_KSTEP_END();
        }
        if (pk.valid) {
            // fewer k-steps than the parked tile needed (small d): the rest in one go
            for (; pk.q < UNITS; ++pk.q) {
                if (pk.mirror) parked_unit<KID, true>(p, pk, pk.q, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
                else           parked_unit<KID, false>(p, pk, pk.q, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
            }
            parked_finish(p, pk, lane, tid, rs + (fin & 1) * 2 * BM);
            ++fin;
        }

        const long ig0 = (p.diag_shift >= 0) ? i0 + p.diag_shift : -(1L << 40);
        const bool edge = (i0 + BM > p.rows_valid) || (j0 + BN > p.cols_valid) ||
                          (ig0 < j0 + BN && ig0 + BM > j0);
        double* Kmir = nullptr; double* rpmir = nullptr; long mir_row0 = 0;
        if (mirror) {
            Kmir = (p.sym == 2) ? p.peerK[peer] : p.K;
            rpmir = (p.sym == 2) ? p.peerRowpart[peer] : p.rowpart;
            mir_row0 = (p.sym == 2) ? (long)peer * p.rows_pad : p.diag_shift;
        }
        if (edge) {
            // ---- burst epilogue with masks: map in registers, shuffle row sums, direct stores ----
            double ni[8], nbv[4][2];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) ni[mt] = (KID == 0) ? -p.km.p0 * p.na[i0 + wm * 64 + mt * 8 + g] : 0.0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                nbv[nt][0] = nbv[nt][1] = 0.0;
                if (KID == 0) {
                    const double2 v = *reinterpret_cast<const double2*>(p.nb + j0 + wn * 32 + nt * 8 + 2 * t);
                    nbv[nt][0] = -p.km.p0 * v.x; nbv[nt][1] = -p.km.p0 * v.y;   // pre-scaled by -gamma (gram_map)
                }
            }
            double* rsb = rs + (fin & 1) * 2 * BM;
            if (mirror) gram_epilogue_mirror<KID, true>(p, acc, ni, i0, j0, wm, wn, g, t, rsb, Kmir, rpmir, mir_row0, tbuf);
            else        gram_epilogue_plain<KID, true>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
            if (p.rowpart) {
                bar_sync_compute();                  // both column halves of every row are in smem
                p.rowpart[(long)tx * p.rows_pad + i0 + tid] = rsb[tid] + rsb[BM + tid];
                rsb[tid] = 0.0;                      // parked tiles ACCUMULATE into these slots
                rsb[BM + tid] = 0.0;
            }
            ++fin;
        } else {
            // ---- park the raw dot products in TMEM; they are mapped during the next tile ----
            tmem_park(taddr, acc);
            pk.valid = true; pk.mirror = mirror; pk.i0 = i0; pk.j0 = j0; pk.tx = tx; pk.q = 0;
            pk.Kmir = Kmir; pk.rpmir = rpmir; pk.mir_row0 = mir_row0; pk.cs0 = pk.cs1 = 0.0;
        }
    }
    if (pk.valid) {
        // the last tile of this CTA has no successor to hide behind
        for (; pk.q < UNITS; ++pk.q) {
            if (pk.mirror) parked_unit<KID, true>(p, pk, pk.q, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
            else           parked_unit<KID, false>(p, pk, pk.q, taddr, wm, wn, g, t, lane, rs + (fin & 1) * 2 * BM, tbuf);
        }
        parked_finish(p, pk, lane, tid, rs + (fin & 1) * 2 * BM);
    }
#undef FX_GRAM_KSTEP
#undef FX_GRAM_KSTEP_END
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // mirrored tiles have left the SM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    bar_sync_compute();                                          // every MMA warp is done with its TMEM quarter
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(*tmem_base));
}

// ---- naive cross-check variant: one thread per output element -------------------------------
__global__ void gram_naive_kernel(const GramParams p, int d_pad) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long i = blockIdx.y;
    if (j >= p.ldk) return;
    double v = 0.0;
    if (i < p.rows_valid && j < p.cols_valid) {
        double dot = 0.0;
        for (int f = 0; f < d_pad; ++f) dot += p.A[(long)f * p.lda + i] * p.B[(long)f * p.ldb + j];
        const double ni = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.na[i] : 0.0;
        const double nj = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.nb[j] : 0.0;
        const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
        v = kernel_map(p.km, dot, ni, nj, ig == j);
    }
    p.K[i * p.ldk + j] = v;
}
// naive per-tile row sums so both variants feed the same reduction
__global__ void gram_naive_rowpart_kernel(const double* K, long ldk, double* rowpart, long rows_pad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long jt = blockIdx.y;
    if (i >= rows_pad) return;
    double s = 0.0;
    for (int c = 0; c < 64; ++c) s += K[i * ldk + jt * 64 + c];
    rowpart[jt * rows_pad + i] = s;
}

__global__ void rownorms_kernel(const double* P, long ld, long nobs, int d_pad, double* norms) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nobs) return;
    double s = 0.0;
    for (int f = 0; f < d_pad; ++f) {
        const double x = P[(long)f * ld + i];
        s = fma(x, x, s);
    }
    norms[i] = s;
}

bool use_v3() {
    static const bool v = []() { const char* e = getenv("FXB200_GRAM_V3"); return e && e[0] == '1'; }();
    return v;
}

template <int KID, int UPS>
int launch_k2(fx_ctx* ctx, const GramParams& p, unsigned grid) {
    FX_SET_SMEM_ONCE(ctx, (gram_dmma_kernel<KID, UPS>), SMEM_BYTES);
    gram_dmma_kernel<KID, UPS><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p);
    return 0;
}
template <int KID>
int launch_k(fx_ctx* ctx, const GramParams& p, unsigned grid) {
    const int kt = p.kt;
    if (kt >= 32) return launch_k2<KID, 1>(ctx, p, grid);
    if (kt >= 16) return launch_k2<KID, 2>(ctx, p, grid);
    if (kt >= 8) return launch_k2<KID, 4>(ctx, p, grid);
    return launch_k2<KID, 8>(ctx, p, grid);
}

}  // namespace

// slots of GramArgs::rowpart per row: one partial row sum per 64-column output tile
int64_t fx_gram_rowpart_slots(int64_t ldk) { return ldk / BN; }

int fx_launch_rownorms(fx_ctx* ctx, const double* P, int64_t ld, int64_t nobs, int64_t d_pad,
                       double* norms) {
    if (nobs <= 0) return 0;
    rownorms_kernel<<<(unsigned)ceil_div(nobs, 256), 256, 0, ctx->stream>>>(P, ld, nobs, (int)d_pad, norms);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_gram(fx_ctx* ctx, const GramArgs& a, int variant) {
    if (use_v3()) return fx_launch_gram_v3(ctx, a, variant);
    if (a.rows_pad % BM || a.ldk % BN || a.d_pad % BK || a.lda % 2 || a.ldb % 2) {
        fx_set_error("gram: bad padding rows_pad=%lld ldk=%lld d_pad=%lld", (long long)a.rows_pad,
                     (long long)a.ldk, (long long)a.d_pad);
        return FX_ERR_CUDA;
    }
    GramParams p;
    p.A = a.A; p.lda = a.lda; p.B = a.B; p.ldb = a.ldb; p.na = a.na; p.nb = a.nb;
    p.K = a.K; p.ldk = a.ldk; p.rowpart = a.rowpart; p.rows_pad = a.rows_pad;
    p.rows_valid = a.rows_valid; p.cols_valid = a.cols_valid; p.diag_shift = a.diag_shift;
    p.kt = (int)(a.d_pad / BK);
    p.tiles_x = (int)(a.ldk / BN);
    p.tiles_y = (int)(a.rows_pad / BM);
    p.tiles_total = (long)(a.ldk / BN) * (a.rows_pad / BM);
    // whole symmetric matrix on one rank: only tiles on / right of the diagonal block are computed
    p.sym = (a.sym_source && a.diag_shift >= 0 && a.diag_shift % BM == 0 &&
             a.diag_shift + a.rows_pad <= a.ldk && !a.no_symmetry) ? 1 : 0;
    p.rank = 0; p.world = 1;
    for (int c = 0; c < 16; ++c) { p.peerK[c] = nullptr; p.peerRowpart[c] = nullptr; }
    if (p.sym && a.peer_world > 1) {
        // cross-rank symmetry: needs every rank's block to be rows_pad wide and all peers mapped
        if (a.peer_world > 16 || a.ldk != (int64_t)a.peer_world * a.rows_pad || a.diag_shift % a.rows_pad ||
            a.peer_rank != (int)(a.diag_shift / a.rows_pad) || variant != 0) {
            fx_set_error("gram: inconsistent peer layout (world %d, rank %d)", a.peer_world, a.peer_rank);
            return FX_ERR_CUDA;
        }
        p.sym = 2; p.world = a.peer_world; p.rank = a.peer_rank;
        for (int c = 0; c < a.peer_world; ++c) { p.peerK[c] = a.peerK[c]; p.peerRowpart[c] = a.peerRowpart[c]; }
    }
    p.km.id = a.kernel_id; p.km.p0 = a.kp[0]; p.km.p1 = a.kp[1]; p.km.p2 = a.kp[2];
    double tmp[3];
    if (fx_make_kernel_map(a.kernel_id, a.kp, &p.km.ipow, tmp)) {
        fx_set_error("gram: invalid kernel id / parameters");
        return -1;
    }
    if (a.rows_pad == 0 || a.ldk == 0) return 0;
    if (variant == 0) {
        long want = 2L * (ctx->sm_count > 0 ? ctx->sm_count : 148);
        const unsigned grid = (unsigned)(p.tiles_total < want ? p.tiles_total : want);
        switch (kernel_variant(p.km.id, p.km.ipow)) {
            case 0: FX_TRY(launch_k<0>(ctx, p, grid)); break;
            case 1: FX_TRY(launch_k<1>(ctx, p, grid)); break;
            case 2: FX_TRY(launch_k<2>(ctx, p, grid)); break;
            case 3: FX_TRY(launch_k<3>(ctx, p, grid)); break;
            case 4: FX_TRY(launch_k<4>(ctx, p, grid)); break;
            case 5: FX_TRY(launch_k<5>(ctx, p, grid)); break;
            default: FX_TRY(launch_k<6>(ctx, p, grid)); break;
        }
        ctx->launches++;
    } else {
        dim3 grid((unsigned)ceil_div(a.ldk, 256), (unsigned)a.rows_pad);
        gram_naive_kernel<<<grid, 256, 0, ctx->stream>>>(p, (int)a.d_pad);
        ctx->launches++;
        if (a.rowpart) {
            dim3 g2((unsigned)ceil_div(a.rows_pad, 256), (unsigned)(a.ldk / 64));
            gram_naive_rowpart_kernel<<<g2, 256, 0, ctx->stream>>>(a.K, a.ldk, a.rowpart, a.rows_pad);
            ctx->launches++;
        }
    }
    FX_LAUNCH_CHECK();
    return 0;
}
