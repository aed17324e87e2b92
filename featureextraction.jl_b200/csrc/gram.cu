// gram.cu -- kernel-matrix construction: K = map(A B') with the kernel map, zero padding and
// per-tile row sums fused into the epilogue.  Replaces MLKernels.kernelmatrix as called at
// /root/reference/src/kpca.jl:51 (symmetric) and kpca.jl:65 (cross).
//
// Data path (persistent CTAs, 2 per SM, each looping over 128x64 output tiles):
//   producer warpgroup (one active warp; gives its registers to the MMA warps via setmaxnreg):
//     HBM --cp.async.bulk (TMA engine, UBLKCP) + mbarrier full/empty--> 3-stage smem ring of
//     16-feature slabs; it runs ahead across tile boundaries, so the next tile's first slabs
//     land while the MMA warps are still in the epilogue of the previous one
//   4 MMA warps: LDS.64 conflict-free (row stride = 4 mod 16 doubles) --> DMMA.8x8x4
//     (mma.sync f64), 64x32 register tile per warp, no CTA-wide barrier in the main loop
//   epilogue: kernel map applied IN REGISTERS (64 independent elements per thread, branch-free
//     exp) --> row sums by 2-step warp shuffle --> 16-byte stores straight from the accumulator
//     fragments (one store instruction covers 8 rows x 64 contiguous bytes = 16 full sectors).
//
// Symmetry.  K is symmetric when A is a row block of B.  sym = 1 (one rank, or no peer access):
// inside this rank's square diagonal block only the tiles on / right of the diagonal are computed
// and the strictly-right ones are mirrored by transposed stores from the fragments.  sym = 2
// (several ranks, peer arenas mapped through CUDA IPC): the off-diagonal blocks are shared out
// too -- of each pair of blocks (r,c) / (c,r) exactly one rank computes every 128x128 super-tile
// and stores the transposed values, and the column sums that are the peer's row-sum partials,
// STRAIGHT INTO THE PEER'S K over NVLink (plain st.global to the IPC-mapped address), so every
// rank executes ~half of its row block and no separate exchange step exists.
#include "fx_kernels.cuh"
#include "gram_tiles.cuh"

namespace {

constexpr int BM = FX_GRAM_BM, BN = FX_GRAM_BN, BK = 16, STAGES = 3;
constexpr int AS = BM + 4;                      // 132: (t*AS + g) mod 16 distinct for t,g in 0..3
constexpr int BS = BN + 4;                      // 68
constexpr int STAGE_D = BK * AS + BK * BS;      // doubles per stage (3200 -> 25600 B)
// mirror staging: per MMA warp two buffers of 8 transposed rows x 64 (+2 pad: conflict-free
// 8-byte stores for lanes (g,t) -> word (2t+q)*TS + g) -- sent off with 512-byte bulk stores
constexpr int TS = 66, TBUF_D = 8 * TS;
constexpr int SMEM_BYTES = STAGES * STAGE_D * 8 + 2 * 2 * BM * 8 + 128 + 4 * 2 * TBUF_D * 8;

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
    unsigned zero;                    // 0 at run time, opaque to the compiler (see the release of a stage)
    int rank, world;                  // sym == 2: this rank's block index and the number of blocks
    double* peerK[16];                // sym == 2: K of rank c (own entry = K)
    double* peerRowpart[16];          //           rowpart of rank c (or null)
    KernelMap km;
};

// tile_role() / tile_coords(): gram_tiles.cuh (host-testable)

__device__ __forceinline__ void bar_sync_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Map of one element; for the Gaussian kernel (KID 0) the norms arrive pre-scaled by -gamma.
template <int KID>
__device__ __forceinline__ double gram_map(const KernelMap& km, double dot, double ni, double nj, bool same) {
    if (KID == 0) return gaussian_map_prescaled(2.0 * km.p0, dot, ni, nj, same);
    return kernel_map_t<KID>(km, dot, ni, nj, same);
}

// Kernel map + row sums + stores for one warp's 64x32 accumulator tile.  EDGE tiles (ragged
// borders or touching the diagonal) carry the validity masks and the same-point test; interior
// tiles run the bare arithmetic.
template <int KID, bool EDGE>
__device__ __forceinline__ void gram_epilogue_plain(const GramParams& p, double (&acc)[8][4][2],
                                                    const double (&ni)[8], const double (&nbv)[4][2], long i0,
                                                    long j0, int wm, int wn, int g, int t, double* rsb) {
    const long jb = j0 + wn * 32 + 2 * t;
    if (KID == 0 && !EDGE) {
        // interior Gaussian tiles: the 8 elements of a fragment row go through the staged exp together
        // (8 interleaved dependency chains per warp instead of one after the other)
        const double two_gamma = 2.0 * p.km.p0;
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            double* krow = p.K + (i0 + wm * 64 + mt * 8 + g) * p.ldk + jb;
            double x[8];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                x[2 * nt] = fma(two_gamma, acc[mt][nt][0], ni[mt] + nbv[nt][0]);
                x[2 * nt + 1] = fma(two_gamma, acc[mt][nt][1], ni[mt] + nbv[nt][1]);
            }
            fx_exp_nonpos_vec<8>(x);
            double sacc = 0.0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<double2*>(krow + nt * 8) = make_double2(x[2 * nt], x[2 * nt + 1]);
                sacc += x[2 * nt] + x[2 * nt + 1];
            }
            if (p.rowpart) {
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
                if (t == 0) rsb[wn * BM + wm * 64 + mt * 8 + g] = sacc;
            }
        }
        return;
    }
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
        if (KID == 0 && !EDGE) {
            // interior Gaussian tiles: 2 x 8 elements through the staged exp (interleaved chains)
            const double two_gamma = 2.0 * p.km.p0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double x[8];
#pragma unroll
                for (int m4 = 0; m4 < 4; ++m4) {
                    const int mt = half * 4 + m4;
                    x[2 * m4] = fma(two_gamma, acc[mt][nt][0], ni[mt] + nb0);
                    x[2 * m4 + 1] = fma(two_gamma, acc[mt][nt][1], ni[mt] + nb1);
                }
                fx_exp_nonpos_vec<8>(x);
#pragma unroll
                for (int m4 = 0; m4 < 4; ++m4) {
                    const int mt = half * 4 + m4;
                    const long i = i0 + wm * 64 + mt * 8 + g;
                    const double v0 = x[2 * m4], v1 = x[2 * m4 + 1];
                    *reinterpret_cast<double2*>(p.K + i * p.ldk + jb + nt * 8) = make_double2(v0, v1);
                    tb[(2 * t) * TS + mt * 8 + g] = v0;              // transposed: row = column of the tile
                    tb[(2 * t + 1) * TS + mt * 8 + g] = v1;
                    cs0 += v0;
                    cs1 += v1;
                    sacc[mt] += v0 + v1;
                }
            }
        } else {
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

template <int KID>
__global__ void __launch_bounds__(256, 2) gram_dmma_kernel(const GramParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* rs = smem + (size_t)STAGES * STAGE_D;                  // [2 bufs][2 halves][BM]
    uint64_t* full = reinterpret_cast<uint64_t*>(rs + 4 * BM);
    uint64_t* empty = full + STAGES;
    double* tstage = rs + 4 * BM + 16;                             // after the 128 bytes of mbarriers

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int KT = p.kt;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        fence_mbar_init();
    }
    __syncthreads();

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
    int it = 0, tcount = 0;
    for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        long i0, j0; int tx;
        tile_coords(p, tile, i0, j0, tx);
        int peer;
        const int role = tile_role(p, (int)(i0 / BM), tx, peer);
        if (role == 0) continue;
        const bool mirror = role == 2;
        ++tcount;
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        // squared norms for the Gaussian map: requested now, consumed after the main loop
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

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const double* As = smem + (size_t)s * STAGE_D + wm * 64 + g;
            const double* Bs = smem + (size_t)s * STAGE_D + BK * AS + wn * 32 + g;
            unsigned dep = 0;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double a[8], b[4];
                const int kr = kk * 4 + t;
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) a[mt] = As[kr * AS + mt * 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[kr * BS + nt * 8];
                if (kk == BK / 4 - 1 && p.sym == 2) {
                    // see the release below: the barrier address is made to depend on the last fragments
#pragma unroll
                    for (int mt = 0; mt < 8; ++mt) dep |= (unsigned)__double2hiint(a[mt]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dep |= (unsigned)__double2hiint(b[nt]);
                    dep &= p.zero;
                }
#pragma unroll
                for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
            }
            // Release of stage s.  ptxas schedules the arrive right behind the last LDS of the stage, ahead
            // of the DMMAs that consume them.  With the load/store queue backed up by REMOTE stores (mirror
            // epilogues of the cross-rank symmetric schedule, sym == 2) the arrive was observed to overtake
            // those loads: the producer refilled the stage and one warp read the NEXT slab for its last
            // 8-column fragment -- one 64 x 8 sub-block of K off, about 1 fit in 50 on four GPUs
            // (profiles/r02_peer_sym_investigation.md).  In that mode the barrier address DEPENDS on the
            // last fragments (p.zero is 0, but only at run time), so the arrive cannot issue before those
            // loads have returned: 0 events in 500 fits.  Without remote stores the effect was never
            // seen (every single-GPU bench line compares two fits bit for bit) and the dependency costs
            // 2-6 % of the kernel, so the other modes release as before.
            __syncwarp();
            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(&empty[s]) + dep));
        }

        // ---- epilogue: map in registers, shuffle row sums, direct 16-byte stores ----
        double* rsb = rs + (tcount & 1) * 2 * BM;
        const long ig0 = (p.diag_shift >= 0) ? i0 + p.diag_shift : -(1L << 40);
        const bool edge = (i0 + BM > p.rows_valid) || (j0 + BN > p.cols_valid) ||
                          (ig0 < j0 + BN && ig0 + BM > j0);
        if (mirror) {
            double* Kmir = (p.sym == 2) ? p.peerK[peer] : p.K;
            double* rpmir = (p.sym == 2) ? p.peerRowpart[peer] : p.rowpart;
            const long mir_row0 = (p.sym == 2) ? (long)peer * p.rows_pad : p.diag_shift;
            if (edge) gram_epilogue_mirror<KID, true>(p, acc, ni, i0, j0, wm, wn, g, t, rsb, Kmir, rpmir, mir_row0, tstage + warp * 2 * TBUF_D);
            else      gram_epilogue_mirror<KID, false>(p, acc, ni, i0, j0, wm, wn, g, t, rsb, Kmir, rpmir, mir_row0, tstage + warp * 2 * TBUF_D);
        } else {
            if (edge) gram_epilogue_plain<KID, true>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
            else      gram_epilogue_plain<KID, false>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
        }
        if (p.rowpart) {
            bar_sync_compute();                  // both column halves of every row are in smem
            p.rowpart[(long)tx * p.rows_pad + i0 + tid] = rsb[tid] + rsb[BM + tid];
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // mirrored tiles have left the SM
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
    p.rank = 0; p.world = 1; p.zero = 0;
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
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<0>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<1>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<2>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<3>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<4>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<5>, SMEM_BYTES);
        FX_SET_SMEM_ONCE(ctx, gram_dmma_kernel<6>, SMEM_BYTES);
        long want = 2L * (ctx->sm_count > 0 ? ctx->sm_count : 148);
        const unsigned grid = (unsigned)(p.tiles_total < want ? p.tiles_total : want);
        switch (kernel_variant(p.km.id, p.km.ipow)) {
            case 0: gram_dmma_kernel<0><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            case 1: gram_dmma_kernel<1><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            case 2: gram_dmma_kernel<2><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            case 3: gram_dmma_kernel<3><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            case 4: gram_dmma_kernel<4><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            case 5: gram_dmma_kernel<5><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
            default: gram_dmma_kernel<6><<<grid, 256, SMEM_BYTES, ctx->stream>>>(p); break;
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
