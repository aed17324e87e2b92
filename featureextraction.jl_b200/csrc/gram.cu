// gram.cu -- kernel-matrix construction: K = map(A B') with the kernel map, zero padding and
// per-tile row sums fused into the epilogue.  Replaces MLKernels.kernelmatrix as called at
// /root/reference/src/kpca.jl:51 (symmetric) and kpca.jl:65 (cross).
//
// Data path (persistent CTAs, 2 per SM, each looping over 128x64 output tiles):
//   producer warpgroup (one active warp; gives its registers to the MMA warps via setmaxnreg):
//     HBM --cp.async.bulk (TMA engine, UBLKCP) + mbarrier full/empty--> 4-stage smem ring of
//     16-feature slabs; it runs ahead across tile boundaries, so the next tile's first slabs
//     land while the MMA warps are still in the epilogue of the previous one
//   4 MMA warps: LDS.64 conflict-free (row stride = 4 mod 16 doubles) --> DMMA.8x8x4
//     (mma.sync f64), 64x32 register tile per warp, no CTA-wide barrier in the main loop
//   epilogue: kernel map applied IN REGISTERS (64 independent elements per thread, branch-free
//     exp) --> row sums by 2-step warp shuffle --> 16-byte stores straight from the accumulator
//     fragments (one store instruction covers 8 rows x 64 contiguous bytes = 16 full sectors).
#include "fx_kernels.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 4;
constexpr int AS = BM + 4;                      // 132: (t*AS + g) mod 16 distinct for t,g in 0..3
constexpr int BS = BN + 4;                      // 68
constexpr int STAGE_D = BK * AS + BK * BS;      // doubles per stage (3200 -> 25600 B)
constexpr int SMEM_BYTES = STAGES * STAGE_D * 8 + 2 * 2 * BM * 8 + 128;

struct GramParams {
    const double* A; long lda;
    const double* B; long ldb;
    const double* na; const double* nb;
    double* K; long ldk;
    double* rowpart; long rows_pad;
    long rows_valid, cols_valid, diag_shift;
    int kt;
    int tiles_x, tiles_y; long tiles_total;
    int sym;                          // 1: symmetric source (A == B): inside this rank's square diagonal
                                      //    block compute only the tiles on / right of the diagonal and
                                      //    mirror the strictly-right ones; other column blocks in full
    KernelMap km;
};

// Tile order: column panels of PANEL tile-columns swept over all row blocks, so that the panel's
// slab of the column operand (PANEL x 131 KB at d=256) stays L2-resident while the row operand
// streams; a plain row-major sweep re-reads all of X from HBM once per row block.
constexpr int PANEL = 128;
__device__ __forceinline__ void tile_coords(const GramParams& p, long tile, long& i0, long& j0, int& tx) {
    const long per_panel = (long)PANEL * p.tiles_y;
    const int panel = (int)(tile / per_panel);
    const long within = tile - (long)panel * per_panel;
    const int pw = min(PANEL, p.tiles_x - panel * PANEL);      // last panel may be narrower
    const int ty = (int)(within / pw);
    tx = panel * PANEL + (int)(within - (long)ty * pw);
    i0 = (long)ty * BM;
    j0 = (long)tx * BN;
}

__device__ __forceinline__ void bar_sync_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Kernel map + row sums + stores for one warp's 64x32 accumulator tile.  EDGE tiles (ragged
// borders or touching the diagonal) carry the validity masks and the same-point test; interior
// tiles run the bare arithmetic.  MIRROR tiles (symmetric mode, strictly right of the diagonal
// block) also write the transposed element K[j][i] (8 lanes x 8 B = 64 contiguous bytes per
// column) and the column sums, which are the row-sum partials of the mirrored rows.
template <int KID, bool EDGE, bool MIRROR>
__device__ __forceinline__ void gram_epilogue(const GramParams& p, double (&acc)[8][4][2],
                                              const double (&ni)[8], const double (&nbv)[4][2], long i0,
                                              long j0, int wm, int wn, int g, int t, double* rsb) {
    const long jb = j0 + wn * 32 + 2 * t;
    double csum[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) csum[nt][0] = csum[nt][1] = 0.0;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
        const long i = i0 + wm * 64 + mt * 8 + g;
        double* krow = p.K + i * p.ldk + jb;
        double* kcol = p.K + (jb - p.diag_shift) * p.ldk + i + p.diag_shift;   // transposed position of (i, jb)
        const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
        const bool iv = i < p.rows_valid;
        double sacc = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            double v0, v1;
            if (EDGE) {
                const long j = jb + nt * 8;
                v0 = kernel_map_t<KID>(p.km, acc[mt][nt][0], ni[mt], nbv[nt][0], ig == j);
                v1 = kernel_map_t<KID>(p.km, acc[mt][nt][1], ni[mt], nbv[nt][1], ig == j + 1);
                v0 = (iv && j < p.cols_valid) ? v0 : 0.0;
                v1 = (iv && j + 1 < p.cols_valid) ? v1 : 0.0;
            } else {
                v0 = kernel_map_t<KID>(p.km, acc[mt][nt][0], ni[mt], nbv[nt][0], false);
                v1 = kernel_map_t<KID>(p.km, acc[mt][nt][1], ni[mt], nbv[nt][1], false);
            }
            *reinterpret_cast<double2*>(krow + nt * 8) = make_double2(v0, v1);
            if (MIRROR) {
                kcol[(long)(nt * 8) * p.ldk] = v0;
                kcol[(long)(nt * 8 + 1) * p.ldk] = v1;
                csum[nt][0] += v0;
                csum[nt][1] += v1;
            }
            sacc += v0 + v1;
        }
        if (p.rowpart) {
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (t == 0) rsb[wn * BM + wm * 64 + mt * 8 + g] = sacc;
        }
    }
    if (MIRROR && p.rowpart) {
        // column sums over this warp's 64 rows = row-sum partial of mirrored row j for the 64-column
        // slot [i0 + wm*64, +64) of that row: slot index = 2*(i0/BM) + wm
        const long slot = p.diag_shift / BN + 2 * (i0 / BM) + wm;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double v = csum[nt][q];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0) p.rowpart[slot * p.rows_pad + (jb - p.diag_shift) + nt * 8 + q] = v;
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
            if (p.sym) {                                          // left of the diagonal inside the
                const int rel = tx - (int)(p.diag_shift / BN);    // diagonal block: produced by a mirror store
                if (rel >= 0 && rel < 2 * (int)(i0 / BM)) continue;
            }
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
        const int rel = tx - (int)(p.diag_shift / BN);            // column tile relative to the diagonal block
        if (p.sym && rel >= 0 && rel < 2 * (int)(i0 / BM)) continue;
        const bool mirror = p.sym && rel >= 2 * (int)(i0 / BM) + 2 && rel < 2 * p.tiles_y;
        ++tcount;
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        // squared norms for the Gaussian map: requested now, consumed after the main loop
        double ni[8], nbv[4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) ni[mt] = (KID == 0) ? p.na[i0 + wm * 64 + mt * 8 + g] : 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            nbv[nt][0] = nbv[nt][1] = 0.0;
            if (KID == 0) {
                const double2 v = *reinterpret_cast<const double2*>(p.nb + j0 + wn * 32 + nt * 8 + 2 * t);
                nbv[nt][0] = v.x; nbv[nt][1] = v.y;
            }
        }

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const double* As = smem + (size_t)s * STAGE_D + wm * 64 + g;
            const double* Bs = smem + (size_t)s * STAGE_D + BK * AS + wn * 32 + g;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double a[8], b[4];
                const int kr = kk * 4 + t;
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) a[mt] = As[kr * AS + mt * 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[kr * BS + nt * 8];
#pragma unroll
                for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);               // this warp is done with stage s
        }

        // ---- epilogue: map in registers, shuffle row sums, direct 16-byte stores ----
        double* rsb = rs + (tcount & 1) * 2 * BM;
        const long ig0 = (p.diag_shift >= 0) ? i0 + p.diag_shift : -(1L << 40);
        const bool edge = (i0 + BM > p.rows_valid) || (j0 + BN > p.cols_valid) ||
                          (ig0 < j0 + BN && ig0 + BM > j0);
        if (mirror) {
            if (edge) gram_epilogue<KID, true, true>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
            else      gram_epilogue<KID, false, true>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
        } else {
            if (edge) gram_epilogue<KID, true, false>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
            else      gram_epilogue<KID, false, false>(p, acc, ni, nbv, i0, j0, wm, wn, g, t, rsb);
        }
        if (p.rowpart) {
            bar_sync_compute();                  // both column halves of every row are in smem
            p.rowpart[(long)tx * p.rows_pad + i0 + tid] = rsb[tid] + rsb[BM + tid];
        }
    }
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
    p.km.id = a.kernel_id; p.km.p0 = a.kp[0]; p.km.p1 = a.kp[1]; p.km.p2 = a.kp[2];
    double tmp[3];
    if (fx_make_kernel_map(a.kernel_id, a.kp, &p.km.ipow, tmp)) {
        fx_set_error("gram: invalid kernel id / parameters");
        return -1;
    }
    if (a.rows_pad == 0 || a.ldk == 0) return 0;
    if (variant == 0) {
        static bool attr_set = false;
        if (!attr_set) {
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            FX_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set = true;
        }
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
