// symm.cu -- Y = K V for the block eigensolver: K is this rank's row block of the symmetric
// kernel matrix (row-major, streamed from HBM exactly once per pass), V the all-gathered
// n x b block.  This is the only heavy operation of the top-k eigensolve that replaces the
// dense dsyevd_ call of /root/reference/src/common.jl:128-131 (via components_eig!, :161-165).
//
// Work unit = (128-row tile, k-split).  One producer warp drives the TMA engine with tensor
// maps (cp.async.bulk.tensor, SASS UTMALDG): per stage ONE 3-D box {16 x 128 rows x 4} of K
// (64 KB, L2 evict-first: streamed once) and ONE 4-D box of V (L2 evict-last: reused by every
// row tile), 128-byte swizzle, completion on mbarrier full/empty pairs -- no CTA-wide barrier in
// the main loop.  8 MMA warps each own 16 rows x b columns of the output and feed DMMA.8x8x4
// straight from LDS.128 fragments: thread (g,t) owns k = 16 s + 4 t + e for the e-th MMA of a
// 16-wide step, for A and B alike, so one 32-byte load serves four MMAs, and with the swizzle
// chunk' = chunk ^ (row & 7) every quarter-warp hits 8 distinct 16-byte banks.
// K-split partials are reduced by a second tiny kernel in a fixed order (deterministic).
//
// CENTER variant (first pass of a fit): K still holds the uncentred kernel matrix.  The MMA warps
// apply K~ = ((K - rowmean) - colmean) + grand to their A fragments in registers (the 64 column
// means of the chunk ride along in the stage as a third bulk copy), put the centred values back
// into the stage (every element of the tile is owned by exactly one thread, so this is race
// free) and feed the same registers to the DMMAs.  A tenth warp sends each finished stage back
// to K with ONE tensor-map bulk store (cp.async.bulk.tensor ... global.shared::cta, the swizzle
// undone by the TMA engine) and frees the stage once the engine has read it -- register-side
// STG stores of the 8-row fragments choked the LSU queue (lg_throttle; profiles/).  So the
// centring pass of MLKernels.centerkernelmatrix! (/root/reference/src/kpca.jl:52) and the first
// K~ V product share one read of K.  Same arithmetic, element for element, as center.cu.
#include <cuda.h>
#include "fx_kernels.cuh"

namespace {

constexpr int RT = 128;            // rows per tile
constexpr int BKC = 64;            // k-chunk (columns of K per stage)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

struct SymmParams {
    long rows_pad; long ldr;
    double* Yp;                    // [splits][bcols][rows_pad]
    int nchunks, splits, bcols, nstage;
    // CENTER variant only
    const double* colmean;         // [ldk] column (== row) means, index = global sample
    const double* grand;           // device scalar
    long rows_valid, cols_valid, row0;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)),
        "l"(hint)
        : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}

template <int NB, bool CENTER>
__global__ void __launch_bounds__(CENTER ? 320 : 288, 1)
symm_dmma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const SymmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment so that (address bits 7..9) == (row & 7) for the 128-byte swizzle
    // (pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST)
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int K_BYTES = RT * BKC * 8;                 // 65536
    constexpr int V_BYTES = NB * 8 * BKC * 8;             // NB * 4096
    constexpr int C_BYTES = CENTER ? 1024 : 0;            // 64 column means (512 B) + pad
    constexpr int STAGE_BYTES = K_BYTES + V_BYTES + C_BYTES;   // multiple of 1024
    const int NS = p.nstage;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)NS * STAGE_BYTES);
    uint64_t* empty = full + 4;
    uint64_t* done = full + 8;                           // CENTER: stage centred in smem, ready to store

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const long rt = blockIdx.x / p.splits;
    const int ks = blockIdx.x % p.splits;
    const int c_begin = (int)((long)p.nchunks * ks / p.splits);
    const int c_end = (int)((long)p.nchunks * (ks + 1) / p.splits);
    const int nc = c_end - c_begin;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CENTER ? 1 : 8);        // CENTER: the store warp frees the stage
            mbar_init(&done[s], 8);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 8) {
        // ------------------------------ producer warp (one elected lane) ------------------------------
        if (lane == 0) {
            const int slabs_per_rank = (int)(p.ldr / 16);
            for (int ci = 0; ci < nc; ++ci) {
                const int s = ci % NS;
                mbar_wait(&empty[s], ((ci / NS) & 1) ^ 1);
                unsigned char* st = base + (size_t)s * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], K_BYTES + V_BYTES + (CENTER ? BKC * 8 : 0));
                const int slab = (c_begin + ci) * (BKC / 16);             // global 16-column slab index
                tma_load_3d(st, &tmK, 0, (int)(rt * RT), slab, &full[s], L2_EVICT_FIRST);
                const int rk = slab / slabs_per_rank;
                tma_load_4d(st + K_BYTES, &tmV, 0, 0, slab - rk * slabs_per_rank, rk, &full[s], L2_EVICT_LAST);
                if (CENTER)
                    bulk_g2s(st + K_BYTES + V_BYTES, p.colmean + (long)(c_begin + ci) * BKC, BKC * 8, &full[s]);
            }
        }
        return;
    }
    if (CENTER && warp == 9) {
        // ------------------------------ store warp (one elected lane) ------------------------------
        if (lane == 0) {
            for (int ci = 0; ci < nc; ++ci) {
                const int s = ci % NS;
                mbar_wait(&done[s], (ci / NS) & 1);
                tma_store_3d(&tmK, base + (size_t)s * STAGE_BYTES, 0, (int)(rt * RT), (c_begin + ci) * (BKC / 16));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(&empty[s]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }

    // -------------------------------- MMA warps --------------------------------
    double acc[2][NB][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NB; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // per-thread swizzled offsets (doubles) inside a 16-column slab row: chunks 2t and 2t+1
    const int off0 = ((2 * t) ^ g) * 2, off1 = ((2 * t + 1) ^ g) * 2;
    // CENTER: row means of this thread's two rows, the grand mean, and whether the row tile is
    // free of padding rows (then only chunks that cross cols_valid need per-element predicates)
    double rmean[2] = {0.0, 0.0}, gmean = 0.0;
    bool rows_full = true;
    if (CENTER) {
        gmean = *p.grand;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) rmean[mt] = p.colmean[p.row0 + rt * RT + warp * 16 + mt * 8 + g];
        rows_full = (rt + 1) * RT <= p.rows_valid;
    }
    for (int ci = 0; ci < nc; ++ci) {
        const int s = ci % NS;
        mbar_wait(&full[s], (ci / NS) & 1);
        double* Ks = reinterpret_cast<double*>(base + (size_t)s * STAGE_BYTES);
        const double* Vs = Ks + RT * BKC;
        const double* Cs = Vs + NB * 8 * BKC;
        const long colbase = (long)(c_begin + ci) * BKC;
        const bool fast = rows_full && (colbase + BKC <= p.cols_valid);
#pragma unroll
        for (int s16 = 0; s16 < BKC / 16; ++s16) {
            double a[2][4], b[NB][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const double* row = Ks + s16 * (RT * 16) + (warp * 16 + mt * 8 + g) * 16;
                const double2 lo = *reinterpret_cast<const double2*>(row + off0);
                const double2 hi = *reinterpret_cast<const double2*>(row + off1);
                a[mt][0] = lo.x; a[mt][1] = lo.y; a[mt][2] = hi.x; a[mt][3] = hi.y;
            }
            if (CENTER) {
                const double2 c01 = *reinterpret_cast<const double2*>(Cs + s16 * 16 + 4 * t);
                const double2 c23 = *reinterpret_cast<const double2*>(Cs + s16 * 16 + 4 * t + 2);
                const double cm[4] = {c01.x, c01.y, c23.x, c23.y};
                const long col = colbase + s16 * 16 + 4 * t;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const long lrow = rt * RT + warp * 16 + mt * 8 + g;
                    if (fast) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) a[mt][e] = ((a[mt][e] - rmean[mt]) - cm[e]) + gmean;
                    } else {
                        const bool rv = lrow < p.rows_valid;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (rv && col + e < p.cols_valid) a[mt][e] = ((a[mt][e] - rmean[mt]) - cm[e]) + gmean;
                    }
                    double* row = Ks + s16 * (RT * 16) + (warp * 16 + mt * 8 + g) * 16;
                    *reinterpret_cast<double2*>(row + off0) = make_double2(a[mt][0], a[mt][1]);
                    *reinterpret_cast<double2*>(row + off1) = make_double2(a[mt][2], a[mt][3]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NB; ++nt) {
                const double* row = Vs + s16 * (NB * 8 * 16) + (nt * 8 + g) * 16;
                const double2 lo = *reinterpret_cast<const double2*>(row + off0);
                const double2 hi = *reinterpret_cast<const double2*>(row + off1);
                b[nt][0] = lo.x; b[nt][1] = lo.y; b[nt][2] = hi.x; b[nt][3] = hi.y;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NB; ++nt)
                        dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt][e], b[nt][e]);
        }
        if (CENTER) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes -> TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(CENTER ? &done[s] : &empty[s]);
    }

    double* Y = p.Yp + (long)ks * p.bcols * p.rows_pad;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NB; ++nt) {
            const long row = rt * RT + warp * 16 + mt * 8 + g;
            const int col = nt * 8 + 2 * t;
            Y[(long)col * p.rows_pad + row] = acc[mt][nt][0];
            Y[(long)(col + 1) * p.rows_pad + row] = acc[mt][nt][1];
        }
}

__global__ void symm_reduce_kernel(const double* __restrict__ Yp, int splits, long total,
                                   double* __restrict__ Y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += Yp[(long)k * total + i];
    Y[i] = s;
}

// naive cross-check: one thread per (row, col)
__global__ void symm_naive_kernel(const double* K, long ldk, const double* V, long ldr, long v_rank_stride,
                                  long rows_pad, long ncols, double* Y) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (row >= rows_pad) return;
    double s = 0.0;
    for (long j = 0; j < ncols; ++j) {
        const long rk = j / ldr;
        s += K[row * ldk + j] * V[rk * v_rank_stride + (long)c * ldr + (j - rk * ldr)];
    }
    Y[(long)c * rows_pad + row] = s;
}

template <int NB, bool CENTER>
int launch_nb(fx_ctx* ctx, const CUtensorMap& tmK, const CUtensorMap& tmV, SymmParams p) {
    constexpr int STAGE_BYTES = RT * BKC * 8 + NB * 8 * BKC * 8 + (CENTER ? 1024 : 0);
    int ns = (227 * 1024 - 1024 - 128) / STAGE_BYTES;
    if (ns > 4) ns = 4;
    p.nstage = ns;
    const int smem = ns * STAGE_BYTES + 1024 + 128;
    FX_SET_SMEM_ONCE(ctx, (symm_dmma_kernel<NB, CENTER>), smem);
    const unsigned grid = (unsigned)((p.rows_pad / RT) * p.splits);
    symm_dmma_kernel<NB, CENTER><<<grid, CENTER ? 320 : 288, smem, ctx->stream>>>(tmK, tmV, p);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

template <int NB>
int launch_nb_sel(fx_ctx* ctx, const CUtensorMap& tmK, const CUtensorMap& tmV, const SymmParams& p, bool center) {
    return center ? launch_nb<NB, true>(ctx, tmK, tmV, p) : launch_nb<NB, false>(ctx, tmK, tmV, p);
}

}  // namespace

// Pick the k-split count that fills the SMs in whole waves (1 CTA per SM).
int fx_symm_splits(fx_ctx* ctx, int64_t rows_pad, int64_t ldk) {
    const int64_t tiles = rows_pad / RT, nchunks = ldk / BKC;
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= 16 && s <= nchunks; ++s) {
        const int64_t units = tiles * s;
        const int64_t waves = ceil_div(units, sms);
        double eff = (double)units / (double)(waves * sms);
        if (nchunks / s < 8) eff *= 0.8;               // keep per-unit streams long
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

int fx_launch_symm(fx_ctx* ctx, const SymmArgs& a, int variant) {
    if (a.rows_pad % RT || a.ldk % BKC || a.bcols % 8 || a.bcols < 8 || a.bcols > 64 || a.ldr % BKC) {
        fx_set_error("symm: bad shapes rows_pad=%lld ldk=%lld bcols=%d ldr=%lld", (long long)a.rows_pad,
                     (long long)a.ldk, a.bcols, (long long)a.ldr);
        return FX_ERR_CUDA;
    }
    if (a.rows_pad == 0) return 0;
    const bool center = a.center_colmean != nullptr;
    if (center && variant != 0) {
        fx_set_error("symm: the fused centring pass exists only in the DMMA variant");
        return FX_ERR_CUDA;
    }
    if (variant != 0) {
        dim3 grid((unsigned)ceil_div(a.rows_pad, 128), (unsigned)a.bcols);
        symm_naive_kernel<<<grid, 128, 0, ctx->stream>>>(a.K, a.ldk, a.V, a.ldr, a.v_rank_stride, a.rows_pad,
                                                         a.ldk, a.Y);
        ctx->launches++;
        FX_LAUNCH_CHECK();
        return 0;
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        fx_set_error("symm: cuTensorMapEncodeTiled is not available from this driver");
        return FX_ERR_CUDA;
    }
    CUtensorMap tmK, tmV;
    {
        cuuint64_t dims[3] = {16, (cuuint64_t)a.rows_pad, (cuuint64_t)(a.ldk / 16)};
        cuuint64_t strides[2] = {(cuuint64_t)a.ldk * 8, 128};
        cuuint32_t box[3] = {16, RT, BKC / 16};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)a.K, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fx_set_error("symm: tensor map for K rejected (CUresult %d)", (int)r); return FX_ERR_CUDA; }
    }
    {
        const int64_t nranks = a.ldk / a.ldr;
        cuuint64_t dims[4] = {16, (cuuint64_t)a.bcols, (cuuint64_t)(a.ldr / 16), (cuuint64_t)nranks};
        cuuint64_t strides[3] = {(cuuint64_t)a.ldr * 8, 128, (cuuint64_t)a.v_rank_stride * 8};
        cuuint32_t box[4] = {16, (cuuint32_t)a.bcols, BKC / 16, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)a.V, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fx_set_error("symm: tensor map for V rejected (CUresult %d)", (int)r); return FX_ERR_CUDA; }
    }
    SymmParams p;
    p.rows_pad = a.rows_pad; p.ldr = a.ldr; p.nchunks = (int)(a.ldk / BKC); p.bcols = a.bcols;
    p.splits = a.splits < 1 ? 1 : a.splits;
    p.Yp = (p.splits == 1) ? a.Y : a.Ypart;
    p.nstage = 2;
    p.colmean = a.center_colmean; p.grand = a.center_grand;
    p.rows_valid = a.rows_valid; p.cols_valid = a.cols_valid; p.row0 = a.row0;
    switch (a.bcols / 8) {
        case 1: FX_TRY(launch_nb_sel<1>(ctx, tmK, tmV, p, center)); break;
        case 2: FX_TRY(launch_nb_sel<2>(ctx, tmK, tmV, p, center)); break;
        case 3: FX_TRY(launch_nb_sel<3>(ctx, tmK, tmV, p, center)); break;
        case 4: FX_TRY(launch_nb_sel<4>(ctx, tmK, tmV, p, center)); break;
        case 5: FX_TRY(launch_nb_sel<5>(ctx, tmK, tmV, p, center)); break;
        case 6: FX_TRY(launch_nb_sel<6>(ctx, tmK, tmV, p, center)); break;
        case 7: FX_TRY(launch_nb_sel<7>(ctx, tmK, tmV, p, center)); break;
        default: FX_TRY(launch_nb_sel<8>(ctx, tmK, tmV, p, center)); break;
    }
    if (p.splits > 1) {
        const long total = (long)a.bcols * a.rows_pad;
        symm_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, ctx->stream>>>(a.Ypart, p.splits, total, a.Y);
        ctx->launches++;
        FX_LAUNCH_CHECK();
    }
    return 0;
}
