// symm.cu -- Y = K V for the block eigensolver: K is this rank's row block of the symmetric
// kernel matrix (row-major, streamed from HBM exactly once per pass), V the all-gathered
// n x b block.  This is the only heavy operation of the top-k eigensolve that replaces the
// dense dsyevd_ call of /root/reference/src/common.jl:128-131 (via components_eig!, :161-165).
//
// Work unit = (128-row tile, k-split).  Per unit a 2-stage ring of {128 x 64 slab of K, 64 x b
// slab of V} is filled by cp.async.bulk (UBLKCP) row copies completing on mbarriers; 8 warps
// each own 16 rows x b columns of the output and feed DMMA.8x8x4 straight from LDS.128
// fragments (row stride 66 doubles => conflict-free quarter-warps).  Thread (g,t) of a warp
// owns k = 16*s + 4*t + e for the e-th MMA of a 16-wide step, for A and B alike, so one
// 32-byte load serves four MMAs.  K-split partials are reduced by a second tiny kernel in a
// fixed order (deterministic, no atomics).
#include "fx_kernels.cuh"

namespace {

constexpr int RT = 128;            // rows per tile
constexpr int BKC = 64;            // k-chunk (columns of K per stage)
constexpr int KS = BKC + 2;        // 66: smem row stride (doubles)
constexpr int NSTAGE = 2;

struct SymmParams {
    const double* K; long ldk; long rows_pad;
    const double* V; long ldr; long v_rank_stride;
    double* Yp;                    // [splits][bcols][rows_pad]
    int nchunks, splits, bcols;
};

template <int NB>
__global__ void __launch_bounds__(256, 1) symm_dmma_kernel(const SymmParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE_D = (RT + NB * 8) * KS;
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NSTAGE * STAGE_D * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const long rt = blockIdx.x / p.splits;
    const int ks = blockIdx.x % p.splits;
    const int c_begin = (int)((long)p.nchunks * ks / p.splits);
    const int c_end = (int)((long)p.nchunks * (ks + 1) / p.splits);
    const int nc = c_end - c_begin;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    constexpr uint32_t STAGE_TX = (RT + NB * 8) * BKC * 8;
    auto issue = [&](int ci, int s) {            // every thread: at most one row copy
        double* Ks = smem + (size_t)s * STAGE_D;
        const long kg = (long)(c_begin + ci) * BKC;
        if (tid == 0) mbar_arrive_expect_tx(&full[s], STAGE_TX);
        if (tid < RT) {
            bulk_g2s(Ks + tid * KS, p.K + (rt * RT + tid) * p.ldk + kg, BKC * 8, &full[s]);
        } else if (tid < RT + NB * 8) {
            const int c = tid - RT;
            const long rk = kg / p.ldr;
            bulk_g2s(Ks + (RT + c) * KS, p.V + rk * p.v_rank_stride + (long)c * p.ldr + (kg - rk * p.ldr),
                     BKC * 8, &full[s]);
        }
    };
    if (nc > 0) issue(0, 0);
    if (nc > 1) issue(1, 1);

    double acc[2][NB][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NB; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    for (int ci = 0; ci < nc; ++ci) {
        const int s = ci & 1;
        mbar_wait(&full[s], (ci >> 1) & 1);
        const double* Ks = smem + (size_t)s * STAGE_D;
        const double* Ar = Ks + (warp * 16 + g) * KS + 4 * t;
        const double* Br = Ks + (RT + g) * KS + 4 * t;
#pragma unroll
        for (int s16 = 0; s16 < BKC / 16; ++s16) {
            double a[2][4], b[NB][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const double2 lo = *reinterpret_cast<const double2*>(Ar + mt * 8 * KS + s16 * 16);
                const double2 hi = *reinterpret_cast<const double2*>(Ar + mt * 8 * KS + s16 * 16 + 2);
                a[mt][0] = lo.x; a[mt][1] = lo.y; a[mt][2] = hi.x; a[mt][3] = hi.y;
            }
#pragma unroll
            for (int nt = 0; nt < NB; ++nt) {
                const double2 lo = *reinterpret_cast<const double2*>(Br + nt * 8 * KS + s16 * 16);
                const double2 hi = *reinterpret_cast<const double2*>(Br + nt * 8 * KS + s16 * 16 + 2);
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
        __syncthreads();
        if (ci + 2 < nc) issue(ci + 2, s);
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
__global__ void symm_naive_kernel(const SymmParams p, double* Y) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (row >= p.rows_pad) return;
    const long ncols = (long)p.nchunks * BKC;
    double s = 0.0;
    for (long j = 0; j < ncols; ++j) {
        const long rk = j / p.ldr;
        s += p.K[row * p.ldk + j] * p.V[rk * p.v_rank_stride + (long)c * p.ldr + (j - rk * p.ldr)];
    }
    Y[(long)c * p.rows_pad + row] = s;
}

template <int NB>
int launch_nb(fx_ctx* ctx, const SymmParams& p) {
    constexpr int STAGE_D = (RT + NB * 8) * KS;
    constexpr int SMEM = NSTAGE * STAGE_D * 8 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        FX_CUDA(cudaFuncSetAttribute(symm_dmma_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    const unsigned grid = (unsigned)((p.rows_pad / RT) * p.splits);
    symm_dmma_kernel<NB><<<grid, 256, SMEM, ctx->stream>>>(p);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
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
    SymmParams p;
    p.K = a.K; p.ldk = a.ldk; p.rows_pad = a.rows_pad; p.V = a.V; p.ldr = a.ldr;
    p.v_rank_stride = a.v_rank_stride; p.nchunks = (int)(a.ldk / BKC); p.bcols = a.bcols;
    p.splits = a.splits < 1 ? 1 : a.splits;
    if (variant != 0) {
        dim3 grid((unsigned)ceil_div(a.rows_pad, 128), (unsigned)a.bcols);
        symm_naive_kernel<<<grid, 128, 0, ctx->stream>>>(p, a.Y);
        ctx->launches++;
        FX_LAUNCH_CHECK();
        return 0;
    }
    p.Yp = (p.splits == 1) ? a.Y : a.Ypart;
    switch (a.bcols / 8) {
        case 1: FX_TRY(launch_nb<1>(ctx, p)); break;
        case 2: FX_TRY(launch_nb<2>(ctx, p)); break;
        case 3: FX_TRY(launch_nb<3>(ctx, p)); break;
        case 4: FX_TRY(launch_nb<4>(ctx, p)); break;
        case 5: FX_TRY(launch_nb<5>(ctx, p)); break;
        case 6: FX_TRY(launch_nb<6>(ctx, p)); break;
        case 7: FX_TRY(launch_nb<7>(ctx, p)); break;
        default: FX_TRY(launch_nb<8>(ctx, p)); break;
    }
    if (p.splits > 1) {
        const long total = (long)a.bcols * a.rows_pad;
        symm_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, ctx->stream>>>(a.Ypart, p.splits, total, a.Y);
        ctx->launches++;
        FX_LAUNCH_CHECK();
    }
    return 0;
}
