// transform.cu -- fused out-of-sample projection:  out = map(Z X') W - 1 bias'
// One kernel builds each 128 x 64 tile of the cross kernel matrix k(Z, X) in registers
// (DMMA over the feature dimension), applies the kernel map, and immediately contracts it with
// the matching 64 x k slab of W on the tensor pipe again -- the C fragments of the first MMA
// are reused as A fragments of the second (thread t's two accumulator columns become the
// k-indices {2t} and {2t+1} of two MMAs), so k(Z,X) never exists in memory (1.05 TB for C5).
// Replaces transform_kpca / transform_klda (/root/reference/src/kpca.jl:64-68, klda.jl:71):
// kernelmatrix(k, Z, X) * W, with the centring reduced to the k-vector bias r'W.
//
// Warp roles: 8 MMA warps (16 query rows x 64 training columns each) + 1 producer warp that
// feeds a 4-stage ring of 16-feature slabs {Z: 16x128, X: 16x64} and a 2-stage ring of W slabs
// through cp.async.bulk + mbarrier full/empty pairs (no CTA-wide barrier in the main loop, so
// the exp() epilogues of different warps overlap other warps' MMAs).
#include "fx_kernels.cuh"

namespace {

constexpr int TQ = 128, TN = 64, TBK = 16, TST = 4;
// Row pitches (doubles) of the staged slabs, chosen for conflict-free fragment loads: Z rows are
// read 16 bytes at a time at kr*ZS + 2g (kr = 4kk + t; a quarter-warp is g in {2h, 2h+1} x t = 0..3,
// so ZS/2 = 2 (mod 8) gives 8 distinct 16-byte units), X rows 8 bytes at a time at kr*XS + 8nt + g.
constexpr int ZS = TQ + 4, XS = TN + 4, WS = 72;
constexpr int TSTAGE_D = TBK * ZS + TBK * XS;            // 3200 doubles

struct TrParams {
    const double* Zp; long ldz; const double* nz;
    const double* Xp; long ldx; const double* nx;
    const double* W; long ldw;
    const double* bias;
    double* out; long ldo;          // [nsplit][kc + 1][ldo] when nsplit > 1 (bias applied in the reduce;
                                    // the extra column carries the partial row sums of k(Z, X))
    long n_valid;
    int kt, ntiles, nsplit, kc;
    // Exact double centring for components that are not orthogonal to 1:
    //   out -= (rowmean_q - grand) * wsum[c],  rowmean_q = (1/n) sum_j k(z_q, x_j).
    // want_rz != 0 makes the MMA warps accumulate the row sums of the mapped tile.
    int want_rz;
    const double* wsum;             // [kc] 1'W, or null (no correction applied)
    double grand, inv_n;
    double* rz_out;                 // [m_pad] raw row sums of k(Z, X), or null
    KernelMap km;
};

// Kernel map of one warp's 16 x 64 slice of the cross kernel tile, in registers.  Interior tiles
// run the bare arithmetic (no predicates: the compiler otherwise wraps every element in its own
// branch region and the 32 maps of a thread run back to back instead of interleaved); the last
// tile of a ragged training set masks the padding columns.
// Accumulator (nt, q) of thread (g, t) is training column j0 + 8 nt + 2t + q.
template <int KID, bool EDGE>
__device__ __forceinline__ void transform_map(const KernelMap& km, double (&acc1)[2][8][2], const double (&nzs)[2],
                                              const double* __restrict__ nxt, long j0, long n_valid, int t,
                                              bool want_rz, double (&rz)[2]) {
    const double two_g = 2.0 * km.p0, neg_g = -km.p0;
    if (KID == 0) {
        // Gaussian: 8 elements (two column blocks) at a time through the staged exp, so that the
        // dependency chains of the exp are interleaved (fx_exp_nonpos_vec)
#pragma unroll
        for (int n2 = 0; n2 < 8; n2 += 2) {
            double x[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2 v = *reinterpret_cast<const double2*>(nxt + (n2 + h) * 8 + 2 * t);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    x[h * 4 + mt * 2] = fma(two_g, acc1[mt][n2 + h][0], fma(neg_g, v.x, nzs[mt]));
                    x[h * 4 + mt * 2 + 1] = fma(two_g, acc1[mt][n2 + h][1], fma(neg_g, v.y, nzs[mt]));
                }
            }
            fx_exp_nonpos_vec<8>(x);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        double v = x[h * 4 + mt * 2 + q];
                        if (EDGE) v = (j0 + (n2 + h) * 8 + 2 * t + q < n_valid) ? v : 0.0;
                        acc1[mt][n2 + h][q] = v;
                        if (want_rz) rz[mt] += v;
                    }
        }
        return;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long j = j0 + nt * 8 + 2 * t + q;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                double v = kernel_map_t<KID>(km, acc1[mt][nt][q], 0.0, 0.0, false);
                if (EDGE) v = (j < n_valid) ? v : 0.0;
                acc1[mt][nt][q] = v;
                if (want_rz) rz[mt] += v;
            }
        }
    }
}

template <int NBK, int KID>
__global__ void __launch_bounds__(288, 1) transform_kernel(const TrParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* Wsm = smem + TST * TSTAGE_D;                           // 2 x [NBK*8][WS]
    double* nxs = Wsm + 2 * NBK * 8 * WS;                          // 2 x [TN] squared norms of the tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(nxs + 2 * TN);
    uint64_t* full = bars;            // [TST]
    uint64_t* empty = bars + TST;     // [TST]
    uint64_t* wfull = bars + 2 * TST; // [2]
    uint64_t* wempty = wfull + 2;     // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const long q0 = (long)blockIdx.x * TQ;
    const int sp = blockIdx.y;
    const int t_begin = (int)((long)p.ntiles * sp / p.nsplit);
    const int t_end = (int)((long)p.ntiles * (sp + 1) / p.nsplit);
    const int KT = p.kt;

    if (tid == 0) {
        for (int s = 0; s < TST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 8); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 8) {
        // ------------------------------ producer warp ------------------------------
        int it = 0;
        for (int tile = t_begin; tile < t_end; ++tile) {
            const long j0 = (long)tile * TN;
            const int wi = tile - t_begin, ws = wi & 1;
            mbar_wait(&wempty[ws], ((wi >> 1) & 1) ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&wfull[ws], NBK * 8 * TN * 8 + (KID == 0 ? TN * 8 : 0));
            __syncwarp();
            for (int c = lane; c < NBK * 8; c += 32)
                bulk_g2s(Wsm + (ws * NBK * 8 + c) * WS, p.W + (long)c * p.ldw + j0, TN * 8, &wfull[ws]);
            if (KID == 0 && lane == 0) bulk_g2s(nxs + ws * TN, p.nx + j0, TN * 8, &wfull[ws]);
            for (int kt = 0; kt < KT; ++kt, ++it) {
                const int s = it % TST;
                mbar_wait(&empty[s], ((it / TST) & 1) ^ 1);
                double* Zs = smem + (size_t)s * TSTAGE_D;
                double* Xs = Zs + TBK * ZS;
                if (lane == 0) mbar_arrive_expect_tx(&full[s], TBK * (TQ + TN) * 8);
                __syncwarp();
                if (lane < TBK)
                    bulk_g2s(Zs + lane * ZS, p.Zp + (long)(kt * TBK + lane) * p.ldz + q0, TQ * 8, &full[s]);
                else
                    bulk_g2s(Xs + (lane - TBK) * XS, p.Xp + (long)(kt * TBK + lane - TBK) * p.ldx + j0,
                             TN * 8, &full[s]);
            }
        }
        return;
    }

    // -------------------------------- MMA warps --------------------------------
    double acc2[2][NBK][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nk = 0; nk < NBK; ++nk) acc2[mt][nk][0] = acc2[mt][nk][1] = 0.0;
    constexpr bool gauss = (KID == 0);
    // Row mapping: accumulator row (mt, g) is query q0 + warp*16 + 2g + mt (one 16-byte load per
    // k-step feeds both MMA row blocks).
    double nzs[2] = {0.0, 0.0}, rz[2] = {0.0, 0.0};
    if (gauss) {
        nzs[0] = -p.km.p0 * p.nz[q0 + warp * 16 + 2 * g];
        nzs[1] = -p.km.p0 * p.nz[q0 + warp * 16 + 2 * g + 1];
    }
    const bool want_rz = p.want_rz != 0;

    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const long j0 = (long)tile * TN;
        double acc1[2][8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc1[mt][nt][0] = acc1[mt][nt][1] = 0.0;

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % TST;
            mbar_wait(&full[s], (it / TST) & 1);
            const double* Zs = smem + (size_t)s * TSTAGE_D + warp * 16 + 2 * g;
            const double* Xs = smem + (size_t)s * TSTAGE_D + TBK * ZS + g;
#pragma unroll
            for (int kk = 0; kk < TBK / 4; ++kk) {
                const int kr = kk * 4 + t;
                double a[2], b[8];
                {
                    const double2 v = *reinterpret_cast<const double2*>(Zs + kr * ZS);
                    a[0] = v.x; a[1] = v.y;
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) b[nt] = Xs[kr * XS + nt * 8];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) dmma(acc1[mt][nt][0], acc1[mt][nt][1], a[mt], b[nt]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        // kernel map in registers (the tile's W slab and norms were staged by the producer)
        const int wi = tile - t_begin, ws = wi & 1;
        mbar_wait(&wfull[ws], (wi >> 1) & 1);
        if (j0 + TN <= p.n_valid) transform_map<KID, false>(p.km, acc1, nzs, nxs + ws * TN, j0, p.n_valid, t, want_rz, rz);
        else                      transform_map<KID, true>(p.km, acc1, nzs, nxs + ws * TN, j0, p.n_valid, t, want_rz, rz);

        // second contraction: acc2 += Ktile (16 x 64 per warp) * W[j0:j0+64, :]; the C fragments of the
        // first MMA are the A fragments of the second (k index t <-> tile column 8 nt + 2t + q)
        const double* Wt = Wsm + (size_t)ws * NBK * 8 * WS + g * WS + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            double2 wv[NBK];
#pragma unroll
            for (int nk = 0; nk < NBK; ++nk) wv[nk] = *reinterpret_cast<const double2*>(Wt + nk * 8 * WS + nt * 8);
            // the two MMAs that update the same accumulator are kept 2*NBK issues apart (a DMMA that
            // waits for its predecessor's result stalls the warp for the pipe latency)
#pragma unroll
            for (int nk = 0; nk < NBK; ++nk)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma(acc2[mt][nk][0], acc2[mt][nk][1], acc1[mt][nt][0], wv[nk].x);
#pragma unroll
            for (int nk = 0; nk < NBK; ++nk)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma(acc2[mt][nk][0], acc2[mt][nk][1], acc1[mt][nt][1], wv[nk].y);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&wempty[ws]);
    }

    // row sums of k(Z, X) over this split: reduce the four t-lanes of a row
    if (want_rz) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            rz[mt] += __shfl_xor_sync(0xffffffffu, rz[mt], 1);
            rz[mt] += __shfl_xor_sync(0xffffffffu, rz[mt], 2);
        }
    }
    const bool split = p.nsplit > 1;
    double* out = p.out + (long)sp * (split ? p.kc + 1 : p.kc) * p.ldo;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const long row = q0 + warp * 16 + 2 * g + mt;
        double corr = 0.0;                       // (rowmean - grand), applied with wsum below
        if (want_rz) {
            if (split) { if (t == 0) out[(long)p.kc * p.ldo + row] = rz[mt]; }
            else {
                if (p.rz_out && t == 0) p.rz_out[row] = rz[mt];
                corr = rz[mt] * p.inv_n - p.grand;
            }
        }
#pragma unroll
        for (int nk = 0; nk < NBK; ++nk) {
            const int c = nk * 8 + 2 * t;
            double b0 = 0.0, b1 = 0.0;
            if (!split) {
                if (p.bias) { b0 = p.bias[c]; b1 = p.bias[c + 1]; }
                if (p.wsum) { b0 = fma(corr, p.wsum[c], b0); b1 = fma(corr, p.wsum[c + 1], b1); }
            }
            out[(long)c * p.ldo + row] = acc2[mt][nk][0] - b0;
            out[(long)(c + 1) * p.ldo + row] = acc2[mt][nk][1] - b1;
        }
    }
}

// one thread per query: sums the split partials in a fixed order, applies bias and the row-mean term
__global__ void transform_reduce_kernel(const double* part, int nsplit, int kc, long ldo, const double* bias,
                                        int want_rz, const double* wsum, double grand, double inv_n,
                                        double* rz_out, double* out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ldo) return;
    double corr = 0.0;
    if (want_rz) {
        double r = 0.0;
        for (int k = 0; k < nsplit; ++k) r += part[((long)k * (kc + 1) + kc) * ldo + i];
        if (rz_out) rz_out[i] = r;
        corr = r * inv_n - grand;
    }
    for (int c = 0; c < kc; ++c) {
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += part[((long)k * (kc + 1) + c) * ldo + i];
        double b = bias ? bias[c] : 0.0;
        if (wsum) b = fma(corr, wsum[c], b);
        out[(long)c * ldo + i] = s - b;
    }
}

__global__ void transform_naive_kernel(const TrParams p, int d_pad, long m_pad) {
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (q >= m_pad) return;
    double s = 0.0, rsum = 0.0;
    for (long j = 0; j < p.n_valid; ++j) {
        double dot = 0.0;
        for (int f = 0; f < d_pad; ++f) dot += p.Zp[(long)f * p.ldz + q] * p.Xp[(long)f * p.ldx + j];
        const double ni = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.nz[q] : 0.0;
        const double nj = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.nx[j] : 0.0;
        const double kv = kernel_map(p.km, dot, ni, nj, false);
        s += kv * p.W[(long)c * p.ldw + j];
        rsum += kv;
    }
    double b = p.bias ? p.bias[c] : 0.0;
    if (p.want_rz && p.wsum) b += (rsum * p.inv_n - p.grand) * p.wsum[c];
    if (p.want_rz && p.rz_out && c == 0) p.rz_out[q] = rsum;
    p.out[(long)c * p.ldo + q] = s - b;
}

template <int NBK, int KID>
int launch_tr2(fx_ctx* ctx, const TrParams& p, long m_pad) {
    constexpr int SMEM = (TST * TSTAGE_D + 2 * NBK * 8 * WS + 2 * TN) * 8 + 128;
    FX_SET_SMEM_ONCE(ctx, (transform_kernel<NBK, KID>), SMEM);
    dim3 grid((unsigned)(m_pad / TQ), (unsigned)p.nsplit);
    transform_kernel<NBK, KID><<<grid, 288, SMEM, ctx->stream>>>(p);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}
template <int NBK>
int launch_tr(fx_ctx* ctx, const TrParams& p, long m_pad) {
    switch (kernel_variant(p.km.id, p.km.ipow)) {
        case 0: return launch_tr2<NBK, 0>(ctx, p, m_pad);
        case 1: return launch_tr2<NBK, 1>(ctx, p, m_pad);
        case 2: return launch_tr2<NBK, 2>(ctx, p, m_pad);
        case 3: return launch_tr2<NBK, 3>(ctx, p, m_pad);
        case 4: return launch_tr2<NBK, 4>(ctx, p, m_pad);
        case 5: return launch_tr2<NBK, 5>(ctx, p, m_pad);
        default: return launch_tr2<NBK, 6>(ctx, p, m_pad);
    }
}

}  // namespace

// scratch doubles needed for the split partials (0 when one split suffices)
size_t fx_transform_scratch(fx_ctx* ctx, int64_t m_pad, int64_t n_ld, int kc, int* nsplit_out) {
    const int64_t qtiles = m_pad / TQ, ntiles = n_ld / TN;
    int64_t ns = 1;
    if (qtiles < 2 * ctx->sm_count) ns = ceil_div(2 * (int64_t)ctx->sm_count, qtiles);
    if (ns > ntiles / 4) ns = ntiles / 4;
    if (ns < 1) ns = 1;
    if (ns > 64) ns = 64;
    if (nsplit_out) *nsplit_out = (int)ns;
    return ns > 1 ? (size_t)ns * (kc + 1) * m_pad : 0;      // + 1: partial row sums of k(Z, X)
}

int fx_launch_transform_ex(fx_ctx* ctx, const TransformArgs& a, int variant, double* scratch, int nsplit) {
    if (a.m_pad % TQ || a.n_ld % TN || a.d_pad % TBK || a.kcols % 8 || a.kcols < 8 || a.kcols > 64) {
        fx_set_error("transform: bad shapes m_pad=%lld n_ld=%lld d_pad=%lld k=%d", (long long)a.m_pad,
                     (long long)a.n_ld, (long long)a.d_pad, a.kcols);
        return FX_ERR_CUDA;
    }
    if (a.m_pad == 0) return 0;
    TrParams p;
    p.Zp = a.Zp; p.ldz = a.ldz; p.nz = a.nz; p.Xp = a.Xp; p.ldx = a.ldx; p.nx = a.nx;
    p.W = a.W; p.ldw = a.ldw; p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.n_valid = a.n_valid;
    p.kt = (int)(a.d_pad / TBK); p.ntiles = (int)(a.n_ld / TN); p.nsplit = 1; p.kc = a.kcols;
    p.want_rz = (a.wsum || a.rz_out) ? 1 : 0; p.wsum = a.wsum; p.grand = a.grand; p.rz_out = a.rz_out;
    p.inv_n = a.n_valid > 0 ? 1.0 / (double)a.n_valid : 0.0;
    p.km.id = a.kernel_id; p.km.p0 = a.kp[0]; p.km.p1 = a.kp[1]; p.km.p2 = a.kp[2];
    double tmp[3];
    if (fx_make_kernel_map(a.kernel_id, a.kp, &p.km.ipow, tmp)) { fx_set_error("fx_launch_transform_ex: invalid argument 1 (fx_make_kernel_map(a.kernel_id, a.kp, &p.km.ipow, tmp))"); return -1; }
    if (variant != 0) {
        dim3 grid((unsigned)ceil_div(a.m_pad, 128), (unsigned)a.kcols);
        transform_naive_kernel<<<grid, 128, 0, ctx->stream>>>(p, (int)a.d_pad, a.m_pad);
        ctx->launches++;
        FX_LAUNCH_CHECK();
        return 0;
    }
    if (nsplit > 1 && scratch) {
        p.nsplit = nsplit;
        p.out = scratch;
    }
    switch (a.kcols / 8) {
        case 1: FX_TRY(launch_tr<1>(ctx, p, a.m_pad)); break;
        case 2: FX_TRY(launch_tr<2>(ctx, p, a.m_pad)); break;
        case 3: FX_TRY(launch_tr<3>(ctx, p, a.m_pad)); break;
        case 4: FX_TRY(launch_tr<4>(ctx, p, a.m_pad)); break;
        case 5: FX_TRY(launch_tr<5>(ctx, p, a.m_pad)); break;
        case 6: FX_TRY(launch_tr<6>(ctx, p, a.m_pad)); break;
        case 7: FX_TRY(launch_tr<7>(ctx, p, a.m_pad)); break;
        default: FX_TRY(launch_tr<8>(ctx, p, a.m_pad)); break;
    }
    if (p.nsplit > 1) {
        transform_reduce_kernel<<<(unsigned)ceil_div(a.ldo, 128), 128, 0, ctx->stream>>>(
            scratch, p.nsplit, a.kcols, a.ldo, a.bias, p.want_rz, p.wsum, p.grand, p.inv_n, p.rz_out, a.out);
        ctx->launches++;
        FX_LAUNCH_CHECK();
    }
    return 0;
}

int fx_launch_transform(fx_ctx* ctx, const TransformArgs& a, int variant) {
    return fx_launch_transform_ex(ctx, a, variant, nullptr, 1);
}
