// tsgemm.cu -- tall-skinny fp64 products for the block eigensolver's projections
// (H = Q'W, W -= QH, Q S, W M) on DMMA.8x8x4 with fragments loaded straight from global
// memory / shared memory in the MMA's own layout (no transposes, 16-byte loads):
//   ts_tn: both operands are "k-major" (k = the long row dimension is contiguous in a
//          column-major n x m matrix), thread (g,t) owns rows r+4t..r+4t+3 of a 16-row step.
//   ts_nn: the A operand is "m-major"; thread g owns rows (2g, 2g+1) so one 16-byte load
//          feeds two MMAs; the small right factor sits in shared memory.
// These stand in for the O(n m b) part of the work LAPACK's dsyevd does internally
// (/root/reference/src/common.jl:128-131); partial sums are combined in a fixed order.
#include "fx_kernels.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// H = A' B
// ---------------------------------------------------------------------------------------
struct TnParams {
    const double* A; const double* B; long ld;
    long rows;             // multiple of 16
    long rows_per_chunk;   // multiple of 128
    int ma, nb;            // multiples of 8
    double* part;          // [chunks][ma][nb_ld]  (column-major ma x nb per chunk, ld = ma)
    int nchunks;
};

__global__ void __launch_bounds__(256, 1) ts_tn_kernel(const TnParams p) {
    __shared__ double red[64 * 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int chunk = blockIdx.x;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.z * 32;
    const int mtiles = min(8, (p.ma - m0) / 8), ntiles = min(4, (p.nb - n0) / 8);
    const long r_begin = (long)chunk * p.rows_per_chunk;
    const long r_end = min(p.rows, r_begin + p.rows_per_chunk);

    double acc[8][4][2];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    const double* Ab = p.A + (long)(m0 + g) * p.ld + 4 * t;
    const double* Bb = p.B + (long)(n0 + g) * p.ld + 4 * t;
    for (long r = r_begin + warp * 16; r < r_end; r += 8 * 16) {
        double a[8][4], b[4][4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
            if (mt < mtiles) {
                const double2 lo = *reinterpret_cast<const double2*>(Ab + (long)mt * 8 * p.ld + r);
                const double2 hi = *reinterpret_cast<const double2*>(Ab + (long)mt * 8 * p.ld + r + 2);
                a[mt][0] = lo.x; a[mt][1] = lo.y; a[mt][2] = hi.x; a[mt][3] = hi.y;
            }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            if (nt < ntiles) {
                const double2 lo = *reinterpret_cast<const double2*>(Bb + (long)nt * 8 * p.ld + r);
                const double2 hi = *reinterpret_cast<const double2*>(Bb + (long)nt * 8 * p.ld + r + 2);
                b[nt][0] = lo.x; b[nt][1] = lo.y; b[nt][2] = hi.x; b[nt][3] = hi.y;
            }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
                if (mt < mtiles)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        if (nt < ntiles) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt][e], b[nt][e]);
    }

    // cross-warp reduction in a fixed order (warp 0 .. 7)
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int idx = (nt * 8 + 2 * t) * 64 + mt * 8 + g;
                    if (w == 0) {
                        red[idx] = acc[mt][nt][0];
                        red[idx + 64] = acc[mt][nt][1];
                    } else {
                        red[idx] += acc[mt][nt][0];
                        red[idx + 64] += acc[mt][nt][1];
                    }
                }
        }
        __syncthreads();
    }
    double* out = p.part + (long)chunk * p.ma * p.nb;
    for (int idx = tid; idx < 64 * 32; idx += 256) {
        const int c = idx >> 6, r = idx & 63;
        if (r < mtiles * 8 && c < ntiles * 8) out[(long)(n0 + c) * p.ma + m0 + r] = red[idx];
    }
}

// one warp per output element, lanes stride over the chunk partials (fixed order: deterministic)
__global__ void ts_tn_reduce_kernel(const double* __restrict__ part, int nchunks, int ma, int nb,
                                    double* __restrict__ H, int ldh) {
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (idx >= ma * nb) return;
    double s = 0.0;
    for (int c = lane; c < nchunks; c += 32) s += part[(long)c * ma * nb + idx];
    s = warp_sum(s);
    if (lane == 0) H[(long)(idx / ma) * ldh + (idx % ma)] = s;
}

__global__ void ts_tn_naive_kernel(const double* A, const double* B, long ld, long rows, int ma, int nb,
                                   double* H, int ldh) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ma * nb) return;
    const int i = idx % ma, j = idx / ma;
    double s = 0.0;
    for (long r = 0; r < rows; ++r) s += A[(long)i * ld + r] * B[(long)j * ld + r];
    H[(long)j * ldh + i] = s;
}

// ---------------------------------------------------------------------------------------
// C = beta C + alpha A M
// ---------------------------------------------------------------------------------------
struct NnParams {
    const double* A; long ld; int ma;
    const double* M; int ldm; int nc;
    double alpha, beta;
    double* C;
};

__global__ void __launch_bounds__(256, 1) ts_nn_kernel(const NnParams p) {
    extern __shared__ __align__(16) double Ms[];          // [32][ldms], k contiguous per column
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.y * 32;
    const int ntiles = min(4, (p.nc - n0) / 8);
    const int ma16 = (p.ma + 15) & ~15;
    const int ldms = ma16 + 2;
    for (int idx = tid; idx < 32 * ma16; idx += 256) {
        const int c = idx / ma16, k = idx - c * ma16;
        Ms[c * ldms + k] = (k < p.ma && c < ntiles * 8) ? p.M[(long)(n0 + c) * p.ldm + k] : 0.0;
    }
    __syncthreads();

    const long i0 = (long)blockIdx.x * 128 + warp * 16 + 2 * g;
    double acc[2][4][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[e][nt][0] = acc[e][nt][1] = 0.0;

    for (int k0 = 0; k0 < ma16; k0 += 16) {
        double2 av[4];
        double b[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int col = k0 + 4 * t + s;
            av[s] = (col < p.ma) ? *reinterpret_cast<const double2*>(p.A + (long)col * p.ld + i0)
                                 : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const double2 lo = *reinterpret_cast<const double2*>(&Ms[(nt * 8 + g) * ldms + k0 + 4 * t]);
            const double2 hi = *reinterpret_cast<const double2*>(&Ms[(nt * 8 + g) * ldms + k0 + 4 * t + 2]);
            b[nt][0] = lo.x; b[nt][1] = lo.y; b[nt][2] = hi.x; b[nt][3] = hi.y;
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                dmma(acc[0][nt][0], acc[0][nt][1], av[s].x, b[nt][s]);
                dmma(acc[1][nt][0], acc[1][nt][1], av[s].y, b[nt][s]);
            }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
        if (nt < ntiles) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double* dst = p.C + (long)(n0 + nt * 8 + 2 * t + q) * p.ld + i0;
                double2 v = make_double2(p.alpha * acc[0][nt][q], p.alpha * acc[1][nt][q]);
                if (p.beta != 0.0) {
                    const double2 o = *reinterpret_cast<const double2*>(dst);
                    v.x += p.beta * o.x;
                    v.y += p.beta * o.y;
                }
                *reinterpret_cast<double2*>(dst) = v;
            }
        }
}

__global__ void ts_nn_naive_kernel(const NnParams p, long rows) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= rows) return;
    double s = 0.0;
    for (int k = 0; k < p.ma; ++k) s += p.A[(long)k * p.ld + i] * p.M[(long)c * p.ldm + k];
    double* dst = p.C + (long)c * p.ld + i;
    *dst = (p.beta != 0.0 ? p.beta * *dst : 0.0) + p.alpha * s;
}

int tn_chunks(fx_ctx* ctx, int ma, int nb, int64_t rows, int64_t* rows_per_chunk) {
    const int blocks = (int)(ceil_div(ma, 64) * ceil_div(nb, 32));
    int64_t want = ceil_div(2 * (int64_t)ctx->sm_count, blocks);
    int64_t maxc = ceil_div(rows, 128);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    *rows_per_chunk = round_up(ceil_div(rows, want), 128);
    return (int)ceil_div(rows, *rows_per_chunk);
}

}  // namespace

size_t fx_ts_tn_scratch(fx_ctx* ctx, int ma, int nb, int64_t rows) {
    int64_t rpc;
    const int nch = tn_chunks(ctx, ma, nb, rows, &rpc);
    return (size_t)nch * ma * nb;
}

int fx_launch_ts_tn(fx_ctx* ctx, const double* A, int ma, const double* B, int nb, int64_t ld,
                    int64_t rows, double* H, int ldh, double* scratch, size_t scratch_doubles,
                    int variant) {
    if (ma % 8 || nb % 8 || rows % 16 || ld % 2) {
        fx_set_error("ts_tn: bad shapes ma=%d nb=%d rows=%lld", ma, nb, (long long)rows);
        return FX_ERR_CUDA;
    }
    if (variant != 0) {
        ts_tn_naive_kernel<<<(unsigned)ceil_div((int64_t)ma * nb, 128), 128, 0, ctx->stream>>>(A, B, ld, rows, ma, nb, H, ldh);
        ctx->launches++;
        FX_LAUNCH_CHECK();
        return 0;
    }
    TnParams p;
    p.A = A; p.B = B; p.ld = ld; p.rows = rows; p.ma = ma; p.nb = nb; p.part = scratch;
    int64_t rpc;
    p.nchunks = tn_chunks(ctx, ma, nb, rows, &rpc);
    p.rows_per_chunk = rpc;
    if ((size_t)p.nchunks * ma * nb > scratch_doubles) {
        fx_set_error("ts_tn: scratch too small");
        return FX_ERR_CUDA;
    }
    dim3 grid((unsigned)p.nchunks, (unsigned)ceil_div(ma, 64), (unsigned)ceil_div(nb, 32));
    ts_tn_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    ts_tn_reduce_kernel<<<(unsigned)ceil_div((int64_t)ma * nb, 8), 256, 0, ctx->stream>>>(scratch, p.nchunks, ma, nb, H, ldh);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_ts_nn(fx_ctx* ctx, const double* A, int ma, const double* M, int ldm, int nc,
                    double alpha, double beta, double* C, int64_t ld, int64_t rows, int variant) {
    if (ma % 8 || nc % 8 || rows % 128 || ld % 2) {
        fx_set_error("ts_nn: bad shapes ma=%d nc=%d rows=%lld", ma, nc, (long long)rows);
        return FX_ERR_CUDA;
    }
    if (rows == 0 || nc == 0) return 0;
    NnParams p;
    p.A = A; p.ld = ld; p.ma = ma; p.M = M; p.ldm = ldm; p.nc = nc; p.alpha = alpha; p.beta = beta; p.C = C;
    if (variant != 0) {
        dim3 grid((unsigned)ceil_div(rows, 128), (unsigned)nc);
        ts_nn_naive_kernel<<<grid, 128, 0, ctx->stream>>>(p, rows);
        ctx->launches++;
        FX_LAUNCH_CHECK();
        return 0;
    }
    const int ma16 = (ma + 15) & ~15;
    const size_t smem = (size_t)32 * (ma16 + 2) * 8;
    FX_SET_SMEM_ONCE(ctx, ts_nn_kernel, FX_MAX_DYN_SMEM);
    dim3 grid((unsigned)(rows / 128), (unsigned)ceil_div(nc, 32));
    ts_nn_kernel<<<grid, 256, smem, ctx->stream>>>(p);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}
