// center.cu -- double centring of the kernel matrix, K~ = K - 1 r' - r 1' + mu, in place.
// Replaces MLKernels.centerkernelmatrix! as called at /root/reference/src/kpca.jl:52.
// The row sums normally ride on the Gram epilogue (gram.cu, per-tile partials reduced here);
// fx_launch_rowsums is the standalone pass (warp-shuffle reduction) for matrices that did not
// come from the Gram kernel.  The update itself is a pure HBM stream: 16-byte vector loads and
// stores, 8 independent rows in flight per thread, one read + one write of K (16 n^2 bytes).
#include "fx_kernels.cuh"

namespace {

__global__ void rowpart_reduce_kernel(const double* __restrict__ rowpart, long ntiles, long rows_pad,
                                      double* __restrict__ rowsum) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;           // fixed order: deterministic
    long t = 0;
    for (; t + 4 <= ntiles; t += 4) {
        s0 += rowpart[(t + 0) * rows_pad + i];
        s1 += rowpart[(t + 1) * rows_pad + i];
        s2 += rowpart[(t + 2) * rows_pad + i];
        s3 += rowpart[(t + 3) * rows_pad + i];
    }
    for (; t < ntiles; ++t) s0 += rowpart[t * rows_pad + i];
    rowsum[i] = (s0 + s1) + (s2 + s3);
}

// one warp per row, 16-byte loads, warp-shuffle reduction
__global__ void rowsums_kernel(const double* __restrict__ K, long ldk, long rows, long cols_valid,
                               double* __restrict__ rowsum) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* kr = K + row * ldk;
    double s0 = 0.0, s1 = 0.0;
    const long cv2 = cols_valid & ~1L;
    for (long j = 2 * lane; j < cv2; j += 64) {
        const double2 v = *reinterpret_cast<const double2*>(kr + j);
        s0 += v.x;
        s1 += v.y;
    }
    if (lane == 0 && (cols_valid & 1)) s0 += kr[cols_valid - 1];
    const double s = warp_sum(s0 + s1);
    if (lane == 0) rowsum[row] = s;
}

// single CTA: colmean = sums / n, grand = sum(sums) / n^2 (fixed-shape tree: deterministic)
__global__ void means_kernel(const double* __restrict__ sums, long n, long n_ld,
                             double* __restrict__ colmean, double* __restrict__ grand) {
    __shared__ double red[32];
    double s = 0.0;
    const double inv_n = 1.0 / (double)n;
    for (long j = threadIdx.x; j < n_ld; j += blockDim.x) {
        const double v = (j < n) ? sums[j] : 0.0;
        colmean[j] = v * inv_n;
        s += v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *grand = v * inv_n * inv_n;
    }
}

constexpr int CROWS = 8;
__global__ void __launch_bounds__(256) center_kernel(double* __restrict__ K, long ldk, long rows_valid,
                                                     long cols_valid, long row0,
                                                     const double* __restrict__ colmean,
                                                     const double* __restrict__ grand) {
    const long j = 2 * ((long)blockIdx.x * blockDim.x + threadIdx.x);
    const long ib = (long)blockIdx.y * CROWS;
    if (j >= cols_valid) return;
    const double mu = *grand;
    const double c0 = colmean[j];
    const bool v1 = (j + 1) < cols_valid;
    const double c1 = v1 ? colmean[j + 1] : 0.0;
    double2 v[CROWS];
#pragma unroll
    for (int r = 0; r < CROWS; ++r)
        if (ib + r < rows_valid) v[r] = *reinterpret_cast<const double2*>(K + (ib + r) * ldk + j);
#pragma unroll
    for (int r = 0; r < CROWS; ++r)
        if (ib + r < rows_valid) {
            const double ri = colmean[row0 + ib + r];       // row mean == column mean (K symmetric)
            v[r].x = v[r].x - ri - c0 + mu;
            v[r].y = v1 ? (v[r].y - ri - c1 + mu) : v[r].y;
            *reinterpret_cast<double2*>(K + (ib + r) * ldk + j) = v[r];
        }
}

__global__ void trace_kernel(const double* __restrict__ K, long ldk, long rows_valid, long row0,
                             double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (long i = threadIdx.x; i < rows_valid; i += blockDim.x) s += K[i * ldk + row0 + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = v;
    }
}

}  // namespace

int fx_launch_rowpart_reduce(fx_ctx* ctx, const double* rowpart, int64_t ntiles, int64_t rows_pad,
                             double* rowsum) {
    if (rows_pad <= 0) return 0;
    rowpart_reduce_kernel<<<(unsigned)ceil_div(rows_pad, 256), 256, 0, ctx->stream>>>(rowpart, ntiles, rows_pad, rowsum);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_rowsums(fx_ctx* ctx, const double* K, int64_t ldk, int64_t rows, int64_t cols_valid,
                      double* rowsum) {
    if (rows <= 0) return 0;
    rowsums_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, ctx->stream>>>(K, ldk, rows, cols_valid, rowsum);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_means(fx_ctx* ctx, const double* sums, int64_t n, int64_t n_ld, double* colmean,
                    double* grand) {
    means_kernel<<<1, 1024, 0, ctx->stream>>>(sums, n, n_ld, colmean, grand);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_center(fx_ctx* ctx, double* K, int64_t ldk, int64_t rows_valid, int64_t cols_valid,
                     int64_t row0, const double* colmean, const double* grand) {
    if (rows_valid <= 0 || cols_valid <= 0) return 0;
    dim3 grid((unsigned)ceil_div(ceil_div(cols_valid, 2), 256), (unsigned)ceil_div(rows_valid, CROWS));
    center_kernel<<<grid, 256, 0, ctx->stream>>>(K, ldk, rows_valid, cols_valid, row0, colmean, grand);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_trace(fx_ctx* ctx, const double* K, int64_t ldk, int64_t rows_valid, int64_t row0,
                    double* out) {
    trace_kernel<<<1, 1024, 0, ctx->stream>>>(K, ldk, rows_valid, row0, out);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}
