// linear.cu -- the linear solvers of the reference's export list on the same device kernels
// (SURVEY.md 8f rank 2): pca / transform_pca (/root/reference/src/kpca.jl:8-37) and
// lda_matrices! / lda_components! / lda! / transform_lda (src/klda.jl:34-65).  These are p x p
// problems: the Gram-of-features X'X is a tall-skinny product (tsgemm.cu), the eigenproblems run
// on the device Jacobi solver (smalldense.cu); p <= 512.
//   pca: Sigma = H'H (kpca.jl:9), regularize!/perturb! folded into the eigenvalues
//        (kpca.jl:10-11; same eigenvectors), lambda/(n-1) (kpca.jl:22).  For alpha = eps = 0 the
//        reference takes the SVD path (kpca.jl:21): sigma^2 of H = eigenvalues of H'H.
//   lda: the intended pencil (Sigma_b, Sigma_w) -- upstream passes (Sigma_b, Sigma_b), klda.jl:42;
//        generalized symmetric eigenproblem as sygvd itype 1 (common.jl:172): Sigma_w = U D U',
//        Y = U D^-1/2, eig(Y' Sigma_b Y) = Z L Z', V = Y Z  (V' Sigma_w V = I), descending.
#include "fx_kernels.cuh"
#include <math.h>

namespace {

__global__ void ones_kernel(double* o, long ld, long n) {      // [8][ld]: column 0 = 1 on valid rows
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    for (int c = 0; c < 8; ++c) o[(long)c * ld + i] = (c == 0 && i < n) ? 1.0 : 0.0;
}
// M (8 x pp, ld 8): row 0 = -colsum[j]/n
__global__ void neg_mean_kernel(const double* S, int lds, int p, int pp, double inv_n, double* M, double* mu) {
    for (int j = threadIdx.x; j < pp; j += blockDim.x) {
        const double m = (j < p) ? S[j] * inv_n : 0.0;          // S is (pp x 8) col-major, column 0
        for (int r = 0; r < 8; ++r) M[r + j * 8] = (r == 0) ? -m : 0.0;
        if (mu) mu[j] = m;
    }
    (void)lds;
}
// A = (1-alpha) A + (alpha * tr(A)/p + eps) I on the leading p x p block (ld = lda); trace out
__global__ void regularize_kernel(double* A, int lda, int p, double alpha, double eps, double* trace_out) {
    __shared__ double tr;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < p; ++i) s += A[i + (long)i * lda];
        tr = s;
        if (trace_out) *trace_out = s;
    }
    __syncthreads();
    const double shift = alpha * tr / p + eps;
    for (int idx = threadIdx.x; idx < p * p; idx += blockDim.x) {
        const int i = idx % p, j = idx / p;
        double v = A[i + (long)j * lda] * (1.0 - alpha);           // own element only (no cross reads)
        if (i == j) v += shift;
        A[i + (long)j * lda] = v;
    }
}
__global__ void symmetrize_full_kernel(double* A, int lda, int p) {
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < p * p; idx += blockDim.x * gridDim.x) {
        const int i = idx % p, j = idx / p;
        if (i < j) { const double v = 0.5 * (A[i + (long)j * lda] + A[j + (long)i * lda]); A[i + (long)j * lda] = v; A[j + (long)i * lda] = v; }
    }
}
// class tables for LDA: means (kp x pp, ld kp) from sums; H_b rows; mu
__global__ void lda_tables_kernel(const double* sums, int kp, int k, int p, int pp, const double* inv_cnt,
                                  const double* freq, double* means, double* Hb /* [pp cols][ldh rows] */, long ldh,
                                  double* mu) {
    for (int j = threadIdx.x; j < pp; j += blockDim.x) {
        double m = 0.0;
        for (int c = 0; c < k; ++c) {
            const double v = (j < p) ? sums[c + (long)j * kp] * inv_cnt[c] : 0.0;   // klda.jl:48
            means[c + (long)j * kp] = v;
            m += freq[c] * v;                                                          // klda.jl:51
        }
        for (int c = k; c < kp; ++c) means[c + (long)j * kp] = 0.0;
        mu[j] = m;
        for (int c = 0; c < k; ++c)                                                    // klda.jl:52-53
            Hb[(long)j * ldh + c] = (j < p) ? sqrt(freq[c]) * (means[c + (long)j * kp] - m) : 0.0;
    }
}
// sums[c][j] *= inv_cnt[c]  (class sums -> class means, kp x pp column-major)
__global__ void scale_class_rows_kernel(double* sums, int kp, int k, int pp, const double* inv_cnt) {
    for (int idx = threadIdx.x; idx < kp * pp; idx += blockDim.x) {
        const int c = idx % kp;
        sums[idx] = (c < k) ? sums[idx] * inv_cnt[c] : 0.0;
    }
}
// Y[:, j] = U[:, j] / sqrt(d_j); flags non-positive d (B not positive definite)
__global__ void scale_cols_kernel(const double* U, int ldu, const double* d, int p, double* Y, long ldy, int* bad) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *bad = 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < p * p; idx += blockDim.x) {
        const int i = idx % p, j = idx / p;
        double v = 0.0;
        if (d[j] > 0.0) v = U[i + (long)j * ldu] / sqrt(d[j]); else *bad = 1;
        Y[(long)j * ldy + i] = v;
    }
}

struct Lin {                       // device-resident X (ld x pp, column-major) and helpers
    int64_t n, p, ld; int pp;
    double *X, *ones, *scr; size_t scr_doubles;
};

int upload(fx_ctx* ctx, const double* X, int64_t n, int64_t p, int64_t ldx, Lin& L) {
    L.n = n; L.p = p; L.ld = round_up(n, 128); L.pp = (int)round_up(p, 8);
    L.X = (double*)fx_arena_alloc(ctx, (size_t)L.pp * L.ld * 8);
    L.ones = (double*)fx_arena_alloc(ctx, (size_t)8 * L.ld * 8);
    L.scr_doubles = fx_ts_tn_scratch(ctx, L.pp, L.pp, L.ld);
    size_t s2 = fx_ts_tn_scratch(ctx, 64, L.pp, L.ld);
    if (s2 > L.scr_doubles) L.scr_doubles = s2;
    size_t s3 = fx_ts_tn_scratch(ctx, L.pp, L.pp, 512);
    if (s3 > L.scr_doubles) L.scr_doubles = s3;
    L.scr = (double*)fx_arena_alloc(ctx, L.scr_doubles * 8);
    if (!L.X || !L.ones || !L.scr) { fx_set_error("internal: arena too small"); return FX_ERR_OOM; }
    FX_CUDA(cudaMemsetAsync(L.X, 0, (size_t)L.pp * L.ld * 8, ctx->stream));
    FX_CUDA(cudaMemcpy2DAsync(L.X, L.ld * 8, X, ldx * 8, n * 8, p, cudaMemcpyHostToDevice, ctx->stream));
    ones_kernel<<<(unsigned)ceil_div(L.ld, 256), 256, 0, ctx->stream>>>(L.ones, L.ld, n);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}
size_t lin_bytes(fx_ctx* ctx, int64_t n, int64_t p) {
    const int64_t ld = round_up(n, 128), pp = round_up(p, 8);
    size_t s = fx_ts_tn_scratch(ctx, (int)pp, (int)pp, ld);
    size_t s2 = fx_ts_tn_scratch(ctx, 64, (int)pp, ld);
    if (s2 > s) s = s2;
    size_t s3 = fx_ts_tn_scratch(ctx, (int)pp, (int)pp, 512);
    if (s3 > s) s = s3;
    return fx_align256((size_t)pp * ld * 8) + fx_align256((size_t)8 * ld * 8) + fx_align256(s * 8) +
           24 * fx_align256((size_t)(pp + 128) * (pp + 128) * 8) + 65536;
}

}  // namespace

extern "C" int fx_pca_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t p, int64_t ldx, double alpha,
                          double eps, double* V, int64_t ldv, double* lambda, double* mean, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_pca_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!X) { fx_set_error("fx_pca_f64: invalid argument 2 (!X)"); return -2; }
    if (n < 2) { fx_set_error("fx_pca_f64: invalid argument 3 (n < 2)"); return -3; }
    if (p < 1 || p > 512) { fx_set_error("pca: p must lie in 1..512"); return -4; }
    if (ldx < n) { fx_set_error("pca: ldx < n (DimensionMismatch)"); return -5; }
    if (!(alpha >= 0.0 && alpha <= 1.0)) { fx_set_error("fx_pca_f64: invalid argument 6 (!(alpha >= 0.0 && alpha <= 1.0))"); return -6; }
    if (!(eps >= 0.0)) { fx_set_error("fx_pca_f64: invalid argument 7 (!(eps >= 0.0))"); return -7; }
    if (!V) { fx_set_error("fx_pca_f64: invalid argument 8 (!V)"); return -8; }
    if (ldv < p) { fx_set_error("pca: ldv < p (DimensionMismatch)"); return -9; }
    if (!lambda) { fx_set_error("fx_pca_f64: invalid argument 10 (!lambda)"); return -10; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    FX_TRY(fx_arena_reserve(ctx, lin_bytes(ctx, n, p)));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    Lin L;
    FX_TRY(upload(ctx, X, n, p, ldx, L));
    const int pp = L.pp;
    auto dalloc = [&](size_t d) { return (double*)fx_arena_alloc(ctx, d * 8); };
    double* S8 = dalloc((size_t)pp * 8); double* M8 = dalloc((size_t)8 * pp); double* mu = dalloc(pp);
    double* Sig = dalloc((size_t)pp * pp); double* U = dalloc((size_t)pp * pp); double* Uw = dalloc((size_t)pp * pp);
    double* w = dalloc(pp + 8); double* tr = dalloc(8);
    int* info = (int*)fx_arena_alloc(ctx, 64);
    // column means (kpca.jl:19) and centring (kpca.jl:20)
    FX_TRY(fx_launch_ts_tn(ctx, L.X, pp, L.ones, 8, L.ld, L.ld, S8, pp, L.scr, L.scr_doubles, 0));
    neg_mean_kernel<<<1, 256, 0, s>>>(S8, pp, (int)p, pp, 1.0 / (double)n, M8, mu);
    ctx->launches++;
    FX_TRY(fx_launch_ts_nn(ctx, L.ones, 8, M8, 8, pp, 1.0, 1.0, L.X, L.ld, L.ld, 0));
    // Sigma = H'H (kpca.jl:9)
    FX_TRY(fx_launch_ts_tn(ctx, L.X, pp, L.X, pp, L.ld, L.ld, Sig, pp, L.scr, L.scr_doubles, 0));
    symmetrize_full_kernel<<<8, 256, 0, s>>>(Sig, pp, (int)p);
    regularize_kernel<<<1, 1024, 0, s>>>(Sig, pp, (int)p, 0.0, 0.0, tr);      // trace only
    ctx->launches += 2;
    FX_LAUNCH_CHECK();
    FX_TRY(fx_launch_jacobi(ctx, Sig, pp, (int)p, U, pp, w, Uw, info));
    std::vector<double> hw(p);
    double htr = 0.0;
    int hinfo[2] = {0, 0};
    FX_CUDA(cudaMemcpyAsync(hw.data(), w, p * 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(&htr, tr, 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(hinfo, info, 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpy2DAsync(V, ldv * 8, U, (size_t)pp * 8, p * 8, p, cudaMemcpyDeviceToHost, s));
    if (mean) FX_CUDA(cudaMemcpyAsync(mean, mu, p * 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    if (!hinfo[1]) { fx_set_error("pca: Jacobi did not converge"); return FX_ERR_NOT_CONVERGED; }
    // regularize!(S, alpha, tr/p); perturb!(S, eps (n-1)) (kpca.jl:10-11,21) only remap eigenvalues
    for (int64_t i = 0; i < p; ++i)
        lambda[i] = ((1.0 - alpha) * hw[i] + alpha * htr / (double)p + eps * (double)(n - 1)) / (double)(n - 1);
    st->trace = htr;
    st->converged = 1;
    st->dense_path = 1;
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

// out (m x k) = Z (m x p) * W (p x k)   -- transform_pca (kpca.jl:37) and transform_lda (klda.jl:65)
extern "C" int fx_project_f64(fx_ctx* ctx, const double* W, int64_t p, int k, int64_t ldw, const double* Z,
                              int64_t m, int64_t ldz, double* out, int64_t ldo, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_project_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!W) { fx_set_error("fx_project_f64: invalid argument 2 (!W)"); return -2; }
    if (p < 1 || p > 512) { fx_set_error("project: p must lie in 1..512"); return -3; }
    if (k < 1 || k > 512) { fx_set_error("fx_project_f64: invalid argument 4 (k < 1 || k > 512)"); return -4; }
    if (ldw < p) { fx_set_error("project: ldw < p (DimensionMismatch)"); return -5; }
    if (!Z) { fx_set_error("fx_project_f64: invalid argument 6 (!Z)"); return -6; }
    if (m < 1) { fx_set_error("fx_project_f64: invalid argument 7 (m < 1)"); return -7; }
    if (ldz < m) { fx_set_error("project: ldz < m (DimensionMismatch)"); return -8; }
    if (!out) { fx_set_error("fx_project_f64: invalid argument 9 (!out)"); return -9; }
    if (ldo < m) { fx_set_error("project: ldo < m (DimensionMismatch)"); return -10; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int64_t ld = round_up(m, 128);
    const int pp = (int)round_up(p, 8), kp = (int)round_up(k, 8);
    FX_TRY(fx_arena_reserve(ctx, fx_align256((size_t)pp * ld * 8) + fx_align256((size_t)kp * ld * 8) +
                                     fx_align256((size_t)pp * kp * 8) + 8192));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Zd = (double*)fx_arena_alloc(ctx, (size_t)pp * ld * 8);
    double* Od = (double*)fx_arena_alloc(ctx, (size_t)kp * ld * 8);
    double* Wd = (double*)fx_arena_alloc(ctx, (size_t)pp * kp * 8);
    FX_CUDA(cudaMemsetAsync(Zd, 0, (size_t)pp * ld * 8, s));
    FX_CUDA(cudaMemsetAsync(Wd, 0, (size_t)pp * kp * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(Zd, ld * 8, Z, ldz * 8, m * 8, p, cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpy2DAsync(Wd, (size_t)pp * 8, W, ldw * 8, p * 8, k, cudaMemcpyHostToDevice, s));
    FX_TRY(fx_launch_ts_nn(ctx, Zd, pp, Wd, pp, kp, 1.0, 0.0, Od, ld, ld, 0));
    FX_CUDA(cudaMemcpy2DAsync(out, ldo * 8, Od, ld * 8, m * 8, k, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

// Device core of lda! (klda.jl:34-61) on a data matrix that already sits in HBM: Xd is ld x pp
// column-major (pp = roundup(p, 8) columns, zero padded; DESTROYED: it becomes H_w).  Used by fx_lda_f64
// and by the dense path of fx_klda_fit_f64, where the data matrix is the kernel matrix itself
// (SURVEY.md A.3).  Buffers come from the arena (fx_lda_core_bytes).  Leaves the p x p generalized
// eigenvectors (leading dimension *ldv_out, V' Sigma_w' V = I, descending) and eigenvalues on the
// device; hinfo[1], hinfo[3] = Jacobi convergence flags, hinfo[4] = Sigma_w' not positive definite.
size_t fx_lda_core_bytes(fx_ctx* ctx, int64_t n, int64_t p) {
    const int64_t ld = round_up(n, 128), pp = round_up(p, 8), ldy = round_up(pp, 128);
    size_t s = fx_ts_tn_scratch(ctx, (int)pp, (int)pp, ld);
    size_t s2 = fx_ts_tn_scratch(ctx, 64, (int)pp, ld);
    if (s2 > s) s = s2;
    size_t s3 = fx_ts_tn_scratch(ctx, (int)pp, (int)pp, ldy);
    if (s3 > s) s = s3;
    return fx_align256(s * 8) + fx_align256((size_t)64 * ld * 8) + 10 * fx_align256((size_t)pp * ldy * 8) +
           fx_align256((size_t)ld * 4) + 65536;
}

// *out (device) = tr(H_w'H_w) with H_w = (X - M[y]) / (n - k) (klda.jl:49-50); Xd (ld x pp) is destroyed.
__global__ void sumsq_kernel(const double* __restrict__ X, long total, double scale, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (long i = threadIdx.x; i < total; i += blockDim.x) s = fma(X[i], X[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) *out = t * scale;
    }
}
int fx_lda_hw_trace(fx_ctx* ctx, double* Xd, int64_t n, int64_t p, int64_t ld, const int64_t* y, int k,
                    const double* inv_cnt, double* out) {
    cudaStream_t s = ctx->stream;
    const int pp = (int)round_up(p, 8), kp = (int)round_up(k, 8);
    auto dalloc = [&](size_t d) { return (double*)fx_arena_alloc(ctx, d * 8); };
    const size_t scr_doubles = fx_ts_tn_scratch(ctx, 64, pp, ld);
    double* scr = dalloc(scr_doubles);
    double* E = dalloc((size_t)kp * ld);
    double* sums = dalloc((size_t)kp * pp);
    double* dinv = dalloc(64);
    int* ydev = (int*)fx_arena_alloc(ctx, (size_t)ld * 4);
    if (!scr || !E || !sums || !dinv || !ydev) { fx_set_error("internal: arena too small for tr(Sigma_w)"); return FX_ERR_OOM; }
    {
        std::vector<int> y32(ld, 0);
        for (int64_t i = 0; i < n; ++i) y32[i] = (int)y[i];
        FX_CUDA(cudaMemcpyAsync(ydev, y32.data(), (size_t)ld * 4, cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaMemcpyAsync(dinv, inv_cnt, k * 8, cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaStreamSynchronize(s));
    }
    FX_TRY(fx_launch_indicator(ctx, ydev, n, ld, kp, E));
    FX_TRY(fx_launch_ts_tn(ctx, E, kp, Xd, pp, ld, ld, sums, kp, scr, scr_doubles, 0));
    scale_class_rows_kernel<<<1, 256, 0, s>>>(sums, kp, k, pp, dinv);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    FX_TRY(fx_launch_ts_nn(ctx, E, kp, sums, kp, pp, -1.0, 1.0, Xd, ld, ld, 0));          // X -= E * means
    const double sc = 1.0 / ((double)(n - k) * (double)(n - k));
    sumsq_kernel<<<1, 1024, 0, s>>>(Xd, (long)pp * ld, sc, out);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_lda_core(fx_ctx* ctx, double* Xd, int64_t n, int64_t p, int64_t ld, const int64_t* y, int k,
                const double* freq, const double* inv_cnt, double alpha_b, double eps_b, double alpha_w,
                double eps_w, double** Vd_out, int64_t* ldv_out, double** lam_out, int hinfo[8]) {
    cudaStream_t s = ctx->stream;
    const int pp = (int)round_up(p, 8), kp = (int)round_up(k, 8);
    const int64_t ldy = round_up(pp, 128);                     // p x p operands padded to 128 rows for tsgemm
    auto dalloc = [&](size_t d) { return (double*)fx_arena_alloc(ctx, d * 8); };
    size_t scr_doubles = fx_ts_tn_scratch(ctx, pp, pp, ld);
    {
        size_t s2 = fx_ts_tn_scratch(ctx, 64, pp, ld), s3 = fx_ts_tn_scratch(ctx, pp, pp, ldy);
        if (s2 > scr_doubles) scr_doubles = s2;
        if (s3 > scr_doubles) scr_doubles = s3;
    }
    double* scr = dalloc(scr_doubles);
    double* E = dalloc((size_t)kp * ld);
    double* sums = dalloc((size_t)kp * pp); double* means = dalloc((size_t)kp * pp); double* mu = dalloc(pp);
    const int64_t ldh = 128;                                   // H_b stored as (128 rows x pp), rows >= k zero
    double* Hb = dalloc((size_t)pp * ldh);
    double* Sb = dalloc((size_t)pp * pp); double* Sw = dalloc((size_t)pp * pp);
    double* U = dalloc((size_t)pp * pp); double* Uw = dalloc((size_t)pp * pp); double* dvals = dalloc(pp + 8);
    double* Y = dalloc((size_t)pp * ldy); double* T = dalloc((size_t)pp * ldy); double* SbP = dalloc((size_t)pp * ldy);
    double* Mm = dalloc((size_t)pp * pp); double* Zv = dalloc((size_t)pp * pp); double* lam = dalloc(pp + 8);
    double* Vd = dalloc((size_t)pp * ldy);
    double* dfreq = dalloc(64); double* dinv = dalloc(64);
    int* ydev = (int*)fx_arena_alloc(ctx, (size_t)ld * 4);
    int* info = (int*)fx_arena_alloc(ctx, 64);
    if (!info || !Vd || !scr) { fx_set_error("internal: arena too small for the LDA core"); return FX_ERR_OOM; }
    {
        std::vector<int> y32(ld, 0);
        for (int64_t i = 0; i < n; ++i) y32[i] = (int)y[i];
        FX_CUDA(cudaMemcpyAsync(ydev, y32.data(), (size_t)ld * 4, cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaMemcpyAsync(dfreq, freq, k * 8, cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaMemcpyAsync(dinv, inv_cnt, k * 8, cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaStreamSynchronize(s));
    }
    FX_TRY(fx_launch_indicator(ctx, ydev, n, ld, kp, E));
    // class sums -> means, mu, H_b (klda.jl:47-53)
    FX_TRY(fx_launch_ts_tn(ctx, E, kp, Xd, pp, ld, ld, sums, kp, scr, scr_doubles, 0));
    FX_CUDA(cudaMemsetAsync(Hb, 0, (size_t)pp * ldh * 8, s));
    lda_tables_kernel<<<1, 256, 0, s>>>(sums, kp, k, (int)p, pp, dinv, dfreq, means, Hb, ldh, mu);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    // H_w = (X - M[y]) / (n - k)  (klda.jl:49-50):  X -= E * means, then scale
    FX_TRY(fx_launch_ts_nn(ctx, E, kp, means, kp, pp, -1.0, 1.0, Xd, ld, ld, 0));
    FX_TRY(fx_launch_vec_axpy(ctx, (int64_t)pp * ld, 1.0 / (double)(n - k) - 1.0, Xd, Xd));
    // Sigma_b = H_b'H_b, Sigma_w = H_w'H_w (klda.jl:35,39) with shrink / perturb (klda.jl:36-41)
    FX_TRY(fx_launch_ts_tn(ctx, Hb, pp, Hb, pp, ldh, ldh, Sb, pp, scr, scr_doubles, 0));
    FX_TRY(fx_launch_ts_tn(ctx, Xd, pp, Xd, pp, ld, ld, Sw, pp, scr, scr_doubles, 0));
    symmetrize_full_kernel<<<8, 256, 0, s>>>(Sb, pp, (int)p);          // syml! (klda.jl:35,39)
    symmetrize_full_kernel<<<8, 256, 0, s>>>(Sw, pp, (int)p);
    regularize_kernel<<<1, 1024, 0, s>>>(Sb, pp, (int)p, alpha_b, eps_b, nullptr);
    regularize_kernel<<<1, 1024, 0, s>>>(Sw, pp, (int)p, alpha_w, eps_w, nullptr);
    ctx->launches += 4;
    FX_LAUNCH_CHECK();
    // generalized symmetric eigenproblem (common.jl:168-174): Sigma_w = U D U'
    FX_CUDA(cudaMemsetAsync(SbP, 0, (size_t)pp * ldy * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(SbP, ldy * 8, Sb, (size_t)pp * 8, p * 8, p, cudaMemcpyDeviceToDevice, s));
    FX_TRY(fx_launch_jacobi(ctx, Sw, pp, (int)p, U, pp, dvals, Uw, info));
    FX_CUDA(cudaMemsetAsync(Y, 0, (size_t)pp * ldy * 8, s));
    scale_cols_kernel<<<1, 1024, 0, s>>>(U, pp, dvals, (int)p, Y, ldy, info + 4);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    // M = Y' Sigma_b Y
    FX_CUDA(cudaMemsetAsync(T, 0, (size_t)pp * ldy * 8, s));
    FX_TRY(fx_launch_ts_nn(ctx, SbP, pp, Y, (int)ldy, pp, 1.0, 0.0, T, ldy, ldy, 0));
    FX_TRY(fx_launch_ts_tn(ctx, Y, pp, T, pp, ldy, ldy, Mm, pp, scr, scr_doubles, 0));
    symmetrize_full_kernel<<<8, 256, 0, s>>>(Mm, pp, (int)p);
    ctx->launches++;
    FX_TRY(fx_launch_jacobi(ctx, Mm, pp, (int)p, Zv, pp, lam, Uw, info + 2));
    // V = Y Z
    FX_CUDA(cudaMemsetAsync(Vd, 0, (size_t)pp * ldy * 8, s));
    FX_TRY(fx_launch_ts_nn(ctx, Y, pp, Zv, pp, pp, 1.0, 0.0, Vd, ldy, ldy, 0));
    FX_CUDA(cudaMemcpyAsync(hinfo, info, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    *Vd_out = Vd; *ldv_out = ldy; *lam_out = lam;
    return 0;
}

extern "C" int fx_lda_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t p, int64_t ldx, const int64_t* y,
                          int nclass, const double* freq, double alpha_b, double eps_b, double alpha_w,
                          double eps_w, double* V, int64_t ldv, double* lambda, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_lda_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!X) { fx_set_error("fx_lda_f64: invalid argument 2 (!X)"); return -2; }
    if (n < 2) { fx_set_error("fx_lda_f64: invalid argument 3 (n < 2)"); return -3; }
    if (p < 1 || p > 512) { fx_set_error("lda: p must lie in 1..512"); return -4; }
    if (ldx < n) { fx_set_error("lda: ldx < n (DimensionMismatch)"); return -5; }
    if (!y) { fx_set_error("fx_lda_f64: invalid argument 6 (!y)"); return -6; }
    const int k = nclass;
    if (k < 2 || k > 64 || k >= n) { fx_set_error("lda: nclass must lie in 2..64 and be < n"); return -7; }
    if (!freq) { fx_set_error("fx_lda_f64: invalid argument 8 (!freq)"); return -8; }
    if (!(alpha_b >= 0.0 && alpha_b <= 1.0)) { fx_set_error("fx_lda_f64: invalid argument 9 (!(alpha_b >= 0.0 && alpha_b <= 1.0))"); return -9; }
    if (!(eps_b >= 0.0)) { fx_set_error("fx_lda_f64: invalid argument 10 (!(eps_b >= 0.0))"); return -10; }
    if (!(alpha_w >= 0.0 && alpha_w <= 1.0)) { fx_set_error("fx_lda_f64: invalid argument 11 (!(alpha_w >= 0.0 && alpha_w <= 1.0))"); return -11; }
    if (!(eps_w >= 0.0)) { fx_set_error("fx_lda_f64: invalid argument 12 (!(eps_w >= 0.0))"); return -12; }
    if (!V) { fx_set_error("fx_lda_f64: invalid argument 13 (!V)"); return -13; }
    if (ldv < p) { fx_set_error("lda: ldv < p (DimensionMismatch)"); return -14; }
    if (!lambda) { fx_set_error("fx_lda_f64: invalid argument 15 (!lambda)"); return -15; }
    std::vector<double> inv_cnt(k, 0.0);
    {
        std::vector<int64_t> cnt(k, 0);
        for (int64_t i = 0; i < n; ++i) {
            if (y[i] < 1 || y[i] > k) { fx_set_error("lda: Index %lld out of range.", (long long)(i + 1)); return -6; }
            cnt[y[i] - 1]++;                                   // klda.jl:5-12
        }
        for (int c = 0; c < k; ++c) {
            if (!cnt[c]) { fx_set_error("lda: class %d is empty", c + 1); return -6; }
            inv_cnt[c] = 1.0 / (double)cnt[c];
        }
    }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    FX_TRY(fx_arena_reserve(ctx, lin_bytes(ctx, n, p) + fx_lda_core_bytes(ctx, n, p)));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    Lin L;
    FX_TRY(upload(ctx, X, n, p, ldx, L));
    double* Vd = nullptr; double* lam = nullptr; int64_t ldy = 0;
    int hinfo[8] = {0};
    FX_TRY(fx_lda_core(ctx, L.X, n, p, L.ld, y, k, freq, inv_cnt.data(), alpha_b, eps_b, alpha_w, eps_w, &Vd, &ldy,
                       &lam, hinfo));
    FX_CUDA(cudaMemcpy2DAsync(V, ldv * 8, Vd, ldy * 8, p * 8, p, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(lambda, lam, p * 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    st->kernel_launches = ctx->launches - launches0;
    st->dense_path = 1;
    if (hinfo[4]) {
        fx_set_error("lda: Sigma_w is not positive definite (the analogue of sygvd info > n); pass eps_w > 0");
        return FX_ERR_NOT_POSDEF;
    }
    if (!hinfo[1] || !hinfo[3]) { fx_set_error("lda: Jacobi did not converge"); return FX_ERR_NOT_CONVERGED; }
    st->converged = 1;
    return 0;
}
