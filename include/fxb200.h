/* fxb200.h -- C ABI of the B200-native Kernel PCA / Kernel LDA engine.
 *
 * Drop-in boundary for the solvers of trthatcher/FeatureExtraction.jl
 * (reference paths are relative to the upstream repository root):
 *
 *   fx_kpca_fit_f64        replaces  kpca(X, k, alpha, eps)            src/kpca.jl:49-57
 *                          i.e. MLKernels.kernelmatrix (kpca.jl:51), centerkernelmatrix!
 *                          (kpca.jl:52), regularize!/perturb! (kpca.jl:53-54, common.jl:32-50),
 *                          components_eig! -> syevd! -> ccall dsyevd_ (common.jl:161-165,
 *                          :107-143, ccall at :128-131) and scale!(L, 1/(n-1)) (kpca.jl:56)
 *   fx_kpca_transform_f64  replaces  transform_kpca(X, k, W, Z)        src/kpca.jl:64-68
 *   fx_klda_fit_f64        replaces  lda_matrices!/lda_components!/lda! applied to K
 *                          (klda.jl:34-61; `klda!` is exported at FeatureExtraction.jl:11
 *                          but never defined)
 *   fx_klda_transform_f64  replaces  transform_klda(X, k, W, Z)        src/klda.jl:71
 *   fx_pca_f64, fx_lda_f64, fx_project_f64   the linear pca / lda! / transform_pca / transform_lda
 *                          (src/kpca.jl:8-37, src/klda.jl:34-65) on the same kernels, p <= 512
 *   fx_gram_f64            = MLKernels.kernelmatrix(k, X) / (k, Z, X)  kpca.jl:51,65
 *   fx_center_f64          = MLKernels.centerkernelmatrix!(K)          kpca.jl:52
 *   fx_symm_f64, fx_syevd_small_f64   the two building blocks that stand in for the
 *                          dsyevd_ call (common.jl:128-131): K*V and the small dense
 *                          Rayleigh-Ritz eigensolver -- exposed for parity tests
 *
 * Conventions (mirroring the reference's only FFI precedent, the LAPACK ccall at
 * common.jl:118-141): every matrix is column-major double with an explicit leading
 * dimension, sizes are 64-bit, outputs are caller-allocated, the return value is an
 * `info` code: 0 ok; -i = argument i invalid (1-based, as LAPACK); >0 runtime class:
 *   1 CUDA error, 2 NCCL error, 3 out of memory, 4 eigensolver not converged,
 *   5 matrix not positive definite / CG breakdown.  fx_last_error() gives the text.
 * The library never keeps a host pointer after a call returns and never aborts.
 * There is no CPU fallback: every entry point runs hand-written sm_100a kernels.
 *
 * Host entry points take HOST pointers; the *_dev_* variants take DEVICE pointers
 * (inputs already resident in HBM) and run on the context's stream.
 */
#ifndef FXB200_H
#define FXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fx_ctx fx_ctx;       /* one GPU (one rank); not re-entrant */
typedef struct fx_model fx_model;   /* fitted state living in HBM */

/* kernel ids; kparams[3] meaning:
 *   GAUSSIAN   exp(-p0 * ||x-y||^2)                       (p1, p2 ignored)
 *   POLYNOMIAL (p0 * x'y + p1)^p2   -- PolynomialKernel(a,c,d), test_kpca.jl:49
 *   SIGMOID    tanh(p0 * x'y + p1)                         (p2 ignored)            */
enum { FX_KERNEL_GAUSSIAN = 0, FX_KERNEL_POLYNOMIAL = 1, FX_KERNEL_SIGMOID = 2 };

enum { FX_OK = 0, FX_ERR_CUDA = 1, FX_ERR_NCCL = 2, FX_ERR_OOM = 3,
       FX_ERR_NOT_CONVERGED = 4, FX_ERR_NOT_POSDEF = 5 };

/* Eigensolver / CG options.  Zero in any field selects the default. */
typedef struct fx_solver_opts {
    int32_t block;        /* Krylov block width b (default: k rounded up to 8, max 64)      */
    int32_t max_blocks;   /* basis capacity in blocks before a thick restart (0: 5, 4 or 3 by K size) */
    int32_t max_steps;    /* K*V passes allowed (default 300)                                */
    int32_t variant;      /* 0 = tuned kernels, 1 = naive CUDA kernels (debug cross-check)   */
    double  tol;          /* Ritz residual tolerance relative to |theta_1| (default 1e-12)   */
    uint64_t seed;        /* start-block seed (default 0x5eed)                               */
    int32_t cg_max_iters; /* KLDA: CG iterations allowed (default 4000)                      */
    int32_t klda_path;    /* KLDA: 0 = automatic (dense pipeline for n <= 512 on one GPU), 1 = force the iterative path */
    double  cg_tol;       /* KLDA: CG relative residual (default 1e-10)                      */
} fx_solver_opts;

/* Per-call device timings (CUDA events on the context stream) and counters. */
typedef struct fx_stats {
    double total_ms;      /* whole call on the device stream, copies included               */
    double h2d_ms, d2h_ms;
    double gram_ms;       /* Gram + fused kernel-map epilogue (sum over launches)           */
    double center_ms;     /* centering pass (sum over launches)                             */
    double symm_ms;       /* K*V passes (sum over launches)                                 */
    double ortho_ms;      /* tall-skinny projections / orthonormalisation                   */
    double rr_ms;         /* small dense Jacobi (Rayleigh-Ritz + SVQB)                      */
    double comm_ms;       /* NCCL collectives                                               */
    double eig_ms;        /* whole eigensolve                                               */
    double transform_ms;  /* fused transform kernel                                         */
    double max_residual;  /* max relative Ritz residual at exit                             */
    double trace;         /* trace of the (centred) kernel matrix                           */
    int64_t gram_launches, center_launches, symm_launches, transform_launches;
    int64_t kernel_launches;   /* every kernel this call launched                           */
    int64_t h2d_bytes, d2h_bytes;
    int32_t steps;        /* K*V passes used                                                */
    int32_t restarts;
    int32_t converged;
    int32_t block;        /* block width used                                               */
    int32_t cg_iters;
    int32_t dense_path;   /* 1 if the small-n dense Jacobi path ran                         */
} fx_stats;

int         fx_version(void);
const char* fx_last_error(void);

/* ---- contexts ------------------------------------------------------------------- */
int  fx_init(int device, fx_ctx** out);
/* Multi-GPU: one process per GPU.  Rank 0 calls fx_comm_unique_id, the host side ships
 * the 128 bytes to every rank (torch.distributed / MPI / Julia Distributed), then every
 * rank calls fx_init_distributed.  K is sharded by row blocks; X is replicated. */
int  fx_comm_unique_id(void* id128);
int  fx_init_distributed(int device, int rank, int world, const void* id128, fx_ctx** out);
/* Multi-GPU in ONE process: the context drives `ndev` GPUs of this node with one host
 * thread per device; NCCL communicators and NVLink peer access are set up inside the library, the
 * caller makes one call per fit / transform with host buffers and never sees ranks.  fit shards K
 * by row blocks, transform shards the queries; the building blocks and the linear solvers run on
 * devices[0].  Device-pointer entry points (*_dev_*) and fx_set_stream need a single-device context. */
int  fx_init_multi(const int* devices, int ndev, fx_ctx** out);
int  fx_set_stream(fx_ctx* ctx, void* cuda_stream);     /* run on a caller-owned stream    */
int  fx_ctx_info(const fx_ctx* ctx, int* device, int* rank, int* world, int* sm_count);
/* 1 if the last collective fit reached every other rank's workspace over NVLink peer mappings
 * (CUDA IPC; symmetric blocks of K shared out between ranks), 0 if each rank computed its whole
 * row block (one rank, peer access unavailable, or FXB200_NO_PEER=1).  Results are identical. */
int  fx_ctx_peer_access(const fx_ctx* ctx);
void fx_shutdown(fx_ctx* ctx);

/* ---- KPCA ------------------------------------------------------------------------ */
/* X: n x d (rows = observations, kpca.jl:50).  V: n x k, unit 2-norm columns, sign
 * free.  lambda: k values, descending, already ((1-alpha) l + alpha tr/n + eps (n-1))/(n-1).
 * center != 0 reproduces the reference (kpca.jl:52).  model may be NULL.             */
int fx_kpca_fit_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                    int kernel_id, const double* kparams, int k, int center,
                    double alpha, double eps, const fx_solver_opts* opts,
                    double* V, int64_t ldv, double* lambda,
                    fx_model** model, fx_stats* stats);
int fx_kpca_fit_dev_f64(fx_ctx* ctx, const double* dX, int64_t n, int64_t d, int64_t ldx,
                        int kernel_id, const double* kparams, int k, int center,
                        double alpha, double eps, const fx_solver_opts* opts,
                        double* dV, int64_t ldv, double* dlambda,
                        fx_model** model, fx_stats* stats);
/* out: m x k = (k(Z,X) - 1 r' - r_z 1' + mu) W  (full double-centring formula).  For
 * components orthogonal to 1 (every eigenvector of the centred matrix with non-zero eigenvalue)
 * this is k(Z,X) W - 1 (r'W) and only the k-vector bias is applied; otherwise the kernel also
 * accumulates the per-query row means r_z.  Any k (launches of <= 64 components). */
int fx_kpca_transform_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                          int64_t ldz, double* out, int64_t ldo, fx_stats* stats);
int fx_kpca_transform_dev_f64(fx_ctx* ctx, const fx_model* model, const double* dZ, int64_t m,
                              int64_t ldz, double* dout, int64_t ldo, fx_stats* stats);
/* Same call through the naive one-thread-per-output CUDA kernel (debug cross-check). */
int fx_kpca_transform_naive_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                                int64_t ldz, double* out, int64_t ldo, fx_stats* stats);
/* Stateless route of the reference signature transform_kpca(X, k, W, Z): rebuild the
 * model from (X, kernel, W).  center != 0 recomputes the bias r'W on the device.     */
int fx_model_create_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                        int kernel_id, const double* kparams, const double* W, int k,
                        int64_t ldw, int center, fx_model** model, fx_stats* stats);
int  fx_model_info(const fx_model* model, int64_t* n, int64_t* d, int* k, int* kernel_id);
void fx_model_free(fx_model* model);

/* ---- KLDA ------------------------------------------------------------------------ */
/* y: 1-based labels in 1..nclass (klda.jl:8).  A: n x ncomp generalized eigenvectors of
 * (Sigma_b, Sigma_w), a' Sigma_w a = 1 (sygvd itype 1, common.jl:172); lambda descending. */
int fx_klda_fit_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                    int kernel_id, const double* kparams, const int64_t* y, int nclass,
                    const double* freq, double alpha_b, double eps_b, double alpha_w,
                    double eps_w, int ncomp, const fx_solver_opts* opts,
                    double* A, int64_t lda, double* lambda, fx_model** model, fx_stats* stats);
int fx_klda_transform_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                          int64_t ldz, double* out, int64_t ldo, fx_stats* stats);

/* ---- linear solvers of the same export list (p <= 512), on the same device kernels ---- */
/* pca(X, alpha, eps) (kpca.jl:17-31): V p x p (descending), lambda p (already / (n-1)),
 * mean (p, may be NULL) = the column means removed before the decomposition.            */
int fx_pca_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t p, int64_t ldx, double alpha,
               double eps, double* V, int64_t ldv, double* lambda, double* mean, fx_stats* stats);
/* lda!(X, y, k, frequency, ...) (klda.jl:57-61) with the intended pencil (Sigma_b, Sigma_w):
 * V p x p generalized eigenvectors, V' Sigma_w V = I (sygvd itype 1), lambda descending.   */
int fx_lda_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t p, int64_t ldx, const int64_t* y,
               int nclass, const double* freq, double alpha_b, double eps_b, double alpha_w,
               double eps_w, double* V, int64_t ldv, double* lambda, fx_stats* stats);
/* transform_pca(W, Z) (kpca.jl:37) / transform_lda(W, Z) (klda.jl:65): out (m x k) = Z W.   */
int fx_project_f64(fx_ctx* ctx, const double* W, int64_t p, int k, int64_t ldw, const double* Z,
                   int64_t m, int64_t ldz, double* out, int64_t ldo, fx_stats* stats);

/* ---- building blocks, exposed for parity tests and per-kernel benches -------------- */
/* K (m x n) = k(Z, X); Z == NULL means the symmetric case K = k(X, X) (m = n). */
int fx_gram_f64(fx_ctx* ctx, const double* Z, int64_t m, int64_t ldz,
                const double* X, int64_t n, int64_t ldx, int64_t d,
                int kernel_id, const double* kparams, int variant,
                double* K, int64_t ldk, fx_stats* stats);
/* In-place symmetric double centring of a host matrix; colmean (n) and grandmean out. */
int fx_center_f64(fx_ctx* ctx, double* K, int64_t n, int64_t ldk,
                  double* colmean, double* grandmean, fx_stats* stats);
/* Y (n x b) = K (n x n, symmetric) * V (n x b) through the SYMM kernel. */
int fx_symm_f64(fx_ctx* ctx, const double* K, int64_t n, int64_t ldk,
                const double* V, int b, int64_t ldv, double* Y, int64_t ldy,
                int variant, fx_stats* stats);
/* The centring pass as a FIT runs it: K (n x n symmetric, host) is centred in place by the fused
 * kernel that also multiplies, Y (n x b) = K~ V.  Same result as fx_center_f64 + fx_symm_f64;
 * exposed so that the fused write-back is tested element by element (kpca.jl:52). */
int fx_center_symm_f64(fx_ctx* ctx, double* K, int64_t n, int64_t ldk, const double* V, int b,
                       int64_t ldv, double* Y, int64_t ldy, double* colmean, double* grandmean,
                       fx_stats* stats);
/* Device Jacobi eigensolver: A (m x m symmetric) -> w descending, A overwritten by
 * eigenvectors (columns).  m <= 1024. */
int fx_syevd_small_f64(fx_ctx* ctx, double* A, int m, int64_t lda, double* w, int* sweeps);

/* ---- device-resident benches (inputs generated / kept in HBM) ---------------------- */
/* Times `reps` launches of one kernel class with CUDA events on the context stream:
 *   which = 0 Gram(+epilogue), 1 centering, 2 K*V, 3 fused transform, 4 DMMA issue-rate
 *   microbenchmark (register-only), 5 DFMA issue-rate microbenchmark, 6 stream copy,
 *           7 DMMA issue sweep (n = CTAs/SM, b = warps/CTA, d = accumulators/warp, m = iters),
 *           8 DMMA/DFMA datapath sharing probe (d = 0 DMMA only, 1 DFMA only, 2 both).
 * Sizes: n rows, d features, b block width / components, m queries.  ms_out[reps].  */
int fx_bench_kernel(fx_ctx* ctx, int which, int64_t n, int64_t d, int b, int64_t m,
                    int kernel_id, int reps, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* FXB200_H */
