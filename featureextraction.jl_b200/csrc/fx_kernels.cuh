// fx_kernels.cuh -- host-side launchers of the hand-written sm_100a kernels.
#pragma once
#include <functional>
#include "fx_common.cuh"

// ---- api_core.cu: single-process multi-GPU fan-out -----------------------------------------------
int fx_multi_run(fx_ctx* parent, const std::function<int(fx_ctx*, int)>& fn);
// building blocks and the linear solvers run on one GPU: a multi-GPU parent lends its first device
static inline fx_ctx* fx_single(fx_ctx* ctx) { return (ctx && !ctx->sub.empty()) ? ctx->sub[0] : ctx; }

// ---- gram.cu ---------------------------------------------------------------------------
// K[i][j] = map(A_i . B_j) for i in [0, rows_pad), j in [0, ldk); operands are feature-major
// (column-major observations x features): element (obs, f) at P[f * ld + obs].
struct GramArgs {
    const double* A; int64_t lda;       // row operand, already offset to the first local row
    const double* B; int64_t ldb;       // column operand
    const double* na; const double* nb; // squared norms (Gaussian), same offsets
    int64_t d_pad;                      // multiple of 16
    double* K; int64_t ldk;             // out, row-major [rows_pad][ldk], ldk % 64 == 0
    int64_t rows_pad;                   // multiple of 128
    int64_t rows_valid, cols_valid;     // entries outside are written as 0
    int64_t diag_shift;                 // global index of local row 0 (symmetric case), else -1
    double* rowpart;                    // [fx_gram_rowpart_slots(ldk)][rows_pad] per-tile partial row sums or null
    int kernel_id; double kp[3];
    int sym_source = 0;                 // 1: A is the row block [diag_shift, +rows_pad) of B (K symmetric)
    int no_symmetry = 0;                // 1: compute every tile even when the matrix is symmetric
    // Cross-rank symmetry (sym_source only): with peer_world > 1 every rank launches the same
    // call on its own row block; K / rowpart of rank c are reachable at peerK[c] / peerRowpart[c]
    // (CUDA IPC mappings; the own entries are K / rowpart).  Each rank computes about half of its
    // row block and stores the transposed tiles into the owning peer.  The caller provides the
    // cross-rank barriers before (peer buffers free) and after (mirrored tiles landed).
    int peer_world = 0, peer_rank = 0;
    double* peerK[16] = {};
    double* peerRowpart[16] = {};
};
int fx_launch_rownorms(fx_ctx* ctx, const double* P, int64_t ld, int64_t nobs, int64_t d_pad,
                       double* norms);
int fx_launch_gram(fx_ctx* ctx, const GramArgs& a, int variant);
// slots of GramArgs::rowpart per row: one partial row sum per 64 columns
int64_t fx_gram_rowpart_slots(int64_t ldk);

// ---- center.cu -------------------------------------------------------------------------
// rowsum[i] = sum over tiles of rowpart[t][i]
int fx_launch_rowpart_reduce(fx_ctx* ctx, const double* rowpart, int64_t ntiles,
                             int64_t rows_pad, double* rowsum);
// rowsum[i] = sum_j K[i][j] (j < cols_valid) for a row-major matrix (standalone centring)
int fx_launch_rowsums(fx_ctx* ctx, const double* K, int64_t ldk, int64_t rows, int64_t cols_valid,
                      double* rowsum);
// colmean[j] = sums[j] / n ; *grand = sum(sums[0..n)) / n^2 ; entries >= n zeroed
int fx_launch_means(fx_ctx* ctx, const double* sums, int64_t n, int64_t n_ld, double* colmean,
                    double* grand);
// K[i][j] += grand - colmean[row0 + i] - colmean[j]   (valid entries only)
int fx_launch_center(fx_ctx* ctx, double* K, int64_t ldk, int64_t rows_valid, int64_t cols_valid,
                     int64_t row0, const double* colmean, const double* grand);
// *out = sum_i K[i][row0 + i], i < rows_valid
int fx_launch_trace(fx_ctx* ctx, const double* K, int64_t ldk, int64_t rows_valid, int64_t row0,
                    double* out);

// ---- symm.cu ---------------------------------------------------------------------------
// Y[c][i] = sum_j K[i][j] V[j][c];  K row-major [rows_pad][ldk] (ldk % 64 == 0, rows_pad % 128
// == 0); V is the all-gathered block: rank g holds rows [g*ldr, (g+1)*ldr) at
// V + g*v_rank_stride, column c at + c*ldr.  Y column-major with leading dim rows_pad.
struct SymmArgs {
    const double* K; int64_t ldk; int64_t rows_pad;
    const double* V; int64_t ldr; int64_t v_rank_stride;
    int bcols;                          // multiple of 8, <= 64
    double* Y;                          // [bcols][rows_pad]
    double* Ypart;                      // scratch >= splits*bcols*rows_pad doubles (splits>1)
    int splits;                         // k-splits chosen by fx_symm_splits
    // Fused centring (first pass of a fit): when center_colmean != null, K still holds the
    // uncentred matrix; the kernel centres every element it streams, writes it back IN PLACE
    // (K is then an in/out argument) and multiplies with the centred values.
    const double* center_colmean = nullptr;   // [ldk] column means (index = global sample)
    const double* center_grand = nullptr;     // device scalar
    int64_t rows_valid = 0, cols_valid = 0, row0 = 0;
};
int fx_symm_splits(fx_ctx* ctx, int64_t rows_pad, int64_t ldk);
int fx_launch_symm(fx_ctx* ctx, const SymmArgs& a, int variant);

// ---- tsgemm.cu -------------------------------------------------------------------------
// Tall-skinny products on column-major matrices with leading dimension ld (rows % 128 == 0).
// H[ma x nb] (ldh) = A[:, :ma]' * B[:, :nb]      (ma, nb multiples of 8)
int fx_launch_ts_tn(fx_ctx* ctx, const double* A, int ma, const double* B, int nb, int64_t ld,
                    int64_t rows, double* H, int ldh, double* scratch, size_t scratch_doubles,
                    int variant);
size_t fx_ts_tn_scratch(fx_ctx* ctx, int ma, int nb, int64_t rows);
// C[:, :nc] = beta*C + alpha * A[:, :ma] * M[ma x nc] (ldm);  in place (C == A) allowed when
// ma == nc.  ma, nc multiples of 8.
int fx_launch_ts_nn(fx_ctx* ctx, const double* A, int ma, const double* M, int ldm, int nc,
                    double alpha, double beta, double* C, int64_t ld, int64_t rows, int variant);

// ---- smalldense.cu -----------------------------------------------------------------------
// Parallel cyclic Jacobi on a device matrix A (m x m, lda), eigenvalues descending in w,
// eigenvectors (columns, same order) in Vout (ldv).  A is destroyed.  info[0] = sweeps,
// info[1] = converged flag.
// `need` (device int, may be null): the kernel returns at once when *need == 0.
int fx_launch_jacobi(fx_ctx* ctx, double* A, int lda, int m, double* Vout, int ldv, double* w,
                     double* Vwork, int* info, const int* need = nullptr);
// CholQR fast path of the block orthonormalisation; sets *need_svqb = 1 when a pivot < thr
int fx_launch_chol_factors(fx_ctx* ctx, const double* G, int ldg, int b, double thr, double* M, double* Cfac,
                           int* need_svqb, int* ndef);
size_t fx_jacobi_work_doubles(int m);
// One-sided (Hestenes) Jacobi: dense full-spectrum eigensolver for 1024 < n <= 8192 (smalldense.cu)
int fx_hestenes_eig(fx_ctx* ctx, double* A, int64_t ld, int n, int k, double* Vwork, double* lamwork,
                    int* rankwork, unsigned long long* flag, double* Vout, int64_t ldo, double* theta,
                    int info_host[2]);
int fx_launch_hash_fill(fx_ctx* ctx, double* W, int64_t ld, int64_t rows_valid, int64_t row0,
                        int cols, uint64_t seed);
// SVQB helpers (single CTA)
int fx_launch_svqb_prep(fx_ctx* ctx, const double* G, int ldg, int b, double* Gs, double* dscale,
                        const int* need = nullptr);
int fx_launch_svqb_factors(fx_ctx* ctx, const double* U, int ldu, const double* lam,
                           const double* dscale, int b, double tol, double* M, double* Cfac,
                           int* ndef, const int* need = nullptr);
// C = A(b x b) * B(b x b), all column-major ld b
int fx_launch_small_matmul(fx_ctx* ctx, const double* A, const double* B, int b, double* C);
// T[:mq, m0:m0+b] = H (mq x b, ldh); T[m0:m0+b, :m0] = H[:m0]'; diagonal block symmetrised
int fx_launch_T_update(fx_ctx* ctx, double* T, int ldt, const double* H, int ldh, int m0, int b);
// res[i] = || C (b x b) * S[m-b:m, i] || , i < k
int fx_launch_ritz_residuals(fx_ctx* ctx, const double* C, int b, const double* S, int lds,
                             int m, int k, double* res);
// T = diag(theta[0..r)) (ldt x ldt zeroed)
int fx_launch_T_restart(fx_ctx* ctx, double* T, int ldt, int tsize, const double* theta, int r);
// bias[c] = sum_j colmean[j] * W[c][j]
int fx_launch_bias(fx_ctx* ctx, const double* colmean, const double* W, int64_t ld, int64_t n,
                   int k, double* bias);
// sums[c] = sum_j W[c][j], sums[k + c] = sum_j |W[c][j]|   (j < n)
int fx_launch_colsums(fx_ctx* ctx, const double* W, int64_t ld, int64_t n, int k, double* sums);
// Fills m->wsum / m->need_rowmean from m->W (synchronises the stream).
int fx_model_finish_centering(fx_ctx* ctx, fx_model* m);

// ---- transform.cu ------------------------------------------------------------------------
// out[c][q] = sum_j map(Z_q . X_j) W[c][j] - bias[c]
struct TransformArgs {
    const double* Zp; int64_t ldz; const double* nz;   // queries, feature-major, m_pad % 128 == 0
    const double* Xp; int64_t ldx; const double* nx;   // training set, n_ld % 64 == 0
    int64_t d_pad; int64_t m_pad; int64_t n_ld; int64_t n_valid;
    const double* W; int64_t ldw; int kcols;           // [kcols][ldw], kcols % 8 == 0, <= 64
    const double* bias;                                // [kcols] or null
    double* out; int64_t ldo;                          // [kcols][ldo] (column-major m_pad x kcols)
    int kernel_id; double kp[3];
    // exact double centring for components not orthogonal to 1: out -= (rowmean_q - grand) wsum[c]
    const double* wsum = nullptr;                      // [kcols] 1'W or null
    double grand = 0.0;                                // grand mean of the training kernel matrix
    double* rz_out = nullptr;                          // [m_pad] raw row sums of k(Z, X) or null
};
int fx_launch_transform(fx_ctx* ctx, const TransformArgs& a, int variant);

// ---- solver.cu ---------------------------------------------------------------------------
struct EigArgs {
    const double* K; int64_t ldk;       // local rows of the (centred) kernel matrix
    int64_t n, ld, n_ld, row0, rows_valid;
    int k;
    fx_solver_opts opts;
    double* Vout;                       // device, column-major [k][n_ld] (all rows, gathered)
    double* theta;                      // device, k Ritz values (unscaled), descending
    // Deferred centring (fx_eig_topk only): when non-null, K is NOT yet centred and the first
    // K*V pass centres it in place (symm.cu CENTER variant); K is centred when the call returns.
    const double* center_colmean = nullptr;
    const double* center_grand = nullptr;
    // Locking (k > 64): `nlock` converged eigenvectors (local rows, column-major, leading dimension
    // ld) that the iteration is kept orthogonal to -- the solver then returns the NEXT k pairs of the
    // spectrum.  scale_ref > 0 replaces |theta_1| of this call as the reference of the relative
    // residual tolerance (the largest eigenvalue of the whole problem).
    const double* lock = nullptr;
    int nlock = 0;
    double scale_ref = 0.0;
};
int fx_eig_topk(fx_ctx* ctx, const EigArgs& a, FxTimer& tm, fx_stats* st);
int fx_eig_dense(fx_ctx* ctx, const EigArgs& a, FxTimer& tm, fx_stats* st);
size_t fx_eig_workspace_bytes(fx_ctx* ctx, int64_t ld, int64_t n_ld, int k, const fx_solver_opts& o);
void fx_resolve_opts(const fx_solver_opts* in, int k, fx_solver_opts* out);
int fx_launch_vec_axpy(fx_ctx* ctx, int64_t n, double a, const double* x, double* y);
int fx_launch_copy2d(fx_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int rows,
                     int cols);
// dst[:rows, :cols] += src[:rows, :cols]
int fx_launch_copy2d_add(fx_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int rows,
                         int cols);
// G (b x b, ld b) = Hs[mq:mq+b, :] - Hs[:mq, :]' Hs[:mq, :]   (Hs is (mq + b) x b with leading dimension ldh)
int fx_launch_gram_downdate(fx_ctx* ctx, const double* Hs, int ldh, int mq, int b, double* G);
size_t fx_transform_scratch(fx_ctx* ctx, int64_t m_pad, int64_t n_ld, int kc, int* nsplit_out);
int fx_launch_transform_ex(fx_ctx* ctx, const TransformArgs& a, int variant, double* scratch, int nsplit);
int fx_launch_fill(fx_ctx* ctx, double* p, int64_t n, double v);

// ---- api_kpca.cu: stages shared by the KPCA and KLDA fits ---------------------------------------
// Pads X into feature-major device storage and builds this rank's row block of K (optionally
// centred).  All buffers come from the (already reserved) arena.
struct GramStage {
    int64_t n, d, ld, n_ld, d_pad, row0, rows_valid;
    double *Xp, *norms, *K, *rowpart, *rowsum, *colmean, *scal;   // scal[0]=grand mean, scal[1]=trace
    // Set by the caller before fx_gramstage_run: compute the means but leave K uncentred -- the
    // eigensolver's first K*V pass centres it in place (EigArgs::center_colmean); the caller
    // then finishes with fx_gramstage_trace.
    bool defer_center = false;
};
int fx_gramstage_trace(fx_ctx* ctx, GramStage& g, FxTimer& tm, fx_stats* st);
size_t fx_gramstage_bytes(int64_t ld, int64_t n_ld, int64_t d_pad, bool center);
int fx_gramstage_run(fx_ctx* ctx, GramStage& g, const double* X, bool x_dev, int64_t ldx, int kernel_id,
                     const double* kp, bool center, int variant, FxTimer& tm, fx_stats* st);
int fx_model_build(fx_ctx* ctx, const GramStage& g, int kernel_id, const double* kp, int k, const double* W,
                   bool centered, bool is_klda, fx_model** out);
int fx_launch_indicator(fx_ctx* ctx, const int* ydev, int64_t n, int64_t ldr, int cp, double* E);

// ---- linear.cu: device core of lda! (shared with the dense path of fx_klda_fit_f64) ---------------
size_t fx_lda_core_bytes(fx_ctx* ctx, int64_t n, int64_t p);
int fx_lda_hw_trace(fx_ctx* ctx, double* Xd, int64_t n, int64_t p, int64_t ld, const int64_t* y, int k,
                    const double* inv_cnt, double* out);
int fx_lda_core(fx_ctx* ctx, double* Xd, int64_t n, int64_t p, int64_t ld, const int64_t* y, int k,
                const double* freq, const double* inv_cnt, double alpha_b, double eps_b, double alpha_w,
                double eps_w, double** Vd_out, int64_t* ldv_out, double** lam_out, int hinfo[8]);
