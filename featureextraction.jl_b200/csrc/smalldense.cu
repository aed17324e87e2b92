// smalldense.cu -- the small dense pieces of the eigensolver, all on the device:
//   * parallel cyclic Jacobi eigensolver (Rayleigh-Ritz on the projected matrix T and the
//     b x b Gram matrices of the SVQB orthonormalisation); also the dense path for tiny n.
//     This is what stands in for LAPACK dsyevd (/root/reference/src/common.jl:128-131) on the
//     projected problem; descending order like components_eig! (common.jl:164).
//   * SVQB factor construction, T bookkeeping, Ritz residuals, deterministic start block.
// Single-CTA kernels: the matrices are at most a few hundred rows.
#include "fx_kernels.cuh"

namespace {

constexpr int J_SMEM_BOTH_MAX_M = 112;           // A and V both in shared memory up to this order
constexpr int J_SMEM_MAX_M = 160;                // A alone in shared memory up to this order

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// Round-robin tournament: players 0..mp-1 (mp even), player mp-1 fixed.
__device__ __forceinline__ void rr_pair(int rnd, int i, int mp, int& p, int& q) {
    if (i == 0) {
        p = mp - 1;
        q = rnd;
    } else {
        p = rnd + i;
        if (p >= mp - 1) p -= mp - 1;
        q = rnd - i;
        if (q < 0) q += mp - 1;
    }
    if (p > q) { const int x = p; p = q; q = x; }
}

// Symmetric Schur rotation annihilating a_pq:  t = h sgn(d) / (|d| + sqrt(d^2 + h^2)) with
// d = a_qq - a_pp, h = 2 a_pq; c = 1/sqrt(1+t^2), s = t c.  rsqrt/rcp based (no IEEE div/sqrt).
__device__ __forceinline__ void schur_rotation(double app, double aqq, double apq, double& c, double& s) {
    c = 1.0; s = 0.0;
    if (fabs(apq) > 1e-300) {
        const double d = aqq - app, h = 2.0 * apq;
        const double r2 = fma(d, d, h * h);
        const double sq = r2 * rsqrt(r2);
        const double den = fabs(d) + sq;
        const double t = (d >= 0.0 ? h : -h) * __drcp_rn(den);
        c = rsqrt(fma(t, t, 1.0));
        s = t * c;
    }
}

// A (m x m, lda) symmetric, destroyed.  Vg (m x m, ldv) is global scratch for the accumulated
// rotations when they do not fit in shared memory; w / Vout receive the sorted (descending)
// eigenvalues / eigenvectors.  One warp owns a rotation pair for the whole round: it derives the
// rotation itself, rotates its two columns of A and V, and after one barrier its two rows of A.
template <bool SMEM_A, bool SMEM_V>
__global__ void __launch_bounds__(1024, 1) jacobi_kernel(double* Ag, int lda_g, int m, double* Vg, int ldvg,
                                                         double* w, double* Vout, int ldvo, int* info,
                                                         int max_sweeps, const int* need) {
    extern __shared__ __align__(16) double sm[];
    if (need && *need == 0) return;               // the Cholesky fast path already produced the factors
    __shared__ double red[33];
    __shared__ int s_rank[1024];
    __shared__ double rot_c[512], rot_s[512];     // written and read by the pair's owner warp only
    __shared__ int rot_p[512], rot_q[512];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int mp = m + (m & 1);
    const int half = mp / 2;

    const int lda = SMEM_A ? (m | 1) : lda_g;     // odd stride in smem: conflict-free row access
    const int ldv = SMEM_V ? (m | 1) : ldvg;
    double* A = SMEM_A ? sm : Ag;
    double* V = SMEM_V ? (sm + (size_t)m * (m | 1)) : Vg;
    if (SMEM_A)
        for (int j = warp; j < m; j += nwarps)
            for (int i = lane; i < m; i += 32) A[i + j * lda] = Ag[i + (long)j * lda_g];
    for (int j = warp; j < m; j += nwarps)
        for (int i = lane; i < m; i += 32) V[i + (long)j * ldv] = (i == j) ? 1.0 : 0.0;
    __syncthreads();

    int sweeps = 0, converged = 0;
    double prev_off = 0.0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int j = warp; j < m; j += nwarps)
            for (int i = lane; i < m; i += 32) {
                const double a = A[i + (long)j * lda];
                if (i == j) dg += a * a; else off += a * a;
            }
        off = block_sum(off, red);
        dg = block_sum(dg, red);
        if (off <= 1e-28 * dg || off == 0.0) { converged = 1; break; }   // off_F <= 1e-14 |A|_F
        // rounding floor reached: the off-diagonal norm stopped contracting
        if (sweep > 2 && off > 0.25 * prev_off && off <= 1e-24 * dg) { converged = 1; break; }
        prev_off = off;
        sweeps++;
        for (int rnd = 0; rnd < mp - 1; ++rnd) {
            // phase A: rotation + column rotations  A <- A J,  V <- V J   (lanes over rows)
            for (int pr = warp; pr < half; pr += nwarps) {
                int pp, qq;
                rr_pair(rnd, pr, mp, pp, qq);
                double c = 1.0, s = 0.0;
                if (qq < m) {
                    schur_rotation(A[pp + (long)pp * lda], A[qq + (long)qq * lda], A[pp + (long)qq * lda], c, s);
                } else {
                    qq = -1;
                }
                __syncwarp();                     // all lanes have read the 2x2 block before it changes
                if (lane == 0) { rot_c[pr] = c; rot_s[pr] = s; rot_p[pr] = pp; rot_q[pr] = qq; }
                if (qq < 0 || s == 0.0) continue;
                double* Ap = A + (long)pp * lda; double* Aq = A + (long)qq * lda;
                double* Vp = V + (long)pp * ldv; double* Vq = V + (long)qq * ldv;
                for (int r = lane; r < m; r += 32) {
                    const double ap = Ap[r], aq = Aq[r], vp = Vp[r], vq = Vq[r];
                    Ap[r] = c * ap - s * aq;
                    Aq[r] = s * ap + c * aq;
                    Vp[r] = c * vp - s * vq;
                    Vq[r] = s * vp + c * vq;
                }
            }
            __syncthreads();
            // phase B: row rotations  A <- J' A   (same warp, same pairs; lanes over columns)
            for (int pr = warp; pr < half; pr += nwarps) {
                const int pp = rot_p[pr], qq = rot_q[pr];
                const double c = rot_c[pr], s = rot_s[pr];
                if (qq < 0 || s == 0.0) continue;
                for (int cc = lane; cc < m; cc += 32) {
                    const double ap = A[pp + (long)cc * lda], aq = A[qq + (long)cc * lda];
                    double np_ = c * ap - s * aq, nq_ = s * ap + c * aq;
                    if (cc == qq) np_ = 0.0;      // the annihilated pair, exactly
                    if (cc == pp) nq_ = 0.0;
                    A[pp + (long)cc * lda] = np_;
                    A[qq + (long)cc * lda] = nq_;
                }
            }
            __syncthreads();
        }
    }
    // sort descending by rank counting (stable), scatter eigenvectors
    for (int i = tid; i < m; i += nt) {
        const double di = A[i + (long)i * lda];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            const double dj = A[j + (long)j * lda];
            rank += (dj > di) || (dj == di && j < i);
        }
        w[rank] = di;
        s_rank[i] = rank;
    }
    __syncthreads();
    for (int j = warp; j < m; j += nwarps) {
        const int dst = s_rank[j];
        for (int r = lane; r < m; r += 32) Vout[r + (long)dst * ldvo] = V[r + (long)j * ldv];
    }
    if (tid == 0 && info) { info[0] = sweeps; info[1] = converged; }
}

__device__ __forceinline__ uint64_t splitmix(uint64_t row, uint64_t col, uint64_t seed) {
    uint64_t z = (row * 0x9E3779B97F4A7C15ULL) ^ (col * 0xBF58476D1CE4E5B9ULL) ^ seed;
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void hash_fill_kernel(double* W, long ld, long rows_valid, long row0, int cols, uint64_t seed) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= ld) return;
    double v = 0.0;
    if (i < rows_valid)
        v = (double)(splitmix((uint64_t)(row0 + i), (uint64_t)c, seed) >> 11) * 0x1.0p-52 - 1.0;
    W[(long)c * ld + i] = v;
}

// Fast path of the block orthonormalisation: Cholesky of the diagonally scaled Gram matrix
// (CholQR).  G = D Gs D, Gs = L L'  =>  W_orth = W D^-1 L^-T,  W = W_orth (L' D).
// Writes M = D^-1 L^-T and Cfac = L' D and need_svqb = 0; if a pivot falls below `thr` (block
// numerically rank deficient / ill conditioned) writes need_svqb = 1 and the SVQB (Jacobi)
// kernels that follow do the work instead -- they return at once when need_svqb == 0.
__global__ void __launch_bounds__(256) chol_factors_kernel(const double* G, int ldg, int b, double thr,
                                                           double* M, double* Cfac, int* need_svqb,
                                                           int* ndef) {
    extern __shared__ __align__(16) double chol_sm[];
    double (*L)[65] = reinterpret_cast<double (*)[65]>(chol_sm);
    double (*Li)[65] = reinterpret_cast<double (*)[65]>(chol_sm + 64 * 65);
    double* d = chol_sm + 2 * 64 * 65;
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    if (tid < b) d[tid] = sqrt(fmax(G[tid + (long)tid * ldg], 1e-300));
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        L[i][j] = 0.5 * (G[i + (long)j * ldg] + G[j + (long)i * ldg]) / (d[i] * d[j]);
        Li[i][j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < b; ++j) {
        const double piv = L[j][j];
        if (!(piv > thr)) { if (tid == 0) bad = 1; break; }       // uniform: every thread reads the same value
        const double inv = rsqrt(piv);
        __syncthreads();
        for (int i = j + tid; i < b; i += blockDim.x) L[i][j] *= inv;        // column j (incl. diagonal -> sqrt)
        __syncthreads();
        const int rem = b - j - 1;
        for (int idx = tid; idx < rem * rem; idx += blockDim.x) {
            const int i = j + 1 + idx % rem, k = j + 1 + idx / rem;
            if (k <= i) L[i][k] -= L[i][j] * L[k][j];
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad) {
        if (tid == 0) *need_svqb = 1;
        return;
    }
    // Li = L^-1 by forward substitution, one thread per column of the identity
    if (tid < b) {
        const int c = tid;
        for (int i = c; i < b; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= L[i][k] * Li[k][c];
            Li[i][c] = s / L[i][i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        // M = D^-1 L^-T : M[i][j] = Li[j][i] / d[i] (upper triangular);  Cfac = L' D : Cfac[i][j] = L[j][i] d[j]
        M[i + j * b] = (j >= i) ? Li[j][i] / d[i] : 0.0;
        Cfac[i + j * b] = (j >= i) ? L[j][i] * d[j] : 0.0;
    }
    if (tid == 0) { *need_svqb = 0; if (ndef) *ndef = 0; }
}

__global__ void svqb_prep_kernel(const double* G, int ldg, int b, double* Gs, double* dscale, const int* need) {
    __shared__ double d[64];
    if (need && *need == 0) return;
    const int tid = threadIdx.x;
    if (tid < b) {
        const double gii = G[tid + (long)tid * ldg];
        d[tid] = sqrt(fmax(gii, 1e-300));
        dscale[tid] = d[tid];
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        // symmetrise: the two triangles come from different reduction orders
        const double gij = 0.5 * (G[i + (long)j * ldg] + G[j + (long)i * ldg]);
        Gs[i + j * b] = gij / (d[i] * d[j]);
    }
}

// M = D^-1 U L^-1/2 (W_orth = W M),  Cfac = L^1/2 U' D (W = W_orth Cfac)
__global__ void svqb_factors_kernel(const double* U, int ldu, const double* lam, const double* dscale,
                                    int b, double tol, double* M, double* Cfac, int* ndef, const int* need) {
    __shared__ double l[64];
    __shared__ int cnt;
    if (need && *need == 0) return;
    const int tid = threadIdx.x;
    if (tid == 0) cnt = 0;
    __syncthreads();
    if (tid < b) {
        const double lmax = fmax(lam[0], 1e-300);
        double v = lam[tid];
        if (v < tol * lmax) { v = tol * lmax; atomicAdd(&cnt, 1); }
        l[tid] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        M[i + j * b] = U[i + (long)j * ldu] / (sqrt(l[j]) * dscale[i]);
        Cfac[i + j * b] = sqrt(l[i]) * U[j + (long)i * ldu] * dscale[j];
    }
    if (tid == 0 && ndef) *ndef = cnt;
}

__global__ void small_matmul_kernel(const double* A, const double* B, int b, double* C) {
    for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        double s = 0.0;
        for (int k = 0; k < b; ++k) s += A[i + k * b] * B[k + j * b];
        C[idx] = s;
    }
}

__global__ void T_update_kernel(double* T, int ldt, const double* H, int ldh, int m0, int b) {
    const int mq = m0 + b;
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < mq * b; idx += blockDim.x * gridDim.x) {
        const int i = idx % mq, j = idx / mq;
        double v = H[i + (long)j * ldh];
        if (i >= m0) v = 0.5 * (v + H[(m0 + j) + (long)(i - m0) * ldh]);   // symmetrise the diagonal block
        T[i + (long)(m0 + j) * ldt] = v;
        if (i < m0) T[(m0 + j) + (long)i * ldt] = v;
    }
}

__global__ void ritz_residual_kernel(const double* C, int b, const double* S, int lds, int m, int k,
                                     double* res) {
    const int i = blockIdx.x;                    // Ritz pair
    __shared__ double red[33];
    double acc = 0.0;
    for (int r = threadIdx.x; r < b; r += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < b; ++c) s += C[r + c * b] * S[(m - b + c) + (long)i * lds];
        acc += s * s;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && i < k) res[i] = sqrt(acc);
}

__global__ void T_restart_kernel(double* T, int ldt, int tsize, const double* theta, int r) {
    for (long idx = threadIdx.x + (long)blockIdx.x * blockDim.x; idx < (long)tsize * tsize;
         idx += (long)blockDim.x * gridDim.x) {
        const int i = (int)(idx % tsize), j = (int)(idx / tsize);
        T[i + (long)j * ldt] = (i == j && i < r) ? theta[i] : 0.0;
    }
}

__global__ void bias_kernel(const double* colmean, const double* W, long ld, long n, int k, double* bias) {
    __shared__ double red[33];
    const int c = blockIdx.x;
    double s = 0.0;
    for (long j = threadIdx.x; j < n; j += blockDim.x) s += colmean[j] * W[(long)c * ld + j];
    s = block_sum(s, red);
    if (threadIdx.x == 0) bias[c] = s;
}

__global__ void colsums_kernel(const double* W, long ld, long n, int k, double* sums) {
    __shared__ double red[33];
    const int c = blockIdx.x;
    double s = 0.0, a = 0.0;
    for (long j = threadIdx.x; j < n; j += blockDim.x) { const double w = W[(long)c * ld + j]; s += w; a += fabs(w); }
    s = block_sum(s, red);
    a = block_sum(a, red);
    if (threadIdx.x == 0) { sums[c] = s; sums[k + c] = a; }
}

// ---- one-sided (Hestenes) Jacobi for the dense full-spectrum path at 1024 < n <= 8192 --------------------
// G = A V is kept with mutually orthogonal columns in the limit: for symmetric A the column norms are
// |lambda_j| and the columns of V the eigenvectors.  One CTA per column pair of a round-robin round
// (n/2 disjoint pairs: perfectly parallel), three dot products, one plane rotation of the two columns
// of G and of V.  Memory-bound O(n^3): ~10 sweeps x 32 n^3 bytes.
__global__ void __launch_bounds__(256) hestenes_round_kernel(double* __restrict__ G, double* __restrict__ V, long ld,
                                                             int n, int np, int rnd, double floor2,
                                                             unsigned long long* __restrict__ offmax) {
    __shared__ double red[33];
    int p, q;
    rr_pair(rnd, blockIdx.x, np, p, q);
    if (q >= n) return;                                          // the bye of an odd n
    double* gp = G + (long)p * ld; double* gq = G + (long)q * ld;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double x = gp[r], y = gq[r];
        a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
    }
    a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
    const double ab = sqrt(a) * sqrt(b);
    // converged pair: relatively orthogonal, or both columns at the noise floor of the matrix
    if (fabs(c) <= 1e-15 * ab || fabs(c) <= floor2) return;
    if (threadIdx.x == 0) atomicMax(offmax, (unsigned long long)__double_as_longlong(fabs(c) / ab));
    const double zeta = (b - a) / (2.0 * c);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
    const double cs = rsqrt(fma(t, t, 1.0)), sn = cs * t;
    double* vp = V + (long)p * ld; double* vq = V + (long)q * ld;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double x = gp[r], y = gq[r];
        gp[r] = cs * x - sn * y; gq[r] = sn * x + cs * y;
        const double u = vp[r], w = vq[r];
        vp[r] = cs * u - sn * w; vq[r] = sn * u + cs * w;
    }
}
// lam[j] = sign(v_j . g_j) |g_j|  (Rayleigh quotient of the unit vector v_j), one CTA per column
__global__ void __launch_bounds__(256) hestenes_values_kernel(const double* __restrict__ G, const double* __restrict__ V,
                                                              long ld, int n, double* __restrict__ lam) {
    __shared__ double red[33];
    const int j = blockIdx.x;
    double a = 0.0, c = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double x = G[(long)j * ld + r];
        a = fma(x, x, a); c = fma(x, V[(long)j * ld + r], c);
    }
    a = block_sum(a, red); c = block_sum(c, red);
    if (threadIdx.x == 0) lam[j] = (c < 0.0 ? -1.0 : 1.0) * sqrt(a);
}
// descending rank of every eigenvalue (stable), one thread per value
__global__ void rank_desc_kernel(const double* __restrict__ lam, int n, int* __restrict__ rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double di = lam[i];
    int r = 0;
    for (int j = 0; j < n; ++j) { const double dj = lam[j]; r += (dj > di) || (dj == di && j < i); }
    rank[i] = r;
}
// Vout[:, rank[j]] = V[:, j] for rank[j] < k; theta[rank[j]] = lam[j]
__global__ void scatter_sorted_kernel(const double* __restrict__ V, long ldv, const double* __restrict__ lam,
                                      const int* __restrict__ rank, int n, int k, double* __restrict__ Vout, long ldo,
                                      double* __restrict__ theta) {
    const int j = blockIdx.x, d = rank[j];
    if (d >= k) return;
    for (int r = threadIdx.x; r < n; r += blockDim.x) Vout[(long)d * ldo + r] = V[(long)j * ldv + r];
    if (threadIdx.x == 0) theta[d] = lam[j];
}
__global__ void identity_kernel(double* V, long ld, int n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < ld) V[(long)j * ld + i] = (i == j && i < n) ? 1.0 : 0.0;
}
__global__ void __launch_bounds__(1024) max_colnorm2_kernel(const double* __restrict__ G, long ld, int n, double* out) {
    __shared__ double red[33];
    double m = 0.0;
    for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int r = threadIdx.x; r < n; r += blockDim.x) { const double x = G[(long)j * ld + r]; a = fma(x, x, a); }
        a = block_sum(a, red);
        m = fmax(m, a);
    }
    if (threadIdx.x == 0) *out = m;
}

}  // namespace

// Dense symmetric eigendecomposition by one-sided Jacobi, all on the device: A (n x n symmetric,
// column-major with leading dimension ld >= n; DESTROYED) -> the k leading (largest algebraic)
// eigenvalues in theta and eigenvectors in Vout (leading dimension ldo).  Vwork: ld x n scratch,
// lamwork: n, rankwork: n ints, flag: 8 bytes.  info_host[0] = sweeps, [1] = converged.
int fx_hestenes_eig(fx_ctx* ctx, double* A, int64_t ld, int n, int k, double* Vwork, double* lamwork,
                    int* rankwork, unsigned long long* flag, double* Vout, int64_t ldo, double* theta,
                    int info_host[2]) {
    cudaStream_t s = ctx->stream;
    const int np = n + (n & 1);
    dim3 gi((unsigned)ceil_div(ld, 256), (unsigned)n);
    identity_kernel<<<gi, 256, 0, s>>>(Vwork, ld, n);
    max_colnorm2_kernel<<<1, 1024, 0, s>>>(A, ld, n, lamwork);
    ctx->launches += 2;
    FX_LAUNCH_CHECK();
    double scale2 = 0.0;
    FX_CUDA(cudaMemcpyAsync(&scale2, lamwork, 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    const double floor2 = 1e-26 * scale2;                        // (1e-13 |lambda_1|)^2
    int sweeps = 0, conv = 0;
    for (; sweeps < 40 && !conv; ++sweeps) {
        FX_CUDA(cudaMemsetAsync(flag, 0, 8, s));
        for (int rnd = 0; rnd < np - 1; ++rnd) {
            hestenes_round_kernel<<<np / 2, 256, 0, s>>>(A, Vwork, ld, n, np, rnd, floor2, flag);
            ctx->launches++;
        }
        FX_LAUNCH_CHECK();
        unsigned long long bits = 0;
        FX_CUDA(cudaMemcpyAsync(&bits, flag, 8, cudaMemcpyDeviceToHost, s));
        FX_CUDA(cudaStreamSynchronize(s));
        double off;
        memcpy(&off, &bits, 8);
        conv = off <= 1e-14;                                     // a sweep whose largest rotation was at rounding level
    }
    hestenes_values_kernel<<<n, 256, 0, s>>>(A, Vwork, ld, n, lamwork);
    rank_desc_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(lamwork, n, rankwork);
    scatter_sorted_kernel<<<n, 256, 0, s>>>(Vwork, ld, lamwork, rankwork, n, k, Vout, ldo, theta);
    ctx->launches += 3;
    FX_LAUNCH_CHECK();
    info_host[0] = sweeps; info_host[1] = conv;
    return 0;
}

size_t fx_jacobi_work_doubles(int m) { return (size_t)m * m; }

int fx_launch_jacobi(fx_ctx* ctx, double* A, int lda, int m, double* Vout, int ldv, double* w,
                     double* Vwork, int* info, const int* need) {
    if (m < 1 || m > 1024) {
        fx_set_error("jacobi: order %d out of range", m);
        return FX_ERR_CUDA;
    }
    const int mode = (m <= J_SMEM_BOTH_MAX_M) ? 2 : ((m <= J_SMEM_MAX_M) ? 1 : 0);
    const size_t smem = (size_t)mode * m * (m | 1) * 8;
    // balanced: every warp owns the same number of rotation pairs (at most 16)
    const int half = (m + 1) / 2;
    int per = (half + 31) / 32;
    int nwarps = (half + per - 1) / per;
    if (nwarps < 2) nwarps = 2;
    const int threads = 32 * nwarps;
    // largest dynamic request of each mode (the kernel also holds ~17 KB of static shared memory)
    FX_SET_SMEM_ONCE(ctx, (jacobi_kernel<true, true>), 2 * J_SMEM_BOTH_MAX_M * (J_SMEM_BOTH_MAX_M | 1) * 8);
    FX_SET_SMEM_ONCE(ctx, (jacobi_kernel<true, false>), J_SMEM_MAX_M * (J_SMEM_MAX_M | 1) * 8);
    if (mode == 2) jacobi_kernel<true, true><<<1, threads, smem, ctx->stream>>>(A, lda, m, Vwork, m, w, Vout, ldv, info, 40, need);
    else if (mode == 1) jacobi_kernel<true, false><<<1, threads, smem, ctx->stream>>>(A, lda, m, Vwork, m, w, Vout, ldv, info, 40, need);
    else jacobi_kernel<false, false><<<1, threads, 0, ctx->stream>>>(A, lda, m, Vwork, m, w, Vout, ldv, info, 40, need);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_hash_fill(fx_ctx* ctx, double* W, int64_t ld, int64_t rows_valid, int64_t row0,
                        int cols, uint64_t seed) {
    if (ld <= 0) return 0;
    dim3 grid((unsigned)ceil_div(ld, 256), (unsigned)cols);
    hash_fill_kernel<<<grid, 256, 0, ctx->stream>>>(W, ld, rows_valid, row0, cols, seed);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_chol_factors(fx_ctx* ctx, const double* G, int ldg, int b, double thr, double* M, double* Cfac,
                           int* need_svqb, int* ndef) {
    if (b < 1 || b > 64) {
        fx_set_error("chol_factors: block width %d out of range", b);
        return FX_ERR_CUDA;
    }
    const size_t smem = (2 * 64 * 65 + 64) * sizeof(double);
    FX_SET_SMEM_ONCE(ctx, chol_factors_kernel, smem);
    chol_factors_kernel<<<1, 256, smem, ctx->stream>>>(G, ldg, b, thr, M, Cfac, need_svqb, ndef);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_svqb_prep(fx_ctx* ctx, const double* G, int ldg, int b, double* Gs, double* dscale, const int* need) {
    svqb_prep_kernel<<<1, 256, 0, ctx->stream>>>(G, ldg, b, Gs, dscale, need);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_svqb_factors(fx_ctx* ctx, const double* U, int ldu, const double* lam,
                           const double* dscale, int b, double tol, double* M, double* Cfac,
                           int* ndef, const int* need) {
    svqb_factors_kernel<<<1, 256, 0, ctx->stream>>>(U, ldu, lam, dscale, b, tol, M, Cfac, ndef, need);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_small_matmul(fx_ctx* ctx, const double* A, const double* B, int b, double* C) {
    small_matmul_kernel<<<1, 256, 0, ctx->stream>>>(A, B, b, C);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_T_update(fx_ctx* ctx, double* T, int ldt, const double* H, int ldh, int m0, int b) {
    T_update_kernel<<<8, 256, 0, ctx->stream>>>(T, ldt, H, ldh, m0, b);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_ritz_residuals(fx_ctx* ctx, const double* C, int b, const double* S, int lds,
                             int m, int k, double* res) {
    ritz_residual_kernel<<<k, 64, 0, ctx->stream>>>(C, b, S, lds, m, k, res);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_T_restart(fx_ctx* ctx, double* T, int ldt, int tsize, const double* theta, int r) {
    T_restart_kernel<<<16, 256, 0, ctx->stream>>>(T, ldt, tsize, theta, r);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_bias(fx_ctx* ctx, const double* colmean, const double* W, int64_t ld, int64_t n,
                   int k, double* bias) {
    if (k <= 0) return 0;
    bias_kernel<<<k, 256, 0, ctx->stream>>>(colmean, W, ld, n, k, bias);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_colsums(fx_ctx* ctx, const double* W, int64_t ld, int64_t n, int k, double* sums) {
    if (k <= 0) return 0;
    colsums_kernel<<<k, 256, 0, ctx->stream>>>(W, ld, n, k, sums);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

namespace {
__global__ void vec_axpy_kernel(long n, double a, const double* x, double* y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
__global__ void copy2d_add_kernel(const double* src, long lds, double* dst, long ldd, int rows, int cols) {
    for (long idx = threadIdx.x + (long)blockIdx.x * blockDim.x; idx < (long)rows * cols;
         idx += (long)blockDim.x * gridDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        dst[i + j * ldd] += src[i + j * lds];
    }
}
// one CTA: G[i][j] = Hs[mq + i][j] - sum_r Hs[r][i] Hs[r][j]
__global__ void gram_downdate_kernel(const double* Hs, int ldh, int mq, int b, double* G) {
    for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
        const int i = idx % b, j = idx / b;
        double s = 0.0;
        for (int r = 0; r < mq; ++r) s = fma(Hs[r + (long)i * ldh], Hs[r + (long)j * ldh], s);
        G[i + j * b] = Hs[mq + i + (long)j * ldh] - s;
    }
}
__global__ void copy2d_kernel(const double* src, long lds, double* dst, long ldd, int rows, int cols) {
    for (long idx = threadIdx.x + (long)blockIdx.x * blockDim.x; idx < (long)rows * cols;
         idx += (long)blockDim.x * gridDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        dst[i + j * ldd] = src[i + j * lds];
    }
}
}  // namespace

int fx_launch_vec_axpy(fx_ctx* ctx, int64_t n, double a, const double* x, double* y) {
    if (n <= 0) return 0;
    vec_axpy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(n, a, x, y);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_copy2d_add(fx_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int rows,
                         int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    copy2d_add_kernel<<<32, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_gram_downdate(fx_ctx* ctx, const double* Hs, int ldh, int mq, int b, double* G) {
    gram_downdate_kernel<<<1, 1024, 0, ctx->stream>>>(Hs, ldh, mq, b, G);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_copy2d(fx_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int rows,
                     int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    copy2d_kernel<<<32, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

namespace {
__global__ void fill_kernel(double* p, long n, double v) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // namespace
int fx_launch_fill(fx_ctx* ctx, double* p, int64_t n, double v) {
    if (n <= 0) return 0;
    fill_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(p, n, v);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}
