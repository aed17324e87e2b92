/* Plain C99 caller of the drop-in boundary (include/fxb200.h): proves the header is valid C and
 * that a non-Python host can drive a fit + transform through the C ABI.  Built and run by
 * tests/test_library_abi.py (no GPU: fx_init must fail loudly) and tests/test_gpu_parity.py (GPU:
 * in-sample identity transform(X) = V diag(lambda (n-1)), the property of kpca.jl:49-57 + :64-68
 * that SURVEY.md A.2 probes).  Test infrastructure, not product. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/fxb200.h"

int main(void) {
    fx_ctx* ctx = NULL;
    if (fx_version() < 100) { printf("bad version\n"); return 2; }
    int rc = fx_init(0, &ctx);
    if (rc != 0) {
        /* no usable sm_100 GPU: the library must say so, not fall back */
        printf("NO_GPU rc=%d msg=%s\n", rc, fx_last_error());
        return 3;
    }
    const int64_t n = 300, d = 5;
    const int k = 3;
    double* X = (double*)malloc(sizeof(double) * n * d);       /* column-major n x d */
    unsigned long long st = 42;
    for (int64_t i = 0; i < n * d; ++i) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        X[i] = (double)(st >> 11) / 9007199254740992.0 + ((i % n) % 3);   /* three loose groups */
    }
    double kp[3] = {0.2, 0.0, 0.0};
    double* V = (double*)malloc(sizeof(double) * n * k);
    double lam[3];
    fx_model* model = NULL;
    fx_stats stats;
    memset(&stats, 0, sizeof(stats));
    rc = fx_kpca_fit_f64(ctx, X, n, d, n, FX_KERNEL_GAUSSIAN, kp, k, 1, 0.0, 0.0, NULL, V, n, lam, &model, &stats);
    if (rc != 0) { printf("fit failed rc=%d %s\n", rc, fx_last_error()); return 4; }
    double worst_orth = 0.0;
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            double s = 0.0;
            for (int64_t i = 0; i < n; ++i) s += V[a * n + i] * V[b * n + i];
            const double e = fabs(s - (a == b ? 1.0 : 0.0));
            if (e > worst_orth) worst_orth = e;
        }
    double* T = (double*)malloc(sizeof(double) * n * k);
    rc = fx_kpca_transform_f64(ctx, model, X, n, n, T, n, NULL);
    if (rc != 0) { printf("transform failed rc=%d %s\n", rc, fx_last_error()); return 5; }
    double worst_id = 0.0, scale = lam[0] * (double)(n - 1);
    for (int c = 0; c < k; ++c)
        for (int64_t i = 0; i < n; ++i) {
            const double e = fabs(T[c * n + i] - V[c * n + i] * lam[c] * (double)(n - 1));
            if (e > worst_id) worst_id = e;
        }
    /* argument errors come back as -i with a message, never abort */
    const int bad = fx_kpca_fit_f64(ctx, X, n, d, n, 7, kp, k, 1, 0.0, 0.0, NULL, V, n, lam, NULL, NULL);
    printf("C_ABI_OK lambda=%.6e %.6e %.6e orth=%.2e insample=%.2e launches=%lld bad_kernel_rc=%d\n", lam[0], lam[1],
           lam[2], worst_orth, worst_id / scale, (long long)stats.kernel_launches, bad);
    const int ok = lam[0] >= lam[1] && lam[1] >= lam[2] && lam[2] > 0 && worst_orth < 1e-12 &&
                   worst_id / scale < 1e-12 && bad < 0 && stats.kernel_launches > 0;
    fx_model_free(model);
    fx_shutdown(ctx);
    free(X); free(V); free(T);
    return ok ? 0 : 6;
}
