// api_klda.cu -- C ABI: Kernel LDA fit (placeholder until klda.cu lands in this round).
#include "fx_kernels.cuh"

extern "C" int fx_klda_fit_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                               int kernel_id, const double* kparams, const int64_t* y, int nclass,
                               const double* freq, double alpha_b, double eps_b, double alpha_w,
                               double eps_w, int ncomp, const fx_solver_opts* opts, double* A,
                               int64_t lda, double* lambda, fx_model** model, fx_stats* stats) {
    (void)ctx; (void)X; (void)n; (void)d; (void)ldx; (void)kernel_id; (void)kparams; (void)y; (void)nclass;
    (void)freq; (void)alpha_b; (void)eps_b; (void)alpha_w; (void)eps_w; (void)ncomp; (void)opts; (void)A;
    (void)lda; (void)lambda; (void)model; (void)stats;
    fx_set_error("fx_klda_fit_f64: not implemented in this build");
    return FX_ERR_CUDA;
}
