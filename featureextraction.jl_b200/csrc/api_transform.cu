// api_transform.cu -- C ABI: out-of-sample projection (transform_kpca, kpca.jl:64-68;
// transform_klda, klda.jl:71) through the fused kernel of transform.cu, and the stateless
// model constructor that serves the reference's signature transform_kpca(X, k, W, Z).
#include "fx_kernels.cuh"

namespace {
constexpr int64_t QCHUNK = 131072;      // queries per launch (bounds device scratch)

int transform_impl(fx_ctx* ctx, const fx_model* md, const double* Z, bool on_dev, int64_t m,
                   int64_t ldz, double* out, int64_t ldo, fx_stats* stats, int variant) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("transform: ctx is NULL"); return -1; }
    if (!md) { fx_set_error("transform: model is NULL"); return -2; }
    if (!ctx->sub.empty()) {
        // multi-GPU context: the queries are sharded over the devices (kpca.jl:64-68 has no coupling
        // between queries), every device holds a replica of the model, no collective
        if (md->sub.size() != ctx->sub.size()) { fx_set_error("transform: the model was not fitted on this multi-GPU context"); return -2; }
        if (on_dev) { fx_set_error("transform: device-pointer entry points need a single-device context"); return -1; }
        if (!Z) { fx_set_error("transform_impl: invalid argument 3 (!Z)"); return -3; }
        if (m < 1) { fx_set_error("transform_impl: invalid argument 4 (m < 1)"); return -4; }
        if (!out) { fx_set_error("transform_impl: invalid argument 6 (!out)"); return -6; }
        const int G = (int)ctx->sub.size();
        std::vector<fx_stats> sts(G);
        for (auto& x : sts) memset(&x, 0, sizeof(x));
        const int rc = fx_multi_run(ctx, [&](fx_ctx* c, int r) {
            const int64_t base = m / G, rem = m % G;
            const int64_t q0 = r * base + (r < rem ? r : rem), cnt = base + (r < rem ? 1 : 0);
            if (cnt == 0) return 0;
            return transform_impl(c, md->sub[r], Z + q0, false, cnt, ldz, out + q0, ldo, &sts[r], variant);
        });
        if (stats) {
            *stats = sts[0];
            for (int r = 1; r < G; ++r) {
                if (sts[r].total_ms > stats->total_ms) stats->total_ms = sts[r].total_ms;
                if (sts[r].transform_ms > stats->transform_ms) stats->transform_ms = sts[r].transform_ms;
                stats->h2d_bytes += sts[r].h2d_bytes; stats->d2h_bytes += sts[r].d2h_bytes;
                stats->transform_launches += sts[r].transform_launches; stats->kernel_launches += sts[r].kernel_launches;
            }
        }
        return rc;
    }
    if (!Z) { fx_set_error("transform_impl: invalid argument 3 (!Z)"); return -3; }
    if (m < 1) { fx_set_error("transform_impl: invalid argument 4 (m < 1)"); return -4; }
    if (ldz < m) { fx_set_error("transform: ldz < m (DimensionMismatch)"); return -5; }
    if (!out) { fx_set_error("transform_impl: invalid argument 6 (!out)"); return -6; }
    if (ldo < m) { fx_set_error("transform: ldo < m (DimensionMismatch)"); return -7; }
    if (md->device != ctx->device) { fx_set_error("transform: model lives on another device"); return -2; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int kc = md->k_pad;
    const int kchunk = kc < 64 ? kc : 64;            // components per launch (the kernel holds <= 64 columns)
    const int64_t mc_max = m < QCHUNK ? round_up(m, 128) : QCHUNK;
    int nsplit = 1;
    const size_t scr = fx_transform_scratch(ctx, mc_max, md->n_ld, kchunk, &nsplit);
    size_t need = fx_align256((size_t)md->d_pad * mc_max * 8) + fx_align256((size_t)mc_max * 8) +
                  fx_align256((size_t)kc * mc_max * 8) + fx_align256(scr * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Zp = (double*)fx_arena_alloc(ctx, (size_t)md->d_pad * mc_max * 8);
    double* nz = (double*)fx_arena_alloc(ctx, (size_t)mc_max * 8);
    double* od = (double*)fx_arena_alloc(ctx, (size_t)kc * mc_max * 8);
    double* scratch = scr ? (double*)fx_arena_alloc(ctx, scr * 8) : nullptr;
    const bool rowmean = md->centered && md->need_rowmean;

    cudaEvent_t e0, e1;
    FX_CUDA(cudaEventCreate(&e0));
    FX_CUDA(cudaEventCreate(&e1));
    FX_CUDA(cudaEventRecord(e0, s));
    FxTimer tm(ctx);
    for (int64_t q0 = 0; q0 < m; q0 += QCHUNK) {
        const int64_t mc = (m - q0) < QCHUNK ? (m - q0) : QCHUNK;
        const int64_t mp = round_up(mc, 128);
        int ns = 1;
        fx_transform_scratch(ctx, mp, md->n_ld, kchunk, &ns);
        if (ns > nsplit) ns = nsplit;
        tm.begin(&st->h2d_ms);
        FX_CUDA(cudaMemsetAsync(Zp, 0, (size_t)md->d_pad * mp * 8, s));
        FX_CUDA(cudaMemcpy2DAsync(Zp, mp * 8, Z + q0, ldz * 8, mc * 8, md->d,
                                  on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        tm.end();
        if (!on_dev) st->h2d_bytes += mc * md->d * 8;
        FX_TRY(fx_launch_rownorms(ctx, Zp, mp, mp, md->d_pad, nz));
        for (int c0 = 0; c0 < kc; c0 += 64) {
            TransformArgs ta;
            ta.Zp = Zp; ta.ldz = mp; ta.nz = nz; ta.Xp = md->Xp; ta.ldx = md->n_ld; ta.nx = md->norms;
            ta.d_pad = md->d_pad; ta.m_pad = mp; ta.n_ld = md->n_ld; ta.n_valid = md->n;
            ta.W = md->W + (size_t)c0 * md->n_ld; ta.ldw = md->n_ld; ta.kcols = (kc - c0) < 64 ? (kc - c0) : 64;
            ta.bias = md->centered ? md->bias + c0 : nullptr;
            ta.out = od + (size_t)c0 * mp; ta.ldo = mp; ta.kernel_id = md->kernel_id;
            ta.kp[0] = md->kp[0]; ta.kp[1] = md->kp[1]; ta.kp[2] = md->kp[2];
            if (rowmean) { ta.wsum = md->wsum + c0; ta.grand = md->grandmean; }
            tm.begin(&st->transform_ms);
            FX_TRY(fx_launch_transform_ex(ctx, ta, variant, scratch, ns));
            tm.end();
            st->transform_launches++;
        }
        tm.begin(&st->d2h_ms);
        FX_CUDA(cudaMemcpy2DAsync(out + q0, ldo * 8, od, mp * 8, mc * 8, md->k,
                                  on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        tm.end();
        if (!on_dev) st->d2h_bytes += mc * md->k * 8;
    }
    FX_CUDA(cudaEventRecord(e1, s));
    FX_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    st->total_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}
}  // namespace

extern "C" int fx_kpca_transform_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                                     int64_t ldz, double* out, int64_t ldo, fx_stats* stats) {
    return transform_impl(ctx, model, Z, false, m, ldz, out, ldo, stats, 0);
}
extern "C" int fx_kpca_transform_dev_f64(fx_ctx* ctx, const fx_model* model, const double* dZ, int64_t m,
                                         int64_t ldz, double* dout, int64_t ldo, fx_stats* stats) {
    return transform_impl(ctx, model, dZ, true, m, ldz, dout, ldo, stats, 0);
}
extern "C" int fx_klda_transform_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                                     int64_t ldz, double* out, int64_t ldo, fx_stats* stats) {
    return transform_impl(ctx, model, Z, false, m, ldz, out, ldo, stats, 0);
}
// debug / parity: same call through the naive CUDA kernel
extern "C" int fx_kpca_transform_naive_f64(fx_ctx* ctx, const fx_model* model, const double* Z, int64_t m,
                                           int64_t ldz, double* out, int64_t ldo, fx_stats* stats) {
    return transform_impl(ctx, model, Z, false, m, ldz, out, ldo, stats, 1);
}

extern "C" int fx_model_create_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                                   int kernel_id, const double* kparams, const double* W, int k,
                                   int64_t ldw, int center, fx_model** model, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_model_create_f64: ctx is NULL"); return -1; }
    if (!ctx->sub.empty()) {
        // one replica per device
        if (!model) { fx_set_error("model is NULL"); return -12; }
        const int G = (int)ctx->sub.size();
        std::vector<fx_model*> models(G, nullptr);
        std::vector<fx_stats> sts(G);
        const int rc = fx_multi_run(ctx, [&](fx_ctx* c, int r) {
            return fx_model_create_f64(c, X, n, d, ldx, kernel_id, kparams, W, k, ldw, center, &models[r], &sts[r]);
        });
        if (stats) *stats = sts[0];
        if (rc) { for (fx_model* x : models) fx_model_free(x); *model = nullptr; return rc; }
        fx_model* mm = new fx_model();
        mm->sub = models;
        *model = mm;
        return 0;
    }
    if (!X) { fx_set_error("fx_model_create_f64: invalid argument 2 (!X)"); return -2; }
    if (n < 1) { fx_set_error("fx_model_create_f64: invalid argument 3 (n < 1)"); return -3; }
    if (d < 1) { fx_set_error("fx_model_create_f64: invalid argument 4 (d < 1)"); return -4; }
    if (ldx < n) { fx_set_error("model_create: ldx < n (DimensionMismatch)"); return -5; }
    double kpv[3]; int ipow;
    if (!kparams || fx_make_kernel_map(kernel_id, kparams, &ipow, kpv)) {
        fx_set_error("model_create: unknown kernel id or invalid parameters");
        return -6;
    }
    if (!W) { fx_set_error("fx_model_create_f64: invalid argument 8 (!W)"); return -8; }
    if (k < 1 || k > n) { fx_set_error("model_create: k must lie in 1..n"); return -9; }
    if (ldw < n) { fx_set_error("model_create: ldw < n (DimensionMismatch: W must have n rows)"); return -10; }
    if (!model) { fx_set_error("model is NULL"); return -12; }
    *model = nullptr;
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    cudaStream_t s = ctx->stream;
    const int64_t n_ld = round_up(n, 128), d_pad = round_up(d, 16);
    const int kc = (int)round_up(k, 8);
    fx_model* m = new fx_model();
    m->device = ctx->device; m->n = n; m->d = d; m->n_ld = n_ld; m->d_pad = d_pad; m->k = k; m->k_pad = kc;
    m->kernel_id = kernel_id; m->kp[0] = kpv[0]; m->kp[1] = kpv[1]; m->kp[2] = kpv[2];
    m->centered = center ? 1 : 0;
    auto fail = [&](int rc) { fx_model_free(m); return rc; };
    if (cudaMalloc((void**)&m->Xp, (size_t)d_pad * n_ld * 8) != cudaSuccess ||
        cudaMalloc((void**)&m->norms, (size_t)n_ld * 8) != cudaSuccess ||
        cudaMalloc((void**)&m->W, (size_t)kc * n_ld * 8) != cudaSuccess ||
        cudaMalloc((void**)&m->bias, (size_t)kc * 8) != cudaSuccess ||
        cudaMalloc((void**)&m->wsum, 2 * (size_t)kc * 8) != cudaSuccess ||
        cudaMalloc((void**)&m->colmean, (size_t)n_ld * 8) != cudaSuccess) {
        cudaGetLastError();
        fx_set_error("model_create: out of device memory");
        return fail(FX_ERR_OOM);
    }
    int rc = 0;
#define MC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fx_set_error("CUDA error %s (%s)", cudaGetErrorString(_e), #expr); return fail(FX_ERR_CUDA); } } while (0)
    MC(cudaMemsetAsync(m->Xp, 0, (size_t)d_pad * n_ld * 8, s));
    MC(cudaMemcpy2DAsync(m->Xp, n_ld * 8, X, ldx * 8, n * 8, d, cudaMemcpyHostToDevice, s));
    MC(cudaMemsetAsync(m->W, 0, (size_t)kc * n_ld * 8, s));
    MC(cudaMemcpy2DAsync(m->W, n_ld * 8, W, ldw * 8, n * 8, k, cudaMemcpyHostToDevice, s));
    MC(cudaMemsetAsync(m->bias, 0, (size_t)kc * 8, s));
    MC(cudaMemsetAsync(m->wsum, 0, 2 * (size_t)kc * 8, s));
    MC(cudaMemsetAsync(m->colmean, 0, (size_t)n_ld * 8, s));
    if ((rc = fx_launch_rownorms(ctx, m->Xp, n_ld, n_ld, d_pad, m->norms))) return fail(rc);
    st->h2d_bytes = n * d * 8 + n * k * 8;
    if (center) {
        // Column means r = K 1 / n and the grand mean of the training kernel matrix: the row sums of
        // k(X, X) fall out of one in-sample pass of the fused kernel (its projection onto the first 8
        // components is discarded); then bias = r'W, wsum = 1'W.
        int nsplit = 1;
        const size_t scr = fx_transform_scratch(ctx, n_ld, n_ld, 8, &nsplit);
        size_t need = fx_align256((size_t)8 * n_ld * 8) + fx_align256(scr * 8) + 2 * fx_align256((size_t)n_ld * 8) + 8192;
        if ((rc = fx_arena_reserve(ctx, need))) return fail(rc);
        fx_arena_reset(ctx);
        double* T = (double*)fx_arena_alloc(ctx, (size_t)8 * n_ld * 8);
        double* rsum = (double*)fx_arena_alloc(ctx, (size_t)n_ld * 8);
        double* gdev = (double*)fx_arena_alloc(ctx, 256);
        double* scratch = scr ? (double*)fx_arena_alloc(ctx, scr * 8) : nullptr;
        TransformArgs ta;
        ta.Zp = m->Xp; ta.ldz = n_ld; ta.nz = m->norms; ta.Xp = m->Xp; ta.ldx = n_ld; ta.nx = m->norms;
        ta.d_pad = d_pad; ta.m_pad = n_ld; ta.n_ld = n_ld; ta.n_valid = n; ta.W = m->W; ta.ldw = n_ld;
        ta.kcols = 8; ta.bias = nullptr; ta.out = T; ta.ldo = n_ld; ta.kernel_id = kernel_id;
        ta.kp[0] = kpv[0]; ta.kp[1] = kpv[1]; ta.kp[2] = kpv[2];
        ta.rz_out = rsum;
        if ((rc = fx_launch_transform_ex(ctx, ta, 0, scratch, nsplit))) return fail(rc);
        st->transform_launches = 1;
        if ((rc = fx_launch_means(ctx, rsum, n, n_ld, m->colmean, gdev))) return fail(rc);
        if ((rc = fx_launch_bias(ctx, m->colmean, m->W, n_ld, n, kc, m->bias))) return fail(rc);
        MC(cudaMemcpyAsync(ctx->pinned + 160, gdev, sizeof(double), cudaMemcpyDeviceToHost, s));
        if ((rc = fx_model_finish_centering(ctx, m))) return fail(rc);
        m->grandmean = ctx->pinned[160];
    }
    MC(cudaStreamSynchronize(s));
#undef MC
    st->kernel_launches = ctx->launches - launches0;
    *model = m;
    return 0;
}
