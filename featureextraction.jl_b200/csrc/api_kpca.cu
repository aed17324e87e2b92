// api_kpca.cu -- C ABI: KPCA fit (kpca, /root/reference/src/kpca.jl:49-57) and the exposed
// building blocks (Gram, centring, K*V, small eigensolver).
#include <stdlib.h>
#include "fx_kernels.cuh"
#include <math.h>

namespace {

struct TotalTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    int start(fx_ctx* ctx) {
        FX_CUDA(cudaEventCreate(&a));
        FX_CUDA(cudaEventCreate(&b));
        FX_CUDA(cudaEventRecord(a, ctx->stream));
        return 0;
    }
    int stop(fx_ctx* ctx, double* ms) {
        FX_CUDA(cudaEventRecord(b, ctx->stream));
        FX_CUDA(cudaEventSynchronize(b));
        float f = 0.f;
        FX_CUDA(cudaEventElapsedTime(&f, a, b));
        *ms = f;
        return 0;
    }
    ~TotalTimer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

struct ModelGuard {
    fx_model* m = nullptr;
    ~ModelGuard() { if (m) fx_model_free(m); }
    fx_model* release() { fx_model* r = m; m = nullptr; return r; }
};

int dmalloc(double** p, size_t doubles) {
    cudaError_t e = cudaMalloc((void**)p, doubles * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        fx_set_error("out of device memory allocating %.3f GB for the model", doubles * 8 / 1e9);
        return FX_ERR_OOM;
    }
    return 0;
}

}  // namespace

size_t fx_gramstage_bytes(int64_t ld, int64_t n_ld, int64_t d_pad, bool center) {
    size_t t = 0;
    t += fx_align256((size_t)d_pad * n_ld * 8);
    if (n_ld > ld) t += fx_align256((size_t)d_pad * ld * 8) + fx_align256((size_t)d_pad * n_ld * 8);   // sharded upload
    t += fx_align256((size_t)n_ld * 8) * 4;
    t += fx_align256((size_t)ld * n_ld * 8);
    if (center) t += fx_align256((size_t)(n_ld / 64) * ld * 8);
    t += 4096;
    return t;
}

int fx_gramstage_run(fx_ctx* ctx, GramStage& g, const double* X, bool x_dev, int64_t ldx,
                     int kernel_id, const double* kp, bool center, int variant, FxTimer& tm,
                     fx_stats* st) {
    cudaStream_t s = ctx->stream;
    g.Xp = (double*)fx_arena_alloc(ctx, (size_t)g.d_pad * g.n_ld * 8);
    g.norms = (double*)fx_arena_alloc(ctx, (size_t)g.n_ld * 8);
    g.rowsum = (double*)fx_arena_alloc(ctx, (size_t)g.n_ld * 8);
    g.colmean = (double*)fx_arena_alloc(ctx, (size_t)g.n_ld * 8);
    g.scal = (double*)fx_arena_alloc(ctx, 256);
    g.K = (double*)fx_arena_alloc(ctx, (size_t)g.ld * g.n_ld * 8);
    g.rowpart = center ? (double*)fx_arena_alloc(ctx, (size_t)(g.n_ld / 64) * g.ld * 8) : nullptr;
    if (!g.Xp || !g.norms || !g.rowsum || !g.colmean || !g.scal || !g.K || (center && !g.rowpart)) {
        fx_set_error("internal: arena too small for the Gram stage");
        return FX_ERR_OOM;
    }
    // Several ranks: map the peers' arenas so the Gram kernel can share out the symmetric blocks
    // (gram.cu, sym = 2).  Collective; also orders this call after every rank's previous use of
    // its K.  The offsets of K / rowpart inside each arena travel with the handles.
    bool peer_sym = false;
    if (ctx->world > 1 && variant == 0) {
        FX_TRY(fx_peer_exchange(ctx, (size_t)((char*)g.K - ctx->arena),
                                g.rowpart ? (size_t)((char*)g.rowpart - ctx->arena) : (size_t)0));
        peer_sym = ctx->peer_valid;
    }
    if (!x_dev && ctx->world > 1) {
        // Host input on several GPUs: every rank uploads only its own row block over PCIe and the
        // blocks are exchanged over NVLink (all ranks hitting the host with the full X at once is
        // what dominated the end-to-end time at 8 GPUs).
        const int G = ctx->world;
        double* snd = (double*)fx_arena_alloc(ctx, (size_t)g.d_pad * g.ld * 8);
        double* rcv = (double*)fx_arena_alloc(ctx, (size_t)g.d_pad * g.n_ld * 8);
        if (!snd || !rcv) { fx_set_error("internal: arena too small for the sharded upload"); return FX_ERR_OOM; }
        tm.begin(&st->h2d_ms);
        FX_CUDA(cudaMemsetAsync(snd, 0, (size_t)g.d_pad * g.ld * 8, s));
        if (g.rows_valid > 0)
            FX_CUDA(cudaMemcpy2DAsync(snd, g.ld * 8, X + g.row0, ldx * 8, g.rows_valid * 8, g.d,
                                      cudaMemcpyHostToDevice, s));
        tm.end();
        st->h2d_bytes += g.rows_valid * g.d * 8;
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allgather(ctx->comm, snd, rcv, (size_t)g.d_pad * g.ld, s));
        tm.end();
        for (int r = 0; r < G; ++r)
            FX_CUDA(cudaMemcpy2DAsync(g.Xp + (size_t)r * g.ld, g.n_ld * 8, rcv + (size_t)r * g.d_pad * g.ld,
                                      g.ld * 8, g.ld * 8, g.d_pad, cudaMemcpyDeviceToDevice, s));
    } else {
        tm.begin(&st->h2d_ms);
        FX_CUDA(cudaMemsetAsync(g.Xp, 0, (size_t)g.d_pad * g.n_ld * 8, s));
        FX_CUDA(cudaMemcpy2DAsync(g.Xp, g.n_ld * 8, X, ldx * 8, g.n * 8, g.d,
                                  x_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        tm.end();
        if (!x_dev) st->h2d_bytes += g.n * g.d * 8;
    }
    FX_CUDA(cudaMemsetAsync(g.colmean, 0, (size_t)g.n_ld * 8, s));
    FX_CUDA(cudaMemsetAsync(g.scal, 0, 256, s));
    FX_TRY(fx_launch_rownorms(ctx, g.Xp, g.n_ld, g.n_ld, g.d_pad, g.norms));

    GramArgs ga;
    ga.A = g.Xp + g.row0; ga.lda = g.n_ld; ga.B = g.Xp; ga.ldb = g.n_ld;
    ga.na = g.norms + g.row0; ga.nb = g.norms; ga.d_pad = g.d_pad;
    ga.K = g.K; ga.ldk = g.n_ld; ga.rows_pad = g.ld; ga.rows_valid = g.rows_valid;
    ga.cols_valid = g.n; ga.diag_shift = g.row0; ga.rowpart = g.rowpart; ga.sym_source = 1;
    ga.kernel_id = kernel_id; ga.kp[0] = kp[0]; ga.kp[1] = kp[1]; ga.kp[2] = kp[2];
    if (peer_sym) {
        ga.peer_world = ctx->world; ga.peer_rank = ctx->rank;
        for (int c = 0; c < ctx->world; ++c) {
            if (c == ctx->rank) { ga.peerK[c] = g.K; ga.peerRowpart[c] = g.rowpart; continue; }
            ga.peerK[c] = (double*)(ctx->peer_arena[c] + ctx->peer_off[c][0]);
            ga.peerRowpart[c] = g.rowpart ? (double*)(ctx->peer_arena[c] + ctx->peer_off[c][1]) : nullptr;
        }
    }
    tm.begin(&st->gram_ms);
    FX_TRY(fx_launch_gram(ctx, ga, variant));
    tm.end();
    st->gram_launches++;
    if (peer_sym) {
        // the other ranks' mirrored tiles (and row-sum partials) must have landed in this rank's
        // memory before anything reads K: kernel completion + a one-element reduction
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, g.scal + 8, 1, s));
        tm.end();
    }

    if (center) {
        // column sums == row sums (K symmetric): local rows, then all-gather
        double* rs_loc = g.rowsum + g.row0;
        tm.begin(&st->center_ms);
        FX_TRY(fx_launch_rowpart_reduce(ctx, g.rowpart, fx_gram_rowpart_slots(g.n_ld), g.ld, rs_loc));
        tm.end();
        if (ctx->world > 1) {
            tm.begin(&st->comm_ms);
            FX_TRY(fxcomm_allgather(ctx->comm, rs_loc, g.rowsum, (size_t)g.ld, s));
            tm.end();
        }
        if (const char* dd = getenv("FXB200_DEBUG_DUMP")) {
            // diagnostics: the all-gathered row sums and this rank's per-tile partials, one file per fit
            static int fitno = 0;
            std::vector<double> hrs((size_t)g.n_ld), hrp((size_t)fx_gram_rowpart_slots(g.n_ld) * g.ld);
            FX_CUDA(cudaMemcpyAsync(hrs.data(), g.rowsum, hrs.size() * 8, cudaMemcpyDeviceToHost, s));
            FX_CUDA(cudaMemcpyAsync(hrp.data(), g.rowpart, hrp.size() * 8, cudaMemcpyDeviceToHost, s));
            FX_CUDA(cudaStreamSynchronize(s));
            char path[512];
            snprintf(path, sizeof(path), "%s/rowsum_r%d.bin", dd, ctx->rank);
            if (FILE* f = fopen(path, "wb")) { fwrite(hrs.data(), 8, hrs.size(), f); fclose(f); }
            snprintf(path, sizeof(path), "%s/rowpart_r%d.bin", dd, ctx->rank);
            if (FILE* f = fopen(path, "wb")) { fwrite(hrp.data(), 8, hrp.size(), f); fclose(f); }
            ++fitno;
        }
        tm.begin(&st->center_ms);
        FX_TRY(fx_launch_means(ctx, g.rowsum, g.n, g.n_ld, g.colmean, g.scal));
        if (!g.defer_center) {
            FX_TRY(fx_launch_center(ctx, g.K, g.n_ld, g.rows_valid, g.n, g.row0, g.colmean, g.scal));
            st->center_launches++;
        }
        tm.end();
        if (g.defer_center) return 0;
    }
    return fx_gramstage_trace(ctx, g, tm, st);
}

// trace of the (centred) matrix over all ranks -> g.scal[1]
int fx_gramstage_trace(fx_ctx* ctx, GramStage& g, FxTimer& tm, fx_stats* st) {
    FX_TRY(fx_launch_trace(ctx, g.K, g.n_ld, g.rows_valid, g.row0, g.scal + 1));
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, g.scal + 1, 1, ctx->stream));
        tm.end();
    }
    return 0;
}

// Builds the persistent model (own allocations) from arena-resident pieces.
int fx_model_build(fx_ctx* ctx, const GramStage& g, int kernel_id, const double* kp, int k,
                   const double* W /* [kpad][n_ld] device */, bool centered, bool is_klda,
                   fx_model** out) {
    cudaStream_t s = ctx->stream;
    ModelGuard mg;
    mg.m = new fx_model();
    fx_model* m = mg.m;
    m->device = ctx->device; m->n = g.n; m->d = g.d; m->n_ld = g.n_ld; m->d_pad = g.d_pad;
    m->k = k; m->k_pad = (int)round_up(k, 8); m->kernel_id = kernel_id;
    m->kp[0] = kp[0]; m->kp[1] = kp[1]; m->kp[2] = kp[2];
    m->centered = centered ? 1 : 0; m->is_klda = is_klda ? 1 : 0;
    FX_TRY(dmalloc(&m->Xp, (size_t)g.d_pad * g.n_ld));
    FX_TRY(dmalloc(&m->norms, g.n_ld));
    FX_TRY(dmalloc(&m->W, (size_t)m->k_pad * g.n_ld));
    FX_TRY(dmalloc(&m->bias, m->k_pad));
    FX_TRY(dmalloc(&m->wsum, 2 * (size_t)m->k_pad));
    FX_TRY(dmalloc(&m->colmean, g.n_ld));
    FX_CUDA(cudaMemcpyAsync(m->Xp, g.Xp, (size_t)g.d_pad * g.n_ld * 8, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(m->norms, g.norms, (size_t)g.n_ld * 8, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(m->W, W, (size_t)m->k_pad * g.n_ld * 8, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(m->colmean, g.colmean, (size_t)g.n_ld * 8, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemsetAsync(m->bias, 0, (size_t)m->k_pad * 8, s));
    if (centered) {
        FX_TRY(fx_launch_bias(ctx, m->colmean, m->W, g.n_ld, g.n, m->k_pad, m->bias));
        FX_CUDA(cudaMemcpyAsync(ctx->pinned + 160, g.scal, sizeof(double), cudaMemcpyDeviceToHost, s));
        FX_TRY(fx_model_finish_centering(ctx, m));
        m->grandmean = ctx->pinned[160];
    }
    *out = mg.release();
    return 0;
}

// wsum = 1'W and the decision whether the transform needs the per-query row-mean term: a column
// counts as orthogonal to 1 when |1'w| is at rounding level of the sum (64 eps |w|_1).
int fx_model_finish_centering(fx_ctx* ctx, fx_model* m) {
    cudaStream_t s = ctx->stream;
    FX_TRY(fx_launch_colsums(ctx, m->W, m->n_ld, m->n, m->k_pad, m->wsum));
    std::vector<double> h(2 * (size_t)m->k_pad);
    FX_CUDA(cudaMemcpyAsync(h.data(), m->wsum, h.size() * 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    m->need_rowmean = 0;
    for (int c = 0; c < m->k; ++c)
        if (fabs(h[c]) > 64.0 * 2.220446049250313e-16 * h[m->k_pad + c]) m->need_rowmean = 1;
    return 0;
}

static int kpca_fit_impl(fx_ctx* ctx, const double* X, bool on_dev, int64_t n, int64_t d, int64_t ldx,
                         int kernel_id, const double* kparams, int k, int center, double alpha,
                         double eps, const fx_solver_opts* opts, double* V, int64_t ldv,
                         double* lambda, fx_model** model, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (model) *model = nullptr;
    if (!ctx) { fx_set_error("kpca_fit: ctx is NULL"); return -1; }
    if (!X) { fx_set_error("kpca_fit: X is NULL"); return -2; }
    if (n < 2) { fx_set_error("kpca_fit: need n >= 2 observations (got %lld)", (long long)n); return -3; }
    if (d < 1) { fx_set_error("kpca_fit: need d >= 1"); return -4; }
    if (ldx < n) { fx_set_error("kpca_fit: ldx < n (DimensionMismatch)"); return -5; }
    double kpv[3]; int ipow;
    if (!kparams || fx_make_kernel_map(kernel_id, kparams, &ipow, kpv)) {
        fx_set_error("kpca_fit: unknown kernel id %d or invalid kernel parameters", kernel_id);
        return kparams ? -6 : -7;
    }
    if (k < 1 || k > n) { fx_set_error("kpca_fit: k=%d outside 1..n", k); return -8; }
    if (!(alpha >= 0.0 && alpha <= 1.0)) { fx_set_error("kpca_fit: alpha must lie in [0,1]"); return -10; }
    if (!(eps >= 0.0)) { fx_set_error("kpca_fit: eps must be >= 0"); return -11; }
    const bool mute = ctx->parent && ctx->rank != 0;       // in-process rank other than the leader: no host output
    if (!V && !mute) { fx_set_error("kpca_fit: V is NULL"); return -13; }
    if (ldv < n) { fx_set_error("kpca_fit: ldv < n (DimensionMismatch)"); return -14; }
    if (!lambda && !mute) { fx_set_error("kpca_fit: lambda is NULL"); return -15; }

    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    fx_solver_opts o;
    fx_resolve_opts(opts, k, &o);
    if (const char* e = getenv("FXB200_BLOCK")) {                  // tuning knob for experiments
        const int v = atoi(e);
        if (v >= k && v <= 64 && (!opts || opts->block <= 0)) o.block = (int)round_up(v, 8);
    }
    const int G = ctx->world;
    const int64_t ld = round_up(ceil_div(n, G), 128), n_ld = ld * G;
    if (!opts || opts->max_blocks <= 1) {
        // basis depth: a deeper basis saves K*V passes but makes every Rayleigh-Ritz (device
        // Jacobi, ~m^3, single CTA) dearer.  The Rayleigh-Ritz runs beside the next pass, so it is
        // free while it is shorter than a solver step: 5 blocks while a pass over this rank's K
        // costs > 2 ms, 4 down to 0.3 ms (measured on 8 GPUs at n = 65536: 18.1 / 17.5 / 18.5 ms
        // per fit with 3 / 4 / 5), 3 below that
        const double pass_ms = 8.0 * (double)ld * (double)n_ld / 6.5e9;
        o.max_blocks = pass_ms > 2.0 ? 5 : (pass_ms > 0.3 ? 4 : 3);
        if (const char* e = getenv("FXB200_MAX_BLOCKS")) {        // tuning knob for experiments
            const int v = atoi(e);
            if (v >= 2 && v <= 8) o.max_blocks = v;
        }
    }
    const int64_t row0 = ld * ctx->rank;
    int64_t rows_valid = n - row0;
    if (rows_valid < 0) rows_valid = 0;
    if (rows_valid > ld) rows_valid = ld;
    const int64_t d_pad = round_up(d, 16);
    const int kp = (int)round_up(k, 8);
    // Dense paths (all n pairs available, like the reference's dsyevd, common.jl:161-165): two-sided
    // Jacobi in one CTA for n <= 1024; one-sided (Hestenes) Jacobi over all SMs for 1024 < n <= 8192
    // when many components are wanted (k > 1024 or k > n/8).  Otherwise the block Krylov solver: one
    // group for k <= 64, groups of 48 with locking above that.
    const bool dense_small = (n <= 256) || (n < 8 * (int64_t)o.block) || (k > 64 && n <= 1024);
    const bool dense_big = !dense_small && n <= 8192 && (k > 1024 || 8 * (int64_t)k > n) && G == 1;
    const bool dense = dense_small || dense_big;
    const bool grouped = !dense && k > 64;
    if (grouped && k > 1024) {
        fx_set_error("kpca_fit: k=%d > 1024 components is supported up to n = 8192 on one GPU (dense path); "
                     "for larger n the solver returns at most 1024 leading pairs", k);
        return -8;
    }
    if (!dense) {
        int64_t mb = n / (2 * o.block) - 1;
        if (mb < 2) mb = 2;
        if (o.max_blocks > mb) o.max_blocks = (int)mb;
    }

    size_t need = fx_gramstage_bytes(ld, n_ld, d_pad, center != 0);
    need += fx_align256((size_t)kp * n_ld * 8) + fx_align256((size_t)(kp + 8) * 8) + 4096;
    if (dense) need += 3 * fx_align256((size_t)n_ld * n * 8) + fx_align256((size_t)n_ld * n_ld * 8) + 65536;
    else need += fx_eig_workspace_bytes(ctx, ld, n_ld, k, o) + (grouped ? fx_align256((size_t)kp * ld * 8) : 0);
    FX_TRY(fx_agree(ctx, fx_arena_reserve(ctx, need)));          // all ranks or none enter the collectives
    fx_arena_reset(ctx);

    TotalTimer tot;
    FX_TRY(tot.start(ctx));
    FxTimer tm(ctx);
    cudaStream_t s = ctx->stream;

    GramStage g;
    g.n = n; g.d = d; g.ld = ld; g.n_ld = n_ld; g.d_pad = d_pad; g.row0 = row0; g.rows_valid = rows_valid;
    // the block solver's first K*V pass does the centring (one read of K for both); the dense
    // path and the naive cross-check variant centre in a pass of their own
    g.defer_center = center != 0 && !dense && o.variant == 0;
    FX_TRY(fx_gramstage_run(ctx, g, X, on_dev, ldx, kernel_id, kpv, center != 0, o.variant, tm, st));

    double* Vout = (double*)fx_arena_alloc(ctx, (size_t)kp * n_ld * 8);
    double* theta = (double*)fx_arena_alloc(ctx, (size_t)(kp + 8) * 8);
    if (!Vout || !theta) { fx_set_error("internal: arena too small"); return FX_ERR_OOM; }

    EigArgs ea;
    ea.K = g.K; ea.ldk = n_ld; ea.n = n; ea.ld = ld; ea.n_ld = n_ld; ea.row0 = row0;
    ea.rows_valid = rows_valid; ea.k = k; ea.opts = o; ea.Vout = Vout; ea.theta = theta;
    if (g.defer_center) { ea.center_colmean = g.colmean; ea.center_grand = g.scal; }
    tm.begin(&st->eig_ms);
    int eig_rc = 0;
    if (dense) {
        eig_rc = fx_eig_dense(ctx, ea, tm, st);
    } else if (!grouped) {
        eig_rc = fx_eig_topk(ctx, ea, tm, st);
    } else {
        // k > 64: the spectrum is taken in groups of 48 (block width 64: 16 guard vectors, so that a group
        // does not wait for the pair at its own edge).  Converged eigenvectors are LOCKED: every later
        // group iterates in their orthogonal complement, i.e. on the deflated operator, whose leading
        // pairs are the next ones of K~ (descending order overall, like common.jl:164).  Deep in the
        // spectrum the gaps are small and a restarted Krylov method converges slowly, so unless the
        // caller sets one, the Ritz-residual tolerance of this mode is 1e-9 |lambda_1| (eigenvalue error
        // ~ residual^2 / gap stays far below the 1e-10 target) with up to 600 passes per group.
        double* Lloc = (double*)fx_arena_alloc(ctx, (size_t)kp * ld * 8);
        if (!Lloc) { fx_set_error("internal: arena too small"); return FX_ERR_OOM; }
        FX_CUDA(cudaMemsetAsync(Lloc, 0, (size_t)kp * ld * 8, s));
        const size_t mark = ctx->arena_off;
        int done = 0, steps = 0, restarts = 0;
        double scale_ref = 0.0, worst = 0.0;
        fx_solver_opts og = o;
        if (!opts || !(opts->tol > 0.0)) og.tol = 1e-9;
        if (!opts || opts->max_steps <= 0) og.max_steps = 600;
        while (done < k && (eig_rc == 0 || eig_rc == FX_ERR_NOT_CONVERGED)) {
            const int kg = (k - done) < 48 ? (k - done) : 48;
            EigArgs eg = ea;
            eg.opts = og;
            eg.k = kg; eg.Vout = Vout + (size_t)done * n_ld; eg.theta = theta + done;
            eg.lock = Lloc; eg.nlock = done; eg.scale_ref = scale_ref;
            if (done > 0) { eg.center_colmean = nullptr; eg.center_grand = nullptr; }   // K is centred by now
            ctx->arena_off = mark;                                  // the solver carves its workspace anew
            const int rc = fx_eig_topk(ctx, eg, tm, st);
            if (rc != 0 && rc != FX_ERR_NOT_CONVERGED) { eig_rc = rc; break; }
            if (rc == FX_ERR_NOT_CONVERGED) eig_rc = rc;
            steps += st->steps; restarts += st->restarts;
            if (st->max_residual > worst) worst = st->max_residual;
            const int kgp = (int)round_up(kg, 8);
            FX_CUDA(cudaMemcpy2DAsync(Lloc + (size_t)done * ld, ld * 8, eg.Vout + row0, n_ld * 8, ld * 8, kgp,
                                      cudaMemcpyDeviceToDevice, s));
            if (done == 0) {
                FX_CUDA(cudaMemcpyAsync(ctx->pinned + 96, theta, sizeof(double), cudaMemcpyDeviceToHost, s));
                FX_CUDA(cudaStreamSynchronize(s));
                scale_ref = fabs(ctx->pinned[96]);
            }
            done += kg;
        }
        st->steps = steps; st->restarts = restarts; st->max_residual = worst;
        st->converged = (eig_rc == 0) ? 1 : 0;
    }
    tm.end();
    if (eig_rc != 0 && eig_rc != FX_ERR_NOT_CONVERGED) return eig_rc;
    if (g.defer_center) FX_TRY(fx_gramstage_trace(ctx, g, tm, st));

    // eigenvalue post-process: regularize!/perturb! (kpca.jl:53-54) only remap the spectrum,
    // lambda' = ((1-alpha) theta + alpha tr(K~)/n + eps (n-1)) / (n-1)   (kpca.jl:56)
    double* hp = ctx->pinned + 2048;                 // k <= 1024 values (pinned scratch holds 4096)
    double* hs = ctx->pinned + 128;                  // grand mean, trace
    FX_CUDA(cudaMemcpyAsync(hp, theta, k * sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(hs, g.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    const double trace = hs[1];
    st->trace = trace;
    std::vector<double> lam(k);
    for (int i = 0; i < k; ++i)
        lam[i] = ((1.0 - alpha) * hp[i] + alpha * trace / (double)n + eps * (double)(n - 1)) / (double)(n - 1);
    tm.begin(&st->d2h_ms);
    if (mute) {
        // results leave through the leader rank
    } else if (on_dev) {
        for (int i = 0; i < k; ++i) hp[i] = lam[i];
        FX_CUDA(cudaMemcpyAsync(lambda, hp, k * sizeof(double), cudaMemcpyHostToDevice, s));
        FX_CUDA(cudaMemcpy2DAsync(V, ldv * 8, Vout, n_ld * 8, n * 8, k, cudaMemcpyDeviceToDevice, s));
    } else {
        for (int i = 0; i < k; ++i) lambda[i] = lam[i];
        FX_CUDA(cudaMemcpy2DAsync(V, ldv * 8, Vout, n_ld * 8, n * 8, k, cudaMemcpyDeviceToHost, s));
        st->d2h_bytes += n * k * 8 + k * 8;
    }
    tm.end();
    if (model) {
        if (ctx->parent) {
            // in-process ranks: nobody allocates (the model) while a peer may still be inside a collective
            FX_CUDA(cudaStreamSynchronize(s));
            ctx->parent->ms->barrier();
        }
        FX_TRY(fx_model_build(ctx, g, kernel_id, kpv, k, Vout, center != 0, false, model));
    }
    FX_TRY(tot.stop(ctx, &st->total_ms));
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return eig_rc;
}

extern "C" int fx_kpca_fit_f64(fx_ctx* ctx, const double* X, int64_t n, int64_t d, int64_t ldx,
                               int kernel_id, const double* kparams, int k, int center, double alpha,
                               double eps, const fx_solver_opts* opts, double* V, int64_t ldv,
                               double* lambda, fx_model** model, fx_stats* stats) {
    if (ctx && !ctx->sub.empty()) {
        // single-process multi-GPU: one host thread per device runs the rank's share of the fit
        if (!V) { fx_set_error("kpca_fit: V is NULL"); return -13; }
        if (!lambda) { fx_set_error("kpca_fit: lambda is NULL"); return -15; }
        const int G = (int)ctx->sub.size();
        std::vector<fx_model*> models(G, nullptr);
        std::vector<fx_stats> sts(G);
        if (model) *model = nullptr;
        const int rc = fx_multi_run(ctx, [&](fx_ctx* c, int r) {
            return kpca_fit_impl(c, X, false, n, d, ldx, kernel_id, kparams, k, center, alpha, eps, opts,
                                 r == 0 ? V : nullptr, ldv, r == 0 ? lambda : nullptr, model ? &models[r] : nullptr,
                                 &sts[r]);
        });
        if (stats) *stats = sts[0];
        if ((rc == 0 || rc == FX_ERR_NOT_CONVERGED) && model) {
            fx_model* m = new fx_model();
            m->sub = models;
            *model = m;
        } else {
            for (fx_model* m : models) fx_model_free(m);
        }
        return rc;
    }
    return kpca_fit_impl(ctx, X, false, n, d, ldx, kernel_id, kparams, k, center, alpha, eps, opts, V,
                         ldv, lambda, model, stats);
}

extern "C" int fx_kpca_fit_dev_f64(fx_ctx* ctx, const double* dX, int64_t n, int64_t d, int64_t ldx,
                                   int kernel_id, const double* kparams, int k, int center,
                                   double alpha, double eps, const fx_solver_opts* opts, double* dV,
                                   int64_t ldv, double* dlambda, fx_model** model, fx_stats* stats) {
    return kpca_fit_impl(ctx, dX, true, n, d, ldx, kernel_id, kparams, k, center, alpha, eps, opts, dV,
                         ldv, dlambda, model, stats);
}

// ---------------------------------------------------------------------------------------------
// building blocks
// ---------------------------------------------------------------------------------------------
extern "C" int fx_gram_f64(fx_ctx* ctx, const double* Z, int64_t m, int64_t ldz, const double* X,
                           int64_t n, int64_t ldx, int64_t d, int kernel_id, const double* kparams,
                           int variant, double* K, int64_t ldk, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_gram_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    const bool sym = (Z == nullptr);
    if (sym) { m = n; }
    if (m < 1) { fx_set_error("fx_gram_f64: invalid argument 2 (m < 1)"); return -2; }
    if (!sym && ldz < m) { fx_set_error("fx_gram_f64: invalid argument 3 (!sym && ldz < m; DimensionMismatch)"); return -3; }
    if (!X) { fx_set_error("fx_gram_f64: invalid argument 4 (!X)"); return -4; }
    if (n < 1) { fx_set_error("fx_gram_f64: invalid argument 5 (n < 1)"); return -5; }
    if (ldx < n) { fx_set_error("fx_gram_f64: invalid argument 6 (ldx < n; DimensionMismatch)"); return -6; }
    if (d < 1) { fx_set_error("fx_gram_f64: invalid argument 7 (d < 1)"); return -7; }
    double kpv[3]; int ipow;
    if (!kparams || fx_make_kernel_map(kernel_id, kparams, &ipow, kpv)) {
        fx_set_error("gram: unknown kernel id or invalid parameters");
        return -8;
    }
    if (!K) { fx_set_error("fx_gram_f64: invalid argument 11 (!K)"); return -11; }
    if (ldk < m) { fx_set_error("fx_gram_f64: invalid argument 12 (ldk < m; DimensionMismatch)"); return -12; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    // device result is row-major [n_pad][m_ld] = column-major m x n with leading dimension m_ld
    const int64_t n_pad = round_up(n, 128), m_ld = round_up(m, 128), d_pad = round_up(d, 16);
    size_t need = fx_align256((size_t)d_pad * n_pad * 8) + fx_align256((size_t)d_pad * m_ld * 8) +
                  fx_align256((size_t)n_pad * 8) + fx_align256((size_t)m_ld * 8) +
                  fx_align256((size_t)n_pad * m_ld * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Xp = (double*)fx_arena_alloc(ctx, (size_t)d_pad * n_pad * 8);
    double* Zp = sym ? Xp : (double*)fx_arena_alloc(ctx, (size_t)d_pad * m_ld * 8);
    double* nx = (double*)fx_arena_alloc(ctx, (size_t)n_pad * 8);
    double* nz = sym ? nx : (double*)fx_arena_alloc(ctx, (size_t)m_ld * 8);
    double* Kd = (double*)fx_arena_alloc(ctx, (size_t)n_pad * m_ld * 8);
    TotalTimer tot;
    FX_TRY(tot.start(ctx));
    FxTimer tm(ctx);
    FX_CUDA(cudaMemsetAsync(Xp, 0, (size_t)d_pad * n_pad * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(Xp, n_pad * 8, X, ldx * 8, n * 8, d, cudaMemcpyHostToDevice, s));
    FX_TRY(fx_launch_rownorms(ctx, Xp, n_pad, n_pad, d_pad, nx));
    if (!sym) {
        FX_CUDA(cudaMemsetAsync(Zp, 0, (size_t)d_pad * m_ld * 8, s));
        FX_CUDA(cudaMemcpy2DAsync(Zp, m_ld * 8, Z, ldz * 8, m * 8, d, cudaMemcpyHostToDevice, s));
        FX_TRY(fx_launch_rownorms(ctx, Zp, m_ld, m_ld, d_pad, nz));
    }
    GramArgs ga;
    ga.A = Xp; ga.lda = n_pad; ga.B = Zp; ga.ldb = sym ? n_pad : m_ld; ga.na = nx; ga.nb = nz;
    ga.d_pad = d_pad; ga.K = Kd; ga.ldk = m_ld; ga.rows_pad = n_pad; ga.rows_valid = n;
    ga.cols_valid = m; ga.diag_shift = sym ? 0 : -1; ga.rowpart = nullptr; ga.sym_source = sym ? 1 : 0;
    ga.kernel_id = kernel_id; ga.kp[0] = kpv[0]; ga.kp[1] = kpv[1]; ga.kp[2] = kpv[2];
    tm.begin(&st->gram_ms);
    FX_TRY(fx_launch_gram(ctx, ga, variant));
    tm.end();
    st->gram_launches = 1;
    FX_CUDA(cudaMemcpy2DAsync(K, ldk * 8, Kd, m_ld * 8, m * 8, n, cudaMemcpyDeviceToHost, s));
    FX_TRY(tot.stop(ctx, &st->total_ms));
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

extern "C" int fx_center_f64(fx_ctx* ctx, double* K, int64_t n, int64_t ldk, double* colmean,
                             double* grandmean, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_center_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!K) { fx_set_error("fx_center_f64: invalid argument 2 (!K)"); return -2; }
    if (n < 1) { fx_set_error("fx_center_f64: invalid argument 3 (n < 1)"); return -3; }
    if (ldk < n) { fx_set_error("center: ldk < n (DimensionMismatch)"); return -4; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int64_t n_ld = round_up(n, 64);
    size_t need = fx_align256((size_t)n * n_ld * 8) + 3 * fx_align256((size_t)n_ld * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Kd = (double*)fx_arena_alloc(ctx, (size_t)n * n_ld * 8);
    double* rs = (double*)fx_arena_alloc(ctx, (size_t)n_ld * 8);
    double* cm = (double*)fx_arena_alloc(ctx, (size_t)n_ld * 8);
    double* sc = (double*)fx_arena_alloc(ctx, 256);
    TotalTimer tot;
    FX_TRY(tot.start(ctx));
    FxTimer tm(ctx);
    FX_CUDA(cudaMemsetAsync(Kd, 0, (size_t)n * n_ld * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(Kd, n_ld * 8, K, ldk * 8, n * 8, n, cudaMemcpyHostToDevice, s));
    tm.begin(&st->center_ms);
    FX_TRY(fx_launch_rowsums(ctx, Kd, n_ld, n, n, rs));
    FX_TRY(fx_launch_means(ctx, rs, n, n_ld, cm, sc));
    FX_TRY(fx_launch_center(ctx, Kd, n_ld, n, n, 0, cm, sc));
    tm.end();
    st->center_launches = 1;
    FX_CUDA(cudaMemcpy2DAsync(K, ldk * 8, Kd, n_ld * 8, n * 8, n, cudaMemcpyDeviceToHost, s));
    if (colmean) FX_CUDA(cudaMemcpyAsync(colmean, cm, n * 8, cudaMemcpyDeviceToHost, s));
    if (grandmean) FX_CUDA(cudaMemcpyAsync(grandmean, sc, 8, cudaMemcpyDeviceToHost, s));
    FX_TRY(tot.stop(ctx, &st->total_ms));
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

extern "C" int fx_symm_f64(fx_ctx* ctx, const double* K, int64_t n, int64_t ldk, const double* V, int b,
                           int64_t ldv, double* Y, int64_t ldy, int variant, fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_symm_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!K) { fx_set_error("fx_symm_f64: invalid argument 2 (!K)"); return -2; }
    if (n < 1) { fx_set_error("fx_symm_f64: invalid argument 3 (n < 1)"); return -3; }
    if (ldk < n) { fx_set_error("fx_symm_f64: invalid argument 4 (ldk < n; DimensionMismatch)"); return -4; }
    if (!V) { fx_set_error("fx_symm_f64: invalid argument 5 (!V)"); return -5; }
    if (b < 1 || b > 64) { fx_set_error("symm: block width must be 1..64"); return -6; }
    if (ldv < n) { fx_set_error("fx_symm_f64: invalid argument 7 (ldv < n; DimensionMismatch)"); return -7; }
    if (!Y) { fx_set_error("fx_symm_f64: invalid argument 8 (!Y)"); return -8; }
    if (ldy < n) { fx_set_error("fx_symm_f64: invalid argument 9 (ldy < n; DimensionMismatch)"); return -9; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int64_t np = round_up(n, 128);
    const int bp = (int)round_up(b, 8);
    const int splits = fx_symm_splits(ctx, np, np);
    size_t need = fx_align256((size_t)np * np * 8) + 2 * fx_align256((size_t)bp * np * 8) +
                  fx_align256((size_t)splits * bp * np * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Kd = (double*)fx_arena_alloc(ctx, (size_t)np * np * 8);
    double* Vd = (double*)fx_arena_alloc(ctx, (size_t)bp * np * 8);
    double* Yd = (double*)fx_arena_alloc(ctx, (size_t)bp * np * 8);
    double* Yp = (double*)fx_arena_alloc(ctx, (size_t)splits * bp * np * 8);
    TotalTimer tot;
    FX_TRY(tot.start(ctx));
    FxTimer tm(ctx);
    FX_CUDA(cudaMemsetAsync(Kd, 0, (size_t)np * np * 8, s));
    FX_CUDA(cudaMemsetAsync(Vd, 0, (size_t)bp * np * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(Kd, np * 8, K, ldk * 8, n * 8, n, cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpy2DAsync(Vd, np * 8, V, ldv * 8, n * 8, b, cudaMemcpyHostToDevice, s));
    SymmArgs sa;
    sa.K = Kd; sa.ldk = np; sa.rows_pad = np; sa.V = Vd; sa.ldr = np; sa.v_rank_stride = (int64_t)bp * np;
    sa.bcols = bp; sa.Y = Yd; sa.Ypart = Yp; sa.splits = splits;
    tm.begin(&st->symm_ms);
    FX_TRY(fx_launch_symm(ctx, sa, variant));
    tm.end();
    st->symm_launches = 1;
    FX_CUDA(cudaMemcpy2DAsync(Y, ldy * 8, Yd, np * 8, n * 8, b, cudaMemcpyDeviceToHost, s));
    FX_TRY(tot.stop(ctx, &st->total_ms));
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

// The fused centring pass of a fit, exposed for elementwise parity tests: K is centred IN PLACE by
// the same kernel instance that computes Y = K~ V (symm.cu, CENTER variant: centred in registers,
// written back through the TMA store warp).  fx_center_f64 exercises the standalone kernel.
extern "C" int fx_center_symm_f64(fx_ctx* ctx, double* K, int64_t n, int64_t ldk, const double* V, int b,
                                  int64_t ldv, double* Y, int64_t ldy, double* colmean, double* grandmean,
                                  fx_stats* stats) {
    fx_stats local;
    fx_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof(*st));
    if (!ctx) { fx_set_error("fx_center_symm_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!K) { fx_set_error("fx_center_symm_f64: invalid argument 2 (!K)"); return -2; }
    if (n < 1) { fx_set_error("fx_center_symm_f64: invalid argument 3 (n < 1)"); return -3; }
    if (ldk < n) { fx_set_error("fx_center_symm_f64: invalid argument 4 (ldk < n; DimensionMismatch)"); return -4; }
    if (!V) { fx_set_error("fx_center_symm_f64: invalid argument 5 (!V)"); return -5; }
    if (b < 1 || b > 64) { fx_set_error("fx_center_symm_f64: block width must be 1..64"); return -6; }
    if (ldv < n) { fx_set_error("fx_center_symm_f64: invalid argument 7 (ldv < n; DimensionMismatch)"); return -7; }
    if (!Y) { fx_set_error("fx_center_symm_f64: invalid argument 8 (!Y)"); return -8; }
    if (ldy < n) { fx_set_error("fx_center_symm_f64: invalid argument 9 (ldy < n; DimensionMismatch)"); return -9; }
    FX_CUDA(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int64_t np = round_up(n, 128);
    const int bp = (int)round_up(b, 8);
    const int splits = fx_symm_splits(ctx, np, np);
    size_t need = fx_align256((size_t)np * np * 8) + 2 * fx_align256((size_t)bp * np * 8) +
                  fx_align256((size_t)splits * bp * np * 8) + 3 * fx_align256((size_t)np * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Kd = (double*)fx_arena_alloc(ctx, (size_t)np * np * 8);
    double* Vd = (double*)fx_arena_alloc(ctx, (size_t)bp * np * 8);
    double* Yd = (double*)fx_arena_alloc(ctx, (size_t)bp * np * 8);
    double* Yp = (double*)fx_arena_alloc(ctx, (size_t)splits * bp * np * 8);
    double* rs = (double*)fx_arena_alloc(ctx, (size_t)np * 8);
    double* cm = (double*)fx_arena_alloc(ctx, (size_t)np * 8);
    double* sc = (double*)fx_arena_alloc(ctx, 256);
    TotalTimer tot;
    FX_TRY(tot.start(ctx));
    FxTimer tm(ctx);
    FX_CUDA(cudaMemsetAsync(Kd, 0, (size_t)np * np * 8, s));
    FX_CUDA(cudaMemsetAsync(Vd, 0, (size_t)bp * np * 8, s));
    FX_CUDA(cudaMemsetAsync(rs, 0, (size_t)np * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(Kd, np * 8, K, ldk * 8, n * 8, n, cudaMemcpyHostToDevice, s));
    FX_CUDA(cudaMemcpy2DAsync(Vd, np * 8, V, ldv * 8, n * 8, b, cudaMemcpyHostToDevice, s));
    tm.begin(&st->center_ms);
    FX_TRY(fx_launch_rowsums(ctx, Kd, np, n, n, rs));
    FX_TRY(fx_launch_means(ctx, rs, n, np, cm, sc));
    SymmArgs sa;
    sa.K = Kd; sa.ldk = np; sa.rows_pad = np; sa.V = Vd; sa.ldr = np; sa.v_rank_stride = (int64_t)bp * np;
    sa.bcols = bp; sa.Y = Yd; sa.Ypart = Yp; sa.splits = splits;
    sa.center_colmean = cm; sa.center_grand = sc; sa.rows_valid = n; sa.cols_valid = n; sa.row0 = 0;
    FX_TRY(fx_launch_symm(ctx, sa, 0));
    tm.end();
    st->center_launches = 1;
    FX_CUDA(cudaMemcpy2DAsync(K, ldk * 8, Kd, np * 8, n * 8, n, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpy2DAsync(Y, ldy * 8, Yd, np * 8, n * 8, b, cudaMemcpyDeviceToHost, s));
    if (colmean) FX_CUDA(cudaMemcpyAsync(colmean, cm, n * 8, cudaMemcpyDeviceToHost, s));
    if (grandmean) FX_CUDA(cudaMemcpyAsync(grandmean, sc, 8, cudaMemcpyDeviceToHost, s));
    FX_TRY(tot.stop(ctx, &st->total_ms));
    tm.resolve();
    st->kernel_launches = ctx->launches - launches0;
    return 0;
}

extern "C" int fx_syevd_small_f64(fx_ctx* ctx, double* A, int m, int64_t lda, double* w, int* sweeps) {
    if (!ctx) { fx_set_error("fx_syevd_small_f64: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (!A) { fx_set_error("fx_syevd_small_f64: invalid argument 2 (!A)"); return -2; }
    if (m < 1 || m > 1024) { fx_set_error("syevd_small: order must be 1..1024"); return -3; }
    if (lda < m) { fx_set_error("syevd_small: lda < m (DimensionMismatch)"); return -4; }
    if (!w) { fx_set_error("fx_syevd_small_f64: invalid argument 5 (!w)"); return -5; }
    FX_CUDA(cudaSetDevice(ctx->device));
    size_t need = 3 * fx_align256((size_t)m * m * 8) + fx_align256((size_t)m * 8) + 8192;
    FX_TRY(fx_arena_reserve(ctx, need));
    fx_arena_reset(ctx);
    cudaStream_t s = ctx->stream;
    double* Ad = (double*)fx_arena_alloc(ctx, (size_t)m * m * 8);
    double* Sd = (double*)fx_arena_alloc(ctx, (size_t)m * m * 8);
    double* Wk = (double*)fx_arena_alloc(ctx, (size_t)m * m * 8);
    double* wd = (double*)fx_arena_alloc(ctx, (size_t)m * 8);
    int* info = (int*)fx_arena_alloc(ctx, 64);
    FX_CUDA(cudaMemcpy2DAsync(Ad, (size_t)m * 8, A, lda * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice, s));
    FX_TRY(fx_launch_jacobi(ctx, Ad, m, m, Sd, m, wd, Wk, info));
    FX_CUDA(cudaMemcpy2DAsync(A, lda * 8, Sd, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaMemcpyAsync(w, wd, (size_t)m * 8, cudaMemcpyDeviceToHost, s));
    int hinfo[2] = {0, 0};
    FX_CUDA(cudaMemcpyAsync(hinfo, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    if (sweeps) *sweeps = hinfo[0];
    if (!hinfo[1]) { fx_set_error("syevd_small: Jacobi did not converge"); return FX_ERR_NOT_CONVERGED; }
    return 0;
}
