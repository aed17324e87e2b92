// solver.cu -- top-k symmetric eigensolve on the device: thick-restart block Krylov
// (Krylov-Schur form) whose only heavy operation is Y = K V (symm.cu), with full CGS2
// re-orthogonalisation on DMMA tall-skinny kernels (tsgemm.cu), CholQR / SVQB block
// orthonormalisation and Rayleigh-Ritz by a device Jacobi solver (smalldense.cu) that runs on a
// side stream beside the next K V pass.
//
// Replaces components_eig! -> syevd! -> ccall dsyevd_ (/root/reference/src/common.jl:161-165,
// :107-143, :128-131), which computes ALL n eigenpairs in O(n^3); the caller of kpca
// (kpca.jl:55) only ever uses the leading ones.  Output order is descending like
// common.jl:164.  tests/proto_krylov.py is the step-for-step NumPy model of this file.
//
// Multi-GPU: K, Q, W are row-sharded; per step one all-gather of the n x b block and a few
// all-reduces of (m+b) x b / b x b matrices; the small dense work runs redundantly on every
// rank from identical all-reduced inputs, so all ranks take identical decisions.
#include <stdlib.h>
#include "fx_kernels.cuh"

void fx_resolve_opts(const fx_solver_opts* in, int k, fx_solver_opts* out) {
    fx_solver_opts o;
    memset(&o, 0, sizeof(o));
    if (in) o = *in;
    int kb = (int)round_up(k, 8);
    if (o.block <= 0) o.block = kb;
    if (o.block < kb) o.block = kb;
    o.block = (int)round_up(o.block, 8);
    if (o.block > 64) o.block = 64;
    if (o.max_blocks <= 1) o.max_blocks = 3;
    if (o.max_steps <= 0) o.max_steps = 300;
    if (!(o.tol > 0.0)) o.tol = 1e-12;
    if (o.seed == 0) o.seed = 0x5eedULL;
    if (o.cg_max_iters <= 0) o.cg_max_iters = 4000;
    if (!(o.cg_tol > 0.0)) o.cg_tol = 1e-10;
    *out = o;
}

constexpr int FX_MAX_LOCK = 1024;                   // locked vectors a call can be deflated against

namespace {
struct Work {
    int b, mcap, mt;
    double *Q, *Wa, *Wb, *Vg, *Ypart, *T, *Tw, *S, *Jw, *theta, *H, *H2, *HL, *G, *Gs, *U, *Uw, *lamb,
        *dscale, *M1, *C1, *M2, *C2, *C, *res, *tnscr;
    int* info;
    size_t tnscr_doubles;
    int splits;
};

size_t carve(fx_ctx* ctx, int64_t ld, int64_t n_ld, int k, const fx_solver_opts& o, Work* w, bool commit) {
    const int b = o.block, mcap = o.max_blocks * b, mt = mcap + b;
    size_t total = 0;
    auto take = [&](size_t doubles) -> double* {
        const size_t bytes = fx_align256(doubles * sizeof(double));
        total += bytes;
        return commit ? (double*)fx_arena_alloc(ctx, bytes) : nullptr;
    };
    const int splits = fx_symm_splits(ctx, ld, n_ld);
    size_t tn = 0;
    for (int ma = 8; ma <= (mt > FX_MAX_LOCK ? mt : FX_MAX_LOCK); ma += 8) {   // chunk count depends on ma: take the max
        const size_t t1 = fx_ts_tn_scratch(ctx, ma, b, ld);
        if (t1 > tn) tn = t1;
    }
    double* Q = take((size_t)mt * ld);
    double* Wa = take((size_t)b * ld);
    double* Wb = take((size_t)b * ld);
    double* Vg = ctx->world > 1 ? take((size_t)ctx->world * b * ld) : nullptr;
    double* Yp = splits > 1 ? take((size_t)splits * b * ld) : nullptr;
    double* T = take((size_t)mt * mt);
    double* Tw = take((size_t)mt * mt);
    double* S = take((size_t)mt * mt);
    double* Jw = take((size_t)mt * mt);
    double* theta = take(mt);
    double* H = take((size_t)mt * b);
    double* H2 = take((size_t)(mt + b) * b);        // stacked [Q W]'W: mq coefficient rows + the b x b Gram
    double* HL = take((size_t)FX_MAX_LOCK * b);     // coefficients against the locked vectors
    double* G = take(64 * 64); double* Gs = take(64 * 64); double* U = take(64 * 64);
    double* Uw = take(64 * 64); double* lamb = take(64); double* dscale = take(64);
    double* M1 = take(64 * 64); double* C1 = take(64 * 64); double* M2 = take(64 * 64);
    double* C2 = take(64 * 64); double* C = take(64 * 64);
    double* res = take(64 + 8);
    double* tnscr = take(tn);
    int* info = (int*)take(8);
    if (w) {
        w->b = b; w->mcap = mcap; w->mt = mt; w->Q = Q; w->Wa = Wa; w->Wb = Wb; w->Vg = Vg;
        w->Ypart = Yp; w->T = T; w->Tw = Tw; w->S = S; w->Jw = Jw; w->theta = theta; w->H = H;
        w->H2 = H2; w->HL = HL; w->G = G; w->Gs = Gs; w->U = U; w->Uw = Uw; w->lamb = lamb; w->dscale = dscale;
        w->M1 = M1; w->C1 = C1; w->M2 = M2; w->C2 = C2; w->C = C; w->res = res; w->tnscr = tnscr;
        w->tnscr_doubles = tn; w->info = info; w->splits = splits;
    }
    return total;
}
}  // namespace

size_t fx_eig_workspace_bytes(fx_ctx* ctx, int64_t ld, int64_t n_ld, int k, const fx_solver_opts& o) {
    return carve(ctx, ld, n_ld, k, o, nullptr, false) + 4096;
}

namespace {

// SVQB pass: dst[:, :b] = src[:, :b] * M with M = D^-1 U L^-1/2 from G = src'src; Cfac out.
int svqb_pass(fx_ctx* ctx, const Work& w, const double* src, double* dst, int64_t ld, double* Mout,
              double* Cout, int variant, FxTimer& tm, fx_stats* st) {
    const int b = w.b;
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_tn(ctx, src, b, src, b, ld, ld, w.G, b, w.tnscr, w.tnscr_doubles, variant));
    tm.end();
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, w.G, (size_t)b * b, ctx->stream));
        tm.end();
    }
    tm.begin(&st->rr_ms);
    // CholQR when the block is well conditioned (pivots of the scaled Gram > 1e-8), else SVQB: the
    // three kernels after the Cholesky return immediately unless it raised the flag (w.info + 6)
    FX_TRY(fx_launch_chol_factors(ctx, w.G, b, b, 1e-8, Mout, Cout, w.info + 6, w.info + 4));
    FX_TRY(fx_launch_svqb_prep(ctx, w.G, b, b, w.Gs, w.dscale, w.info + 6));
    FX_TRY(fx_launch_jacobi(ctx, w.Gs, b, b, w.U, b, w.lamb, w.Uw, w.info + 2, w.info + 6));
    FX_TRY(fx_launch_svqb_factors(ctx, w.U, b, w.lamb, w.dscale, b, 1e-14, Mout, Cout, w.info + 4, w.info + 6));
    tm.end();
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_nn(ctx, src, b, Mout, b, b, 1.0, 0.0, dst, ld, ld, variant));
    tm.end();
    return 0;
}

// W -= Q[:, :mq] (Q[:, :mq]' W);  H (ldh = mq) receives the coefficients when Hout != null
int project_out(fx_ctx* ctx, const Work& w, int mq, double* W, int64_t ld, double* Hout, int variant,
                FxTimer& tm, fx_stats* st) {
    const int b = w.b;
    double* H = Hout ? Hout : w.H2;
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_tn(ctx, w.Q, mq, W, b, ld, ld, H, mq, w.tnscr, w.tnscr_doubles, variant));
    tm.end();
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, H, (size_t)mq * b, ctx->stream));
        tm.end();
    }
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_nn(ctx, w.Q, mq, H, mq, b, -1.0, 1.0, W, ld, ld, variant));
    tm.end();
    return 0;
}

// W -= L (L'W) for the locked eigenvectors (deflation); one CGS pass
int project_lock(fx_ctx* ctx, const Work& w, const double* L, int nlock, double* W, int64_t ld, int variant,
                 FxTimer& tm, fx_stats* st) {
    if (nlock <= 0) return 0;
    const int b = w.b;
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_tn(ctx, L, nlock, W, b, ld, ld, w.HL, nlock, w.tnscr, w.tnscr_doubles, variant));
    tm.end();
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, w.HL, (size_t)nlock * b, ctx->stream));
        tm.end();
    }
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_nn(ctx, L, nlock, w.HL, nlock, b, -1.0, 1.0, W, ld, ld, variant));
    tm.end();
    return 0;
}

// One reduction for a re-projection and the block Gram matrix that follows it:
//   Hs = [Q(:, :mq) W]' W  ((mq + b) x b, one all-reduce), then  G = W'W - H2'H2  is the Gram matrix
// of W - Q H2 (Q orthonormal; H2 is a SECOND-pass coefficient block, tiny, so the subtraction does
// not cancel), the block is orthonormalised from G (CholQR / SVQB) and  dst = (W - Q H2) M.
// Saves one collective and one tall-skinny product per use against project_out + svqb_pass.
int reproject_and_orthonormalise(fx_ctx* ctx, const Work& w, int mq, double* W, double* dst, int64_t ld,
                                 double* Hs, double* Mout, double* Cout, int variant, FxTimer& tm, fx_stats* st) {
    const int b = w.b, ldh = mq + b;
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_tn(ctx, w.Q, mq, W, b, ld, ld, Hs, ldh, w.tnscr, w.tnscr_doubles, variant));
    FX_TRY(fx_launch_ts_tn(ctx, W, b, W, b, ld, ld, Hs + mq, ldh, w.tnscr, w.tnscr_doubles, variant));
    tm.end();
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, Hs, (size_t)ldh * b, ctx->stream));
        tm.end();
    }
    tm.begin(&st->rr_ms);
    FX_TRY(fx_launch_gram_downdate(ctx, Hs, ldh, mq, b, w.G));               // G = Hs[mq:, :] - H2'H2
    FX_TRY(fx_launch_chol_factors(ctx, w.G, b, b, 1e-8, Mout, Cout, w.info + 6, w.info + 4));
    FX_TRY(fx_launch_svqb_prep(ctx, w.G, b, b, w.Gs, w.dscale, w.info + 6));
    FX_TRY(fx_launch_jacobi(ctx, w.Gs, b, b, w.U, b, w.lamb, w.Uw, w.info + 2, w.info + 6));
    FX_TRY(fx_launch_svqb_factors(ctx, w.U, b, w.lamb, w.dscale, b, 1e-14, Mout, Cout, w.info + 4, w.info + 6));
    tm.end();
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_nn(ctx, w.Q, mq, Hs, ldh, b, -1.0, 1.0, W, ld, ld, variant));     // W -= Q H2
    FX_TRY(fx_launch_ts_nn(ctx, W, b, Mout, b, b, 1.0, 0.0, dst, ld, ld, variant));
    tm.end();
    return 0;
}

}  // namespace

int fx_eig_topk(fx_ctx* ctx, const EigArgs& a, FxTimer& tm, fx_stats* st) {
    const fx_solver_opts& o = a.opts;
    const int variant = o.variant;
    const int64_t ld = a.ld;
    const int k = a.k, kp = (int)round_up(k, 8);
    Work w;
    carve(ctx, ld, a.n_ld, k, o, &w, true);
    const int b = w.b, mt = w.mt;
    cudaStream_t s = ctx->stream;
    st->block = b;

    FX_CUDA(cudaMemsetAsync(w.Q, 0, (size_t)mt * ld * 8, s));
    FX_CUDA(cudaMemsetAsync(w.T, 0, (size_t)mt * mt * 8, s));

    // start block: deterministic hash noise (independent of the number of ranks), SVQB twice
    const int nlock = (int)round_up(a.nlock, 8);       // the caller pads the locked block with zero columns
    if (nlock > FX_MAX_LOCK) { fx_set_error("eigensolver: more than %d locked vectors", FX_MAX_LOCK); return -1; }
    FX_TRY(fx_launch_hash_fill(ctx, w.Wa, ld, a.rows_valid, a.row0, b, o.seed + (uint64_t)a.nlock));
    FX_TRY(project_lock(ctx, w, a.lock, nlock, w.Wa, ld, variant, tm, st));
    FX_TRY(project_lock(ctx, w, a.lock, nlock, w.Wa, ld, variant, tm, st));
    FX_TRY(svqb_pass(ctx, w, w.Wa, w.Wb, ld, w.M1, w.C1, variant, tm, st));
    FX_TRY(project_lock(ctx, w, a.lock, nlock, w.Wb, ld, variant, tm, st));
    FX_TRY(svqb_pass(ctx, w, w.Wb, w.Q, ld, w.M2, w.C2, variant, tm, st));

    int m = 0, steps = 0, restarts = 0;
    bool converged = false;
    double maxres = 0.0, res_prev = -1.0, res_prev2 = -1.0;
    double* pinned = ctx->pinned;
    // One K*V pass: gather the block, W = K Qn.  The first pass of a fit with deferred centring
    // also centres K in place (symm.cu CENTER variant) and is booked as the centring pass.
    auto kv_pass = [&](const double* Qn, bool first) -> int {
        const double* Vsrc = Qn;
        if (ctx->world > 1) {
            tm.begin(&st->comm_ms);
            FX_TRY(fxcomm_allgather(ctx->comm, Qn, w.Vg, (size_t)b * ld, s));
            tm.end();
            Vsrc = w.Vg;
        }
        SymmArgs sa;
        sa.K = a.K; sa.ldk = a.ldk; sa.rows_pad = ld; sa.V = Vsrc; sa.ldr = ld;
        sa.v_rank_stride = (int64_t)b * ld; sa.bcols = b; sa.Y = w.Wa; sa.Ypart = w.Ypart;
        sa.splits = w.splits;
        if (first && a.center_colmean) {
            sa.center_colmean = a.center_colmean; sa.center_grand = a.center_grand;
            sa.rows_valid = a.rows_valid; sa.cols_valid = a.n; sa.row0 = a.row0;
            tm.begin(&st->center_ms);
            FX_TRY(fx_launch_symm(ctx, sa, variant));
            tm.end();
            st->center_launches++;
        } else {
            tm.begin(&st->symm_ms);
            FX_TRY(fx_launch_symm(ctx, sa, variant));
            tm.end();
            st->symm_launches++;
        }
        return 0;
    };
    // The Rayleigh-Ritz problem of step s (single-CTA Jacobi, latency bound) runs on the side
    // stream while the main stream already computes K*Qn for step s+1: that product does not
    // depend on the Ritz pairs (only the convergence decision and a thick restart do).  When the
    // residual history says step s probably converges, the solver does not run ahead, so the
    // speculative pass is wasted at most when the prediction fails.
    static const bool overlap_off = []() { const char* e = getenv("FXB200_NO_OVERLAP"); return e && e[0] == '1'; }();
    static const bool lag_off = []() { const char* e = getenv("FXB200_NO_LAG"); return e && e[0] == '1'; }();
    bool pending = false;                               // W = K*Qn for this step is already in Wa
    int wasted = 0, last_read = 0;
    cudaEvent_t ev_rr[2] = {ctx->ev_join, ctx->ev_join2};
    const int hp_off[2] = {0, 1024};                    // pinned slots of even / odd steps

    while (true) {
        double* Qn = w.Q + (size_t)m * ld;
        const int mq = m + b;
        if (!pending) FX_TRY(kv_pass(Qn, steps == 0));
        pending = false;
        steps++;
        // 3-5. CGS2 against [Q Qn]; H accumulates both passes.  The second pass shares its reduction
        // with the Gram matrix of the block orthonormalisation (3 collectives per step, not 5).
        FX_TRY(project_lock(ctx, w, a.lock, nlock, w.Wa, ld, variant, tm, st));
        FX_TRY(project_out(ctx, w, mq, w.Wa, ld, w.H, variant, tm, st));
        FX_TRY(project_lock(ctx, w, a.lock, nlock, w.Wa, ld, variant, tm, st));
        FX_TRY(reproject_and_orthonormalise(ctx, w, mq, w.Wa, w.Wb, ld, w.H2, w.M1, w.C1, variant, tm, st));
        FX_TRY(fx_launch_copy2d_add(ctx, w.H2, mq + b, w.H, mq, mq, b));          // H += H2 (leading mq rows)
        // 6. projected matrix
        FX_TRY(fx_launch_T_update(ctx, w.T, mt, w.H, mq, m, b));
        // 7. re-projection of the normalised block + second orthonormalisation -> next block in place
        FX_TRY(reproject_and_orthonormalise(ctx, w, mq, w.Wb, w.Q + (size_t)mq * ld, ld, w.H2, w.M2, w.C2, variant,
                                            tm, st));
        m = mq;
        // 8. Rayleigh-Ritz on the side stream.  The main stream first waits until the PREVIOUS
        // step's Rayleigh-Ritz has released Tw / C (normally long finished), snapshots T and the
        // residual factor C for this one, and forks.
        const int par = steps & 1;
        double* hp = pinned + hp_off[par];
        if (steps > 1) FX_CUDA(cudaStreamWaitEvent(s, ev_rr[par ^ 1], 0));
        FX_TRY(fx_launch_copy2d(ctx, w.T, mt, w.Tw, mt, m, m));
        FX_TRY(fx_launch_small_matmul(ctx, w.C2, w.C1, b, w.C));
        FX_CUDA(cudaEventRecord(ctx->ev_fork, s));
        FX_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        {
            cudaStream_t keep = ctx->stream;
            ctx->stream = ctx->side;                    // the launchers below use ctx->stream
            tm.begin(&st->rr_ms);
            int rc = fx_launch_jacobi(ctx, w.Tw, mt, m, w.S, mt, w.theta, w.Jw, w.info);
            if (!rc) rc = fx_launch_ritz_residuals(ctx, w.C, b, w.S, mt, m, k, w.res);
            tm.end();
            ctx->stream = keep;
            FX_TRY(rc);
        }
        FX_CUDA(cudaMemcpyAsync(hp, w.res, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->side));
        FX_CUDA(cudaMemcpyAsync(hp + 64, w.theta, sizeof(double), cudaMemcpyDeviceToHost, ctx->side));
        FX_CUDA(cudaMemcpyAsync(hp + 72, w.info, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->side));
        FX_CUDA(cudaEventRecord(ev_rr[par], ctx->side));
        auto read_step = [&](int step) -> int {          // host join on the Rayleigh-Ritz of `step`
            const double* h = pinned + hp_off[step & 1];
            FX_CUDA(cudaEventSynchronize(ev_rr[step & 1]));
            const double scale = a.scale_ref > 0.0 ? a.scale_ref : (fabs(h[64]) > 1e-300 ? fabs(h[64]) : 1e-300);
            double mr = 0.0;
            for (int i = 0; i < k; ++i) {
                if (!(h[i] == h[i])) {
                    fx_set_error("eigensolver: non-finite Ritz residual at step %d", step);
                    return FX_ERR_NOT_CONVERGED;
                }
                if (h[i] > mr) mr = h[i];
            }
            maxres = mr / scale;
            res_prev2 = res_prev;
            res_prev = maxres;
            last_read = step;
            static const bool trace = []() { const char* e = getenv("FXB200_TRACE_SOLVER"); return e && e[0] == '1'; }();
            if (trace && ctx->rank == 0)
                fprintf(stderr, "[fxb200 solver] step %d m=%d nlock=%d max rel residual %.3e%s\n", step,
                        m - (steps - step) * b, a.nlock, maxres, step < steps ? " (read one step late)" : "");
            return 0;
        };
        // The host reads the residuals of the PREVIOUS step here (its Rayleigh-Ritz was queued a
        // whole step ago, so this join does not idle the GPU) and needs THIS step's Ritz pairs now
        // only when the step is expected to converge (residual history), a thick restart follows,
        // or the pass budget is spent.  Otherwise it queues the next step at once: no per-step
        // host join, the Rayleigh-Ritz chain hides behind the main stream.
        converged = false;
        if (last_read < steps - 1) {
            FX_TRY(read_step(steps - 1));
            if (maxres <= o.tol) converged = true;       // a step earlier than predicted
        }
        bool likely_done = converged;
        if (res_prev >= 0.0 && res_prev2 > 0.0) {
            double ratio = res_prev / res_prev2;
            if (ratio > 1.0) ratio = 1.0;
            if (res_prev * ratio <= 100.0 * o.tol) likely_done = true;
        }
        const bool run_ahead = !overlap_off && variant == 0 && steps < o.max_steps && !likely_done;
        const bool need_now = !run_ahead || m + b > w.mcap || lag_off;
        if (run_ahead) {
            FX_TRY(kv_pass(w.Q + (size_t)m * ld, false));
            pending = true;
        }
        if (need_now) {
            // (when the previous step already converged this step's pairs are taken all the same:
            // they span a larger Krylov space, and S / theta of the earlier step are overwritten)
            const bool was = converged;
            FX_TRY(read_step(steps));
            converged = was || maxres <= o.tol;
        }
        if (last_read == steps)
            FX_CUDA(cudaStreamWaitEvent(s, ev_rr[par], 0));    // S, theta are final for the main stream
        if (converged || steps >= o.max_steps) {
            if (pending) wasted = 1;
            break;
        }
        if (m + b > w.mcap) {
            // thick restart: keep the b leading Ritz vectors, T = diag(theta)
            tm.begin(&st->ortho_ms);
            FX_TRY(fx_launch_ts_nn(ctx, w.Q, m, w.S, mt, b, 1.0, 0.0, w.Wb, ld, ld, variant));
            tm.end();
            FX_CUDA(cudaMemcpyAsync(w.Q + (size_t)b * ld, w.Q + (size_t)m * ld, (size_t)b * ld * 8,
                                    cudaMemcpyDeviceToDevice, s));
            FX_CUDA(cudaMemcpyAsync(w.Q, w.Wb, (size_t)b * ld * 8, cudaMemcpyDeviceToDevice, s));
            FX_TRY(fx_launch_T_restart(ctx, w.T, mt, mt, w.theta, b));
            m = b;
            restarts++;
        }
    }
    steps += wasted;                                    // K*V passes actually executed
    st->steps = steps;
    st->restarts = restarts;
    st->converged = converged ? 1 : 0;
    st->max_residual = maxres;

    // Ritz vectors: V_loc = Q[:, :m] S[:, :kp], then gather all rows
    tm.begin(&st->ortho_ms);
    FX_TRY(fx_launch_ts_nn(ctx, w.Q, m, w.S, mt, kp, 1.0, 0.0, w.Wa, ld, ld, variant));
    tm.end();
    if (ctx->world > 1) {
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allgather(ctx->comm, w.Wa, w.Vg, (size_t)kp * ld, s));
        tm.end();
        for (int g = 0; g < ctx->world; ++g)
            FX_CUDA(cudaMemcpy2DAsync(a.Vout + (size_t)g * ld, a.n_ld * 8, w.Vg + (size_t)g * kp * ld,
                                      ld * 8, ld * 8, kp, cudaMemcpyDeviceToDevice, s));
    } else {
        FX_CUDA(cudaMemcpyAsync(a.Vout, w.Wa, (size_t)kp * ld * 8, cudaMemcpyDeviceToDevice, s));
    }
    FX_CUDA(cudaMemcpyAsync(a.theta, w.theta, kp * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (!converged) {
        fx_set_error("eigensolver: not converged after %d K*V passes (max relative residual %.3e, tol %.1e)",
                     steps, maxres, o.tol);
        return FX_ERR_NOT_CONVERGED;
    }
    return 0;
}

// Small-n path: full Jacobi eigendecomposition of the (gathered) matrix, all on the device.
int fx_eig_dense(fx_ctx* ctx, const EigArgs& a, FxTimer& tm, fx_stats* st) {
    const int64_t n = a.n, ld = a.ld, n_ld = a.n_ld;
    const int k = a.k, kp = (int)round_up(k, 8);
    cudaStream_t s = ctx->stream;
    if (n > 1024) {
        // 1024 < n <= 8192, many components: one-sided Jacobi (one GPU; K is destroyed)
        if (ctx->world > 1) { fx_set_error("dense full-spectrum path (k > 1024 or k > n/8) needs a single-GPU context"); return -1; }
        double* Vw = (double*)fx_arena_alloc(ctx, (size_t)n_ld * n * 8);
        double* lw = (double*)fx_arena_alloc(ctx, (size_t)(n + 8) * 8);
        int* rw = (int*)fx_arena_alloc(ctx, (size_t)n * 4);
        unsigned long long* flag = (unsigned long long*)fx_arena_alloc(ctx, 256);
        if (!Vw || !lw || !rw || !flag) { fx_set_error("internal: arena too small for the dense path"); return FX_ERR_OOM; }
        FX_CUDA(cudaMemsetAsync(a.Vout, 0, (size_t)kp * n_ld * 8, s));
        FX_CUDA(cudaMemsetAsync(a.theta, 0, (size_t)kp * 8, s));
        int hi[2] = {0, 0};
        tm.begin(&st->rr_ms);
        FX_TRY(fx_hestenes_eig(ctx, (double*)a.K, a.ldk, (int)n, k, Vw, lw, rw, flag, a.Vout, n_ld, a.theta, hi));
        tm.end();
        st->dense_path = 1; st->steps = hi[0]; st->converged = hi[1]; st->block = 0;
        if (!hi[1]) { fx_set_error("dense one-sided Jacobi did not converge in %d sweeps", hi[0]); return FX_ERR_NOT_CONVERGED; }
        return 0;
    }
    double* Kfull = (double*)a.K;
    if (ctx->world > 1) {
        Kfull = (double*)fx_arena_alloc(ctx, (size_t)n_ld * n_ld * 8);
        tm.begin(&st->comm_ms);
        FX_TRY(fxcomm_allgather(ctx->comm, a.K, Kfull, (size_t)ld * n_ld, s));
        tm.end();
    }
    double* Vw = (double*)fx_arena_alloc(ctx, (size_t)n * n * 8);
    double* S = (double*)fx_arena_alloc(ctx, (size_t)n * n * 8);
    double* th = (double*)fx_arena_alloc(ctx, (size_t)(n + 8) * 8);
    int* info = (int*)fx_arena_alloc(ctx, 64);
    tm.begin(&st->rr_ms);
    FX_TRY(fx_launch_jacobi(ctx, Kfull, (int)a.ldk, (int)n, S, (int)n, th, Vw, info));
    tm.end();
    FX_CUDA(cudaMemsetAsync(a.Vout, 0, (size_t)kp * n_ld * 8, s));
    FX_CUDA(cudaMemcpy2DAsync(a.Vout, n_ld * 8, S, n * 8, n * 8, k, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemsetAsync(a.theta, 0, kp * 8, s));
    FX_CUDA(cudaMemcpyAsync(a.theta, th, k * 8, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(ctx->pinned + 72, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    const int* hi = reinterpret_cast<const int*>(ctx->pinned + 72);
    st->dense_path = 1;
    st->steps = hi[0];
    st->converged = hi[1];
    st->block = 0;
    if (!hi[1]) {
        fx_set_error("dense Jacobi eigensolver did not converge in %d sweeps", hi[0]);
        return FX_ERR_NOT_CONVERGED;
    }
    return 0;
}
