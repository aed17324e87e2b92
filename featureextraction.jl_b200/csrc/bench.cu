// bench.cu -- device-resident timing of individual kernel classes (inputs generated in HBM,
// CUDA events on the context stream) and the fp64 issue-rate microbenchmarks that give the
// roofline denominators for the Gram / SYMM kernels (fp64 tensor peak is not in
// MEASURED_PEAKS.json; BASELINE.md section 2 asks the builder to commit a measurement).
#include <functional>
#include "fx_kernels.cuh"

namespace {

// register-only DMMA chain: 8 independent accumulators per warp, no memory traffic
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// sweep variant: ILP independent accumulators per warp, any block size
template <int ILP>
__global__ void dmma_sweep_kernel(double* out, int iters) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// do DMMA and DFMA share the FP64 datapath?  warps 0-3 run DMMA chains, warps 4-7 DFMA chains;
// mode 0: DMMA only, 1: DFMA only, 2: both concurrently
__global__ void __launch_bounds__(256) fp64_share_kernel(double* out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    double s = 0.0;
    if (warp < 4) {
        if (mode == 1) return;
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
        double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    } else {
        if (mode == 0) return;
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
        const double a = 1.0000001, b = 1e-9;
        for (int it = 0; it < iters * 8; ++it) {           // 8*16 DFMA per DMMA-iteration: equal lane-cycles
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i];
    }
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, long n2) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        const double2 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n2; i += stride) dst[i] = src[i];
}

int time_loop(fx_ctx* ctx, int reps, double* ms_out, const std::function<int()>& fn);

}  // namespace

#include <functional>
namespace {
int time_loop(fx_ctx* ctx, int reps, double* ms_out, const std::function<int()>& fn) {
    cudaEvent_t a, b;
    FX_CUDA(cudaEventCreate(&a));
    FX_CUDA(cudaEventCreate(&b));
    FX_TRY(fn());                                   // warm-up (also sets func attributes)
    FX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < reps; ++r) {
        FX_CUDA(cudaEventRecord(a, ctx->stream));
        FX_TRY(fn());
        FX_CUDA(cudaEventRecord(b, ctx->stream));
        FX_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        FX_CUDA(cudaEventElapsedTime(&ms, a, b));
        ms_out[r] = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return 0;
}
}  // namespace

extern "C" int fx_bench_kernel(fx_ctx* ctx, int which, int64_t n, int64_t d, int b, int64_t m,
                               int kernel_id, int reps, double* ms_out) {
    if (!ctx) { fx_set_error("fx_bench_kernel: ctx is NULL"); return -1; }
    ctx = fx_single(ctx);
    if (reps < 1 || !ms_out) { fx_set_error("fx_bench_kernel: invalid argument 8 (reps < 1 || !ms_out)"); return -8; }
    FX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const double kp_g[3] = {d > 0 ? 1.0 / (double)d : 1.0, 1.0, 3.0};
    const double kp_p[3] = {d > 0 ? 1.0 / (double)d : 1.0, 1.0, 3.0};
    const double kp_s[3] = {d > 0 ? 1.0 / (double)d : 1.0, 0.1, 0.0};
    const double kp_l[3] = {1.0, 0.0, 1.0};          // kernel_id 3 here = linear (polynomial, power 1)
    const double* kp = kernel_id == 0 ? kp_g : (kernel_id == 1 ? kp_p : (kernel_id == 3 ? kp_l : kp_s));
    if (kernel_id == 3) kernel_id = 1;
    if (which == 4 || which == 5) {
        FX_TRY(fx_arena_reserve(ctx, 4096));
        fx_arena_reset(ctx);
        double* out = (double*)fx_arena_alloc(ctx, 256);
        const int iters = 20000;
        const int grid = ctx->sm_count * 8;
        // ms_out[r] = ms; caller converts: DMMA flops = grid*8 warps*iters*8*512 ; DFMA = grid*256*iters*16*2
        return time_loop(ctx, reps, ms_out, [&]() -> int {
            if (which == 4) dmma_peak_kernel<<<grid, 256, 0, s>>>(out, iters);
            else dfma_peak_kernel<<<grid, 256, 0, s>>>(out, iters);
            ctx->launches++;
            FX_LAUNCH_CHECK();
            return 0;
        });
    }
    if (which == 8) {
        FX_TRY(fx_arena_reserve(ctx, 4096));
        fx_arena_reset(ctx);
        double* out = (double*)fx_arena_alloc(ctx, 256);
        const int mode = (int)d, iters = 4000;
        return time_loop(ctx, reps, ms_out, [&]() -> int {
            fp64_share_kernel<<<ctx->sm_count, 256, 0, s>>>(out, iters, mode);
            ctx->launches++;
            FX_LAUNCH_CHECK();
            return 0;
        });
    }
    if (which == 7) {
        // DMMA issue sweep: n = CTAs per SM, b = warps per CTA, d = independent accumulators
        FX_TRY(fx_arena_reserve(ctx, 4096));
        fx_arena_reset(ctx);
        double* out = (double*)fx_arena_alloc(ctx, 256);
        const int grid = ctx->sm_count * (int)n, threads = 32 * b, ilp = (int)d;
        const int iters = (int)m;
        return time_loop(ctx, reps, ms_out, [&]() -> int {
            switch (ilp) {
                case 1: dmma_sweep_kernel<1><<<grid, threads, 0, s>>>(out, iters); break;
                case 2: dmma_sweep_kernel<2><<<grid, threads, 0, s>>>(out, iters); break;
                case 4: dmma_sweep_kernel<4><<<grid, threads, 0, s>>>(out, iters); break;
                case 8: dmma_sweep_kernel<8><<<grid, threads, 0, s>>>(out, iters); break;
                case 16: dmma_sweep_kernel<16><<<grid, threads, 0, s>>>(out, iters); break;
                default: dmma_sweep_kernel<32><<<grid, threads, 0, s>>>(out, iters); break;
            }
            ctx->launches++;
            FX_LAUNCH_CHECK();
            return 0;
        });
    }
    if (which == 6) {
        const size_t bytes = (size_t)n * 8;           // n doubles each way
        FX_TRY(fx_arena_reserve(ctx, 2 * fx_align256(bytes) + 4096));
        fx_arena_reset(ctx);
        double* src = (double*)fx_arena_alloc(ctx, bytes);
        double* dst = (double*)fx_arena_alloc(ctx, bytes);
        FX_CUDA(cudaMemsetAsync(src, 0, bytes, s));
        return time_loop(ctx, reps, ms_out, [&]() -> int {
            copy_kernel<<<ctx->sm_count * 16, 512, 0, s>>>((const double2*)src, (double2*)dst, (long)(n / 2));
            ctx->launches++;
            FX_LAUNCH_CHECK();
            return 0;
        });
    }
    if (n < 128 || n % 128) { fx_set_error("bench: n must be a positive multiple of 128"); return -3; }
    const int64_t d_pad = round_up(d > 0 ? d : 16, 16);
    if (which == 0 || which == 1) {
        size_t need = fx_align256((size_t)d_pad * n * 8) + fx_align256((size_t)n * 8) * 3 +
                      fx_align256((size_t)n * n * 8) + fx_align256((size_t)(n / 64) * n * 8) + 8192;
        FX_TRY(fx_arena_reserve(ctx, need));
        fx_arena_reset(ctx);
        double* Xp = (double*)fx_arena_alloc(ctx, (size_t)d_pad * n * 8);
        double* nr = (double*)fx_arena_alloc(ctx, (size_t)n * 8);
        double* cm = (double*)fx_arena_alloc(ctx, (size_t)n * 8);
        double* sc = (double*)fx_arena_alloc(ctx, 256);
        double* K = (double*)fx_arena_alloc(ctx, (size_t)n * n * 8);
        double* rp = (double*)fx_arena_alloc(ctx, (size_t)(n / 64) * n * 8);
        FX_TRY(fx_launch_hash_fill(ctx, Xp, n, n, 0, (int)d_pad, 77));
        FX_TRY(fx_launch_rownorms(ctx, Xp, n, n, d_pad, nr));
        GramArgs ga;
        ga.A = Xp; ga.lda = n; ga.B = Xp; ga.ldb = n; ga.na = nr; ga.nb = nr; ga.d_pad = d_pad; ga.K = K;
        ga.ldk = n; ga.rows_pad = n; ga.rows_valid = n; ga.cols_valid = n; ga.diag_shift = 0; ga.rowpart = rp;
        ga.sym_source = (b == 1) ? 0 : 1;           // b = 1: time the all-tiles kernel
        ga.kernel_id = kernel_id; ga.kp[0] = kp[0]; ga.kp[1] = kp[1]; ga.kp[2] = kp[2];
        if (which == 0) return time_loop(ctx, reps, ms_out, [&]() -> int { return fx_launch_gram(ctx, ga, 0); });
        FX_TRY(fx_launch_gram(ctx, ga, 0));
        FX_CUDA(cudaMemsetAsync(cm, 0, (size_t)n * 8, s));
        FX_CUDA(cudaMemsetAsync(sc, 0, 256, s));
        return time_loop(ctx, reps, ms_out, [&]() -> int { return fx_launch_center(ctx, K, n, n, n, 0, cm, sc); });
    }
    if (which == 2) {
        const int bp = (int)round_up(b, 8);
        const int splits = fx_symm_splits(ctx, n, n);
        size_t need = fx_align256((size_t)n * n * 8) + 2 * fx_align256((size_t)bp * n * 8) +
                      fx_align256((size_t)splits * bp * n * 8) + 8192;
        FX_TRY(fx_arena_reserve(ctx, need));
        fx_arena_reset(ctx);
        double* K = (double*)fx_arena_alloc(ctx, (size_t)n * n * 8);
        double* V = (double*)fx_arena_alloc(ctx, (size_t)bp * n * 8);
        double* Y = (double*)fx_arena_alloc(ctx, (size_t)bp * n * 8);
        double* Yp = (double*)fx_arena_alloc(ctx, (size_t)splits * bp * n * 8);
        // K filled in column strips of 64 to bound the grid's y dimension
        for (int64_t c0 = 0; c0 < n; c0 += 32768) {
            const int cols = (int)((n - c0) < 32768 ? (n - c0) : 32768);
            FX_TRY(fx_launch_hash_fill(ctx, K + c0 * n, n, n, c0, cols, 78));
        }
        FX_TRY(fx_launch_hash_fill(ctx, V, n, n, 0, bp, 79));
        SymmArgs sa;
        sa.K = K; sa.ldk = n; sa.rows_pad = n; sa.V = V; sa.ldr = n; sa.v_rank_stride = (int64_t)bp * n;
        sa.bcols = bp; sa.Y = Y; sa.Ypart = Yp; sa.splits = splits;
        return time_loop(ctx, reps, ms_out, [&]() -> int { return fx_launch_symm(ctx, sa, 0); });
    }
    if (which == 3) {
        const int kc = (int)round_up(b, 8);
        const int64_t mp = round_up(m, 128);
        int nsplit = 1;
        const size_t scr = fx_transform_scratch(ctx, mp, n, kc, &nsplit);
        size_t need = fx_align256((size_t)d_pad * n * 8) + fx_align256((size_t)d_pad * mp * 8) +
                      fx_align256((size_t)n * 8) + fx_align256((size_t)mp * 8) + fx_align256((size_t)kc * n * 8) +
                      fx_align256((size_t)kc * mp * 8) + fx_align256(scr * 8) + 8192;
        FX_TRY(fx_arena_reserve(ctx, need));
        fx_arena_reset(ctx);
        double* Xp = (double*)fx_arena_alloc(ctx, (size_t)d_pad * n * 8);
        double* Zp = (double*)fx_arena_alloc(ctx, (size_t)d_pad * mp * 8);
        double* nx = (double*)fx_arena_alloc(ctx, (size_t)n * 8);
        double* nz = (double*)fx_arena_alloc(ctx, (size_t)mp * 8);
        double* W = (double*)fx_arena_alloc(ctx, (size_t)kc * n * 8);
        double* O = (double*)fx_arena_alloc(ctx, (size_t)kc * mp * 8);
        double* scratch = scr ? (double*)fx_arena_alloc(ctx, scr * 8) : nullptr;
        FX_TRY(fx_launch_hash_fill(ctx, Xp, n, n, 0, (int)d_pad, 80));
        FX_TRY(fx_launch_hash_fill(ctx, Zp, mp, mp, 1 << 20, (int)d_pad, 81));
        FX_TRY(fx_launch_hash_fill(ctx, W, n, n, 0, kc, 82));
        FX_TRY(fx_launch_rownorms(ctx, Xp, n, n, d_pad, nx));
        FX_TRY(fx_launch_rownorms(ctx, Zp, mp, mp, d_pad, nz));
        TransformArgs ta;
        ta.Zp = Zp; ta.ldz = mp; ta.nz = nz; ta.Xp = Xp; ta.ldx = n; ta.nx = nx; ta.d_pad = d_pad;
        ta.m_pad = mp; ta.n_ld = n; ta.n_valid = n; ta.W = W; ta.ldw = n; ta.kcols = kc; ta.bias = nullptr;
        ta.out = O; ta.ldo = mp; ta.kernel_id = kernel_id; ta.kp[0] = kp[0]; ta.kp[1] = kp[1]; ta.kp[2] = kp[2];
        return time_loop(ctx, reps, ms_out, [&]() -> int { return fx_launch_transform_ex(ctx, ta, 0, scratch, nsplit); });
    }
    fx_set_error("bench: unknown kernel class %d", which);
    return -2;
}
