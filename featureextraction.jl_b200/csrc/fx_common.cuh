// fx_common.cuh -- shared host/device helpers for the fxb200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/fxb200.h"

// ------------------------------------------------------------------------------------------
// error plumbing: every failure sets the thread-local message and returns an FX_ERR_* class
// ------------------------------------------------------------------------------------------
void fx_set_error(const char* fmt, ...);

#define FX_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            fx_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,      \
                         __LINE__, #expr);                                                     \
            return (_e == cudaErrorMemoryAllocation) ? FX_ERR_OOM : FX_ERR_CUDA;               \
        }                                                                                      \
    } while (0)

#define FX_TRY(expr)                                                                           \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) return _r;                                                                \
    } while (0)

#define FX_LAUNCH_CHECK() FX_CUDA(cudaGetLastError())

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (comm.cpp): no link-time dependency, Julia / Python never see NCCL
// ------------------------------------------------------------------------------------------
struct FxComm;
int  fxcomm_unique_id(void* id128);
int  fxcomm_create(int rank, int world, const void* id128, FxComm** out);
void fxcomm_destroy(FxComm* c);
int  fxcomm_allreduce_sum(FxComm* c, double* buf, size_t count, cudaStream_t s);      // in place
int  fxcomm_allgather(FxComm* c, const double* send, double* recv, size_t count_per_rank,
                      cudaStream_t s);

// ------------------------------------------------------------------------------------------
// context: one GPU, one stream, a grow-only device arena reused across calls
// ------------------------------------------------------------------------------------------
struct fx_ctx {
    int device = 0, rank = 0, world = 1, sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    FxComm* comm = nullptr;
    // arena
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // pinned scratch for small device->host reads inside solver loops
    double* pinned = nullptr;
    size_t pinned_cap = 0;
    // timing events (pool)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    int64_t launches = 0;
};

// Arena: reserve() guarantees capacity (may re-allocate: only legal when nothing is live),
// alloc() bumps.  All sizes rounded to 256 B.
int   fx_arena_reserve(fx_ctx* ctx, size_t bytes);
void* fx_arena_alloc(fx_ctx* ctx, size_t bytes);
static inline void fx_arena_reset(fx_ctx* ctx) { ctx->arena_off = 0; }
static inline size_t fx_align256(size_t b) { return (b + 255) & ~size_t(255); }

// Event-pair timer accumulating into a double (resolved at the end of the call).
struct FxTimer {
    fx_ctx* ctx;
    struct Span { cudaEvent_t a, b; double* acc; };
    std::vector<Span> spans;
    explicit FxTimer(fx_ctx* c) : ctx(c) {}
    ~FxTimer() { ctx->ev_used = 0; }
    cudaEvent_t get_event();
    std::vector<int> open_;       // stack of open spans (begin/end nest)
    void begin(double* acc);      // records start on ctx->stream
    void end();                   // records end for the innermost open span
    int  resolve();               // after stream sync: adds elapsed ms to every accumulator
};

// ------------------------------------------------------------------------------------------
// fitted model (device resident, replicated on every rank)
// ------------------------------------------------------------------------------------------
struct fx_model {
    int device = 0;
    int64_t n = 0, d = 0, n_ld = 0, d_pad = 0;
    int k = 0, k_pad = 0;
    int kernel_id = 0;
    double kp[3] = {0, 0, 0};
    int centered = 0;
    int is_klda = 0;
    double* Xp = nullptr;      // [d_pad][n_ld] zero padded, feature-major (Julia column-major X)
    double* norms = nullptr;   // [n_ld] squared row norms
    double* W = nullptr;       // [k_pad][n_ld] components (column-major n_ld x k_pad)
    double* bias = nullptr;    // [k_pad]  r'W  (zero when not centred)
    double* colmean = nullptr; // [n_ld]
    double grandmean = 0.0;
};

// ------------------------------------------------------------------------------------------
// device-side PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16 B aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes),
        "r"(smem_u32(bar))
        : "memory");
}

// fp64 tensor-core MMA, 8x8x4 (SASS DMMA.8x8x4): D = A(8x4,row) * B(4x8,col) + C.
// Lane l: g = l>>2, t = l&3.  a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Branch-free exp(x) for x <= 0 (the Gaussian kernel's argument): Cody-Waite reduction by
// ln2 (hi/lo), degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18), scaling by two
// exact powers of two so results below the normal range flush smoothly to 0.  No slow path, so
// unrolled callers interleave many independent chains on the FP64 pipe.
__device__ __forceinline__ double fx_exp_nonpos(double x) {
    x = fmax(x, -746.0);
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);   // round-to-nearest via magic
    const int n = __double2loint(t);
    const double fn = t - 6755399441055744.0;
    double r = fma(fn, -6.93147180369123816490e-01, x);
    r = fma(fn, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                 // 1/13!
    p = fma(p, r, 2.08767569878680990e-09);            // 1/12!
    p = fma(p, r, 2.50521083854417188e-08);            // 1/11!
    p = fma(p, r, 2.75573192239858907e-07);            // 1/10!
    p = fma(p, r, 2.75573192239858907e-06);            // 1/9!
    p = fma(p, r, 2.48015873015873016e-05);            // 1/8!
    p = fma(p, r, 1.98412698412698413e-04);            // 1/7!
    p = fma(p, r, 1.38888888888888889e-03);            // 1/6!
    p = fma(p, r, 8.33333333333333333e-03);            // 1/5!
    p = fma(p, r, 4.16666666666666667e-02);            // 1/4!
    p = fma(p, r, 1.66666666666666667e-01);            // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int n1 = n >> 1, n2 = n - n1;                // n in [-1077, 0]: both exponents normal
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
    return (p * s1) * s2;
}

// kernel map of a dot product (and squared norms for the Gaussian), see include/fxb200.h
struct KernelMap {
    int id;
    double p0, p1, p2;
    int ipow;      // integer exponent when p2 is a small positive integer, else 0
};
// compile-time kernel selector: 0 Gaussian, 1 polynomial with integer power, 2 sigmoid,
// 3 polynomial with a general real power
template <int KID>
__device__ __forceinline__ double kernel_map_t(const KernelMap& km, double dot, double ni, double nj,
                                               bool same_point) {
    if (KID == 0) {
        double d2 = fmax(ni + nj - 2.0 * dot, 0.0);
        d2 = same_point ? 0.0 : d2;
        return fx_exp_nonpos(-km.p0 * d2);
    } else if (KID == 1) {
        const double base = fma(km.p0, dot, km.p1);
        double r = base;
        for (int e = 1; e < km.ipow; ++e) r *= base;
        return r;
    } else if (KID == 2) {
        return tanh(fma(km.p0, dot, km.p1));
    } else {
        return pow(fma(km.p0, dot, km.p1), km.p2);
    }
}
static __host__ __device__ __forceinline__ int kernel_variant(int id, int ipow) {
    return id == FX_KERNEL_GAUSSIAN ? 0 : (id == FX_KERNEL_SIGMOID ? 2 : (ipow > 0 ? 1 : 3));
}
__device__ __forceinline__ double kernel_map(const KernelMap& km, double dot, double ni,
                                             double nj, bool same_point) {
    switch (kernel_variant(km.id, km.ipow)) {
        case 0: return kernel_map_t<0>(km, dot, ni, nj, same_point);
        case 1: return kernel_map_t<1>(km, dot, ni, nj, same_point);
        case 2: return kernel_map_t<2>(km, dot, ni, nj, same_point);
        default: return kernel_map_t<3>(km, dot, ni, nj, same_point);
    }
}
#endif  // __CUDACC__

static inline int fx_make_kernel_map(int kernel_id, const double* kp, int* ipow_out,
                                     double out[3]) {
    if (kernel_id < 0 || kernel_id > 2 || !kp) return -1;
    out[0] = kp[0]; out[1] = kp[1]; out[2] = kp[2];
    *ipow_out = 0;
    if (kernel_id == FX_KERNEL_GAUSSIAN && !(kp[0] > 0.0)) return -1;
    if (kernel_id == FX_KERNEL_POLYNOMIAL) {
        double p = kp[2];
        if (p >= 1.0 && p <= 16.0 && p == (double)(int)p) *ipow_out = (int)p;
    }
    return 0;
}
