// fx_common.cuh -- shared host/device helpers for the fxb200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <unordered_set>
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
// single process, one communicator per device (ncclGroupStart / ncclCommInitRank / ncclGroupEnd)
int  fxcomm_create_all(const int* devices, int ndev, FxComm** out);
void fxcomm_destroy(FxComm* c);
int  fxcomm_allreduce_sum(FxComm* c, double* buf, size_t count, cudaStream_t s);      // in place
int  fxcomm_allgather(FxComm* c, const double* send, double* recv, size_t count_per_rank,
                      cudaStream_t s);

// ------------------------------------------------------------------------------------------
// context: one GPU, one stream, a grow-only device arena reused across calls
// ------------------------------------------------------------------------------------------
// Host-side rendezvous of the rank threads of a single-process multi-GPU context (fx_init_multi).
struct FxMultiShared {
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0, generation = 0, n = 0;
    char* arena[16] = {};
    size_t off[16][2] = {};
    int ok[16] = {};
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};

struct fx_ctx {
    int device = 0, rank = 0, world = 1, sm_count = 148;
    // Single-process multi-GPU: the PARENT context (fx_init_multi) owns one sub-context per device and
    // no device state of its own; every entry point fans out to one host thread per sub-context.
    std::vector<fx_ctx*> sub;
    fx_ctx* parent = nullptr;
    FxMultiShared* ms = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // side stream for the small Rayleigh-Ritz problem that runs beside the next K*V pass
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    FxComm* comm = nullptr;
    // arena
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // Peer mapping of the other ranks' arenas (CUDA IPC, one process per GPU on one node): the
    // symmetric Gram kernel stores mirrored tiles straight into the owning rank's K over NVLink.
    // Filled by fx_peer_exchange (collective); peer_valid says every peer arena is mapped.
    static constexpr int MAX_PEERS = 16;
    char* peer_arena[MAX_PEERS] = {};
    unsigned char peer_handle[MAX_PEERS][64] = {};
    double peer_gen[MAX_PEERS] = {};
    size_t peer_off[MAX_PEERS][2] = {};          // per rank: arena offsets announced in the last exchange
    bool peer_valid = false, peer_failed = false, arena_exported = false;
    double arena_gen = 0.0;                      // bumped on every arena re-allocation
    std::vector<char*> retired;                  // exported arenas replaced by a bigger one: freed once
                                                 // every peer has dropped its mapping
    double* peer_stage = nullptr;                // device staging for the handle all-gather
    // pinned scratch for small device->host reads inside solver loops
    double* pinned = nullptr;
    size_t pinned_cap = 0;
    // timing events (pool)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    int64_t launches = 0;
    // kernels whose opt-in shared-memory attribute has been set on THIS device (the attribute is
    // per device / context, so the registry lives here and not in a process-wide static)
    std::unordered_set<const void*> attr_done;
};
// Opt a kernel into `bytes` of dynamic shared memory once per context (recorded only on success).
// Pass template instantiations in parentheses: FX_SET_SMEM_ONCE(ctx, (kernel<1, 2>), bytes).
#define FX_SET_SMEM_ONCE(ctx, fn, bytes)                                                               \
    do {                                                                                               \
        if (!(ctx)->attr_done.count((const void*)(fn))) {                                              \
            FX_CUDA(cudaFuncSetAttribute((fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            (ctx)->attr_done.insert((const void*)(fn));                                                \
        }                                                                                              \
    } while (0)
constexpr int FX_MAX_DYN_SMEM = 227 * 1024;

// Arena: reserve() guarantees capacity (may re-allocate: only legal when nothing is live),
// alloc() bumps.  All sizes rounded to 256 B.
int   fx_arena_reserve(fx_ctx* ctx, size_t bytes);
void* fx_arena_alloc(fx_ctx* ctx, size_t bytes);
static inline void fx_arena_reset(fx_ctx* ctx) { ctx->arena_off = 0; }
// Collective over all ranks (no-op for world == 1): announces this rank's arena (IPC handle) and
// the two arena offsets the caller wants peers to reach (K block, row-sum partials), maps every
// peer's arena, and sets ctx->peer_valid.  Also a stream-ordered barrier across ranks.
// Falls back to peer_valid == false (identically on every rank) if IPC is unavailable or
// FXB200_NO_PEER=1.
int fx_peer_exchange(fx_ctx* ctx, size_t off0, size_t off1);
// Collective (no-op for one rank): every rank passes its local status before a section full of
// collectives; if ANY rank failed (e.g. its workspace allocation), all ranks return an error instead
// of some of them blocking in NCCL forever.  Returns rc if non-zero, else FX_ERR_CUDA when a peer failed.
int fx_agree(fx_ctx* ctx, int rc);
void fx_peer_close_all(fx_ctx* ctx);
static inline size_t fx_align256(size_t b) { return (b + 255) & ~size_t(255); }

// Event-pair timer accumulating into a double (resolved at the end of the call).
struct FxTimer {
    fx_ctx* ctx;
    struct Span { cudaEvent_t a, b; double* acc; cudaStream_t s; };
    std::vector<Span> spans;
    explicit FxTimer(fx_ctx* c) : ctx(c) {}
    ~FxTimer() { ctx->ev_used = 0; }
    cudaEvent_t get_event();
    std::vector<int> open_;       // stack of open spans (begin/end nest)
    void begin(double* acc, cudaStream_t on = nullptr);   // records start on `on` (default ctx->stream)
    void end();                   // records end for the innermost open span, on the same stream
    int  resolve();               // after stream sync: adds elapsed ms to every accumulator
};

// ------------------------------------------------------------------------------------------
// fitted model (device resident, replicated on every rank)
// ------------------------------------------------------------------------------------------
struct fx_model {
    std::vector<fx_model*> sub;   // model of a multi-GPU parent context: one replica per device
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
    double* wsum = nullptr;    // [k_pad]  1'W (sum over the n training rows of every component)
    double grandmean = 0.0;
    // 1 when some component is not orthogonal to the constant vector (|1'w| above rounding level):
    // the transform then applies the full double-centring formula with the per-query row mean
    // (SURVEY.md A.2); for eigenvectors of the centred matrix with non-zero eigenvalue 1'w = 0 and
    // the k-vector bias alone is exact.
    int need_rowmean = 0;
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

// Branch-free exp(x) for x <= 0 (the Gaussian kernel's argument), 13 FP64-pipe operations:
// x = (64 m + j) ln2/64 + r with |r| <= ln2/128 (Cody-Waite, hi/lo), exp(x) = 2^m * T[j] * e^r,
// T[j] = 2^(j/64) as a hi+lo pair (one 16-byte load), e^r - 1 by a degree-5 polynomial
// (truncation 3.5e-17).  Max error 0.83 ulp against 80-bit expl() over [-708, 0] (measured on the host
// build of this very function, tests/test_device_math_host.py); exact 1 at x = 0; results
// x < -708 are clamped (result exp(-708) = 3.3e-308 instead of a subnormal / 0: absolute error < 3.4e-308).
// Small positive x (cancellation in ||x||^2+||y||^2-2x'y) are fine.  DMMA and DFMA share one FP64
// datapath on B200, so every operation saved here is DMMA issue time.
#define FX_EXP_TAB_VALUES 1.0, 0.0, 1.0108892860517005, -1.5234778603368577e-17, 1.0218971486541166, 5.109225028973444e-17, 1.0330248790212284, 7.600838874027088e-18, 1.0442737824274138, 8.551889705537965e-17, 1.0556451783605572, 1.759325738772092e-18, 1.0671404006768237, -7.899853966841582e-17, 1.0787607977571199, -6.656660436056593e-17, 1.0905077326652577, -3.046782079812471e-17, 1.102382583307841, 5.2660368715706944e-17, 1.1143867425958924, 1.0410278456845571e-16, 1.1265216186082418, 5.165856758795457e-17, 1.1387886347566916, 8.912812676025408e-17, 1.1511892299529827, 3.250710218863827e-17, 1.1637248587775775, 3.8292048369240935e-17, 1.1763969916502812, 5.554203254218079e-17, 1.189207115002721, 3.982015231465646e-17, 1.202156731452703, 6.644981499252301e-17, 1.215247359980469, -7.712630692681488e-17, 1.22848053610687, -1.89878163130253e-17, 1.241857812073484, 4.658027591836937e-17, 1.255380757024691, -6.7113898212968784e-18, 1.2690509571917332, 2.667932131342186e-18, 1.2828700160787783, 1.713594918243561e-17, 1.2968395546510096, 2.5382502794888315e-17, 1.3109612115247644, -7.181536135519454e-17, 1.3252366431597413, -2.8587312100388614e-17, 1.339667524053303, 8.927282594831732e-17, 1.3542555469368927, 7.70094837980299e-17, 1.3690024229745905, 9.593797919118849e-17, 1.383909881963832, -6.770511658794786e-17, 1.3989796725383112, -9.614213209051323e-17, 1.4142135623730951, -9.667293313452913e-17, 1.42961333839197, -1.2031642489053655e-17, 1.4451808069770467, -3.0237581349939873e-17, 1.460917794180647, -5.600377186075216e-17, 1.4768261459394993, -3.483994556892796e-17, 1.4929077282912648, 1.4192920154284036e-17, 1.5091644275934228, -1.016455327754295e-16, 1.5255981507445384, -1.1024941712342561e-16, 1.5422108254079407, 7.949834809697621e-17, 1.559004400237837, 3.7812070533575275e-17, 1.5759808451078865, -1.0136916471278304e-17, 1.593142151342267, -1.0094406542311964e-16, 1.6104903319492543, 2.4707192569797888e-17, 1.6280274218573478, -6.712955084707084e-17, 1.645755478153965, -1.0125679913674773e-16, 1.6636765803267364, 5.8909926967131e-17, 1.681792830507429, 8.199010020581497e-17, 1.7001063537185235, -8.0237193703977e-18, 1.718619298122478, -1.851380418263111e-17, 1.7373338352737062, 3.164389299292957e-17, 1.7562521603732995, 2.960140695448873e-17, 1.7753764925265212, 6.429731796556572e-17, 1.7947090750031072, 1.8227458427912087e-17, 1.8142521755003989, -9.969531538920349e-17, 1.8340080864093424, 3.283107224245627e-17, 1.8539791250833855, 9.761887490727594e-17, 1.8741676341103, -6.122763413004143e-17, 1.8945759815869656, 3.4034035352165297e-17, 1.9152065613971474, -1.0619946056195963e-16, 1.9360617934922943, 1.0332385960676326e-16, 1.9571441241754002, 8.960767791036668e-17, 1.978456026387951, 4.0388753109278167e-17
static __device__ const double fx_exp_tab[128] = {FX_EXP_TAB_VALUES};
// Host mirror of the table and of the three bit-level intrinsics so that fx_exp_nonpos also
// compiles for the host (tests/test_device_math_host.py measures its error against expl()).
#ifndef __CUDA_ARCH__
static const double fx_exp_tab_host[128] = {FX_EXP_TAB_VALUES};
static inline int fx_host_lo(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(uint32_t)u; }
static inline int fx_host_hi(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(uint32_t)(u >> 32); }
static inline double fx_host_hilo(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v; memcpy(&v, &u, 8); return v;
}
#endif
__host__ __device__ __forceinline__ double fx_exp_nonpos(double x) {
    // clamp at -708 (below, the 2^m scale would leave the normal range).  For x <= 0 the high word,
    // read as unsigned, grows with |x|, so one integer min on it does the job (-708.0 is
    // 0xC0862000'00000000; the low word is kept, i.e. the clamp lands in [-708.0007, -708]);
    // small positive x (high word < 0x80000000) pass unchanged.  fmax() costs five instructions.
#ifdef __CUDA_ARCH__
    x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
#else
    {
        const unsigned h = (unsigned)fx_host_hi(x);
        x = fx_host_hilo((int)(h < 0xC0862000u ? h : 0xC0862000u), fx_host_lo(x));
    }
#endif
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);   // round-to-nearest via magic
#ifdef __CUDA_ARCH__
    const int n = __double2loint(t);
#else
    const int n = fx_host_lo(t);
#endif
    const double fn = t - 6755399441055744.0;
    double r = fma(fn, -0.010830424696905538, x);
    r = fma(fn, 6.563929801064195e-13, r);
    double p = 8.33333333333333333e-03;                // 1/5!
    p = fma(p, r, 4.16666666666666667e-02);
    p = fma(p, r, 1.66666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    const double p1 = p * r;                           // e^r - 1
#ifdef __CUDA_ARCH__
    const double2 th = __ldg(reinterpret_cast<const double2*>(fx_exp_tab) + (n & 63));
#else
    const double2 th = make_double2(fx_exp_tab_host[2 * (n & 63)], fx_exp_tab_host[2 * (n & 63) + 1]);
#endif
    const double v = th.x + fma(th.x, p1, th.y);       // in (0.99, 2): normal
    // scale by 2^(n >> 6) in the exponent field (n >> 6 >= -1022: the result stays normal)
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
#else
    return fx_host_hilo(fx_host_hi(v) + ((n >> 6) << 20), fx_host_lo(v));
#endif
}

// The same function on G independent arguments at once, written stage by stage so that the G
// dependency chains are interleaved in program order (ptxas otherwise emits them one after the
// other -- five dependent DFMAs back to back -- and the map runs at an IPC of 0.3).  Element for
// element the operations and their order are those of fx_exp_nonpos: bit-identical results.
#ifdef __CUDACC__
template <int G>
__device__ __forceinline__ void fx_exp_nonpos_vec(double (&x)[G]) {
    int n[G];
    double r[G], p[G];
#pragma unroll
    for (int i = 0; i < G; ++i)
        x[i] = __hiloint2double((int)min((unsigned)__double2hiint(x[i]), 0xC0862000u), __double2loint(x[i]));
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const double t = fma(x[i], 92.33248261689366, 6755399441055744.0);
        n[i] = __double2loint(t);
        p[i] = t - 6755399441055744.0;                     // fn
    }
#pragma unroll
    for (int i = 0; i < G; ++i) r[i] = fma(p[i], -0.010830424696905538, x[i]);
#pragma unroll
    for (int i = 0; i < G; ++i) r[i] = fma(p[i], 6.563929801064195e-13, r[i]);
#pragma unroll
    for (int i = 0; i < G; ++i) p[i] = fma(8.33333333333333333e-03, r[i], 4.16666666666666667e-02);
#pragma unroll
    for (int i = 0; i < G; ++i) p[i] = fma(p[i], r[i], 1.66666666666666667e-01);
#pragma unroll
    for (int i = 0; i < G; ++i) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < G; ++i) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < G; ++i) p[i] = p[i] * r[i];       // e^r - 1
    double2 th[G];
#pragma unroll
    for (int i = 0; i < G; ++i) th[i] = __ldg(reinterpret_cast<const double2*>(fx_exp_tab) + (n[i] & 63));
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const double v = th[i].x + fma(th[i].x, p[i], th[i].y);
        x[i] = __hiloint2double(__double2hiint(v) + ((n[i] >> 6) << 20), __double2loint(v));
    }
}
#endif

// kernel map of a dot product (and squared norms for the Gaussian), see include/fxb200.h
struct KernelMap {
    int id;
    double p0, p1, p2;
    int ipow;      // integer exponent when p2 is a small positive integer, else 0
};
// compile-time kernel selector: 0 Gaussian, 1 polynomial with an integer power <= 16
// (branch-free square-and-multiply), 2 sigmoid, 3 polynomial with a general real power,
// 4 / 5 / 6 polynomial with power 1 / 2 / 3
template <int KID>
__device__ __forceinline__ double kernel_map_t(const KernelMap& km, double dot, double ni, double nj,
                                               bool same_point) {
    if (KID == 0) {
        // no clamp at 0: a rounding-negative squared distance gives exp(+1e-16) = 1 + 1e-16
        double d2 = ni + nj - 2.0 * dot;
        d2 = same_point ? 0.0 : d2;
        return fx_exp_nonpos(-km.p0 * d2);
    } else if (KID == 1) {
        double bb = fma(km.p0, dot, km.p1), r = 1.0;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            r = ((km.ipow >> bit) & 1) ? r * bb : r;
            bb *= bb;
        }
        return r;
    } else if (KID == 2) {
        return tanh(fma(km.p0, dot, km.p1));
    } else if (KID == 3) {
        return pow(fma(km.p0, dot, km.p1), km.p2);
    } else {
        const double base = fma(km.p0, dot, km.p1);
        return KID == 4 ? base : (KID == 5 ? base * base : (base * base) * base);
    }
}
// Gaussian map with the norms already multiplied by -gamma (gram.cu scales them once per tile):
// exp(-gamma (ni + nj - 2 dot)) = exp(fma(2 gamma, dot, ni' + nj')), two FP64 operations before
// the exp instead of three.
__device__ __forceinline__ double gaussian_map_prescaled(double two_gamma, double dot, double nis, double njs,
                                                         bool same_point) {
    double x = fma(two_gamma, dot, nis + njs);
    x = same_point ? 0.0 : x;
    return fx_exp_nonpos(x);
}
static __host__ __device__ __forceinline__ int kernel_variant(int id, int ipow) {
    if (id == FX_KERNEL_GAUSSIAN) return 0;
    if (id == FX_KERNEL_SIGMOID) return 2;
    if (ipow >= 1 && ipow <= 3) return 3 + ipow;
    return ipow > 0 ? 1 : 3;
}
__device__ __forceinline__ double kernel_map(const KernelMap& km, double dot, double ni,
                                             double nj, bool same_point) {
    switch (kernel_variant(km.id, km.ipow)) {
        case 0: return kernel_map_t<0>(km, dot, ni, nj, same_point);
        case 1: return kernel_map_t<1>(km, dot, ni, nj, same_point);
        case 2: return kernel_map_t<2>(km, dot, ni, nj, same_point);
        case 3: return kernel_map_t<3>(km, dot, ni, nj, same_point);
        case 4: return kernel_map_t<4>(km, dot, ni, nj, same_point);
        case 5: return kernel_map_t<5>(km, dot, ni, nj, same_point);
        default: return kernel_map_t<6>(km, dot, ni, nj, same_point);
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
