// api_core.cu -- contexts, error text, device arena, event timers.
#include <stdarg.h>
#include <stdlib.h>
#include <thread>
#include "fx_kernels.cuh"

static thread_local char g_err[1024] = "";

void fx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fx_last_error(void) { return g_err; }
extern "C" int fx_version(void) { return 100; }

int fx_arena_reserve(fx_ctx* ctx, size_t bytes) {
    bytes = fx_align256(bytes) + 4096;
    if (bytes <= ctx->arena_cap) return 0;
    if (ctx->arena) {
        FX_CUDA(cudaStreamSynchronize(ctx->stream));
        // an arena whose IPC handle went out to peers may still be mapped there: park it until
        // the next collective exchange has made every peer drop the mapping
        if (ctx->arena_exported) ctx->retired.push_back(ctx->arena);
        else FX_CUDA(cudaFree(ctx->arena));
        ctx->arena_exported = false;
        ctx->arena = nullptr;
        ctx->arena_cap = 0;
    }
    cudaError_t e = cudaMalloc((void**)&ctx->arena, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fx_set_error("out of device memory: workspace of %.2f GB refused (%s)", bytes / 1e9,
                     cudaGetErrorString(e));
        return FX_ERR_OOM;
    }
    ctx->arena_cap = bytes;
    ctx->arena_off = 0;
    ctx->arena_gen += 1.0;
    return 0;
}

// Cross-rank symmetric Gram (every off-diagonal super-tile computed once, its transpose stored into the
// owner's K over NVLink from inside the kernel): OPT-IN since round 2 (FXB200_PEER_SYM=1).  With the
// remote stores in flight a fit occasionally comes out with one 64 x 8 sub-block of K off (about 1 fit in
// 50 on four GPUs, every second fit on eight; bit-identical to the one-GPU result otherwise): eigenvalues
// then differ in the 7th digit, far outside the 1e-10 parity bar.  The rank-local schedule (each rank
// computes its whole row block, symmetric only inside its diagonal block) was bit-reproducible in 600 fits
// of the same experiment (profiles/r02_peer_sym_investigation.md), so it is the default.
static bool fx_peer_sym_enabled() {
    static const bool on = []() {
        const char* e = getenv("FXB200_PEER_SYM");
        const char* off = getenv("FXB200_NO_PEER");
        return e && e[0] == '1' && !(off && off[0] == '1');
    }();
    return on;
}

static int peer_stage_alloc(fx_ctx* ctx) {
    if (!ctx->peer_stage) FX_CUDA(cudaMalloc((void**)&ctx->peer_stage, (size_t)(12 + 12 * fx_ctx::MAX_PEERS) * 8));
    return 0;
}

int fx_agree(fx_ctx* ctx, int rc) {
    if (ctx->world <= 1 || !ctx->comm) return rc;
    if (ctx->parent) {
        // rank threads of one process: agree on the host (also keeps every cudaMalloc of the workspace
        // phase ahead of the first collective: with peer access enabled an allocation synchronises
        // with the peer devices and must not meet a rank that already spins inside an NCCL kernel)
        FxMultiShared* ms = ctx->parent->ms;
        ms->ok[ctx->rank] = rc ? 0 : 1;
        ms->barrier();
        int failed = 0;
        for (int c = 0; c < ctx->world; ++c) failed += ms->ok[c] ? 0 : 1;
        ms->barrier();
        if (rc) return rc;
        if (failed) {
            fx_set_error("another device of this multi-GPU context failed before the collective section of this call "
                         "(%d of %d, e.g. out of device memory)", failed, ctx->world);
            return FX_ERR_CUDA;
        }
        return 0;
    }
    if (peer_stage_alloc(ctx)) return rc ? rc : FX_ERR_OOM;       // cannot even talk: local verdict
    double* hp = ctx->pinned + 3400;                    // (0..3071: solver / eigenvalue scratch, 3072..3291: peer exchange)
    hp[0] = rc ? 1.0 : 0.0;
    cudaStream_t s = ctx->stream;
    if (cudaMemcpyAsync(ctx->peer_stage, hp, 8, cudaMemcpyHostToDevice, s) != cudaSuccess) return rc ? rc : FX_ERR_CUDA;
    if (fxcomm_allreduce_sum(ctx->comm, ctx->peer_stage, 1, s)) return rc ? rc : FX_ERR_NCCL;
    if (cudaMemcpyAsync(hp, ctx->peer_stage, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return rc ? rc : FX_ERR_CUDA;
    if (rc) return rc;
    if (hp[0] != 0.0) {
        fx_set_error("another rank failed before the collective section of this call (%d of %d ranks reported an "
                     "error, e.g. out of device memory); this rank gives up with it", (int)hp[0], ctx->world);
        return FX_ERR_CUDA;
    }
    return 0;
}

void fx_peer_close_all(fx_ctx* ctx) {
    for (int c = 0; c < fx_ctx::MAX_PEERS; ++c) {
        if (ctx->peer_arena[c]) {
            cudaIpcCloseMemHandle(ctx->peer_arena[c]);
            cudaGetLastError();
            ctx->peer_arena[c] = nullptr;
        }
        memset(ctx->peer_handle[c], 0, 64);
        ctx->peer_gen[c] = 0.0;
    }
    ctx->peer_valid = false;
}

// In-process variant (fx_init_multi): the rank threads meet on a host barrier and hand each other
// plain device pointers (unified addressing + cudaDeviceEnablePeerAccess, no IPC handles).
static int fx_peer_exchange_inproc(fx_ctx* ctx, size_t off0, size_t off1) {
    ctx->peer_valid = false;
    const int G = ctx->world, r = ctx->rank;
    FxMultiShared* ms = ctx->parent->ms;
    const bool disabled = !fx_peer_sym_enabled();
    int ok = (!disabled && G <= fx_ctx::MAX_PEERS && ctx->arena) ? 1 : 0;
    for (int c = 0; c < G && ok; ++c) {
        if (c == r) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, ctx->parent->sub[c]->device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctx->parent->sub[c]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
        cudaGetLastError();
    }
    ms->arena[r] = ctx->arena; ms->off[r][0] = off0; ms->off[r][1] = off1; ms->ok[r] = ok;
    ms->barrier();
    bool all_ok = true;
    for (int c = 0; c < G; ++c) all_ok = all_ok && ms->ok[c];
    if (all_ok)
        for (int c = 0; c < G; ++c) {
            ctx->peer_arena[c] = (c == r) ? nullptr : ms->arena[c];
            ctx->peer_off[c][0] = ms->off[c][0];
            ctx->peer_off[c][1] = ms->off[c][1];
        }
    ms->barrier();                                         // everybody has read the table
    if (!all_ok) return 0;
    // stream-ordered barrier across the ranks: no rank writes into a peer before that peer is done
    // with its previous use of the buffers
    FX_TRY(peer_stage_alloc(ctx));
    FX_CUDA(cudaMemsetAsync(ctx->peer_stage, 0, 8, ctx->stream));
    FX_TRY(fxcomm_allreduce_sum(ctx->comm, ctx->peer_stage, 1, ctx->stream));
    ctx->peer_valid = true;
    return 0;
}

int fx_peer_exchange(fx_ctx* ctx, size_t off0, size_t off1) {
    if (ctx->parent) return fx_peer_exchange_inproc(ctx, off0, off1);
    ctx->peer_valid = false;
    const int G = ctx->world;
    if (G <= 1 || G > fx_ctx::MAX_PEERS || !ctx->comm || ctx->peer_failed) return 0;
    if (!fx_peer_sym_enabled()) return 0;
    constexpr int W = 12;                                  // doubles per rank in the exchange
    cudaStream_t s = ctx->stream;
    FX_TRY(peer_stage_alloc(ctx));
    double* hp = ctx->pinned + 3072;                       // [W] send, then [G][W] received
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    double ok = 1.0;
    if (!ctx->arena || cudaIpcGetMemHandle(&h, ctx->arena) != cudaSuccess) { cudaGetLastError(); ok = 0.0; }
    else ctx->arena_exported = true;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(hp, &h, 64);
    hp[8] = ok; hp[9] = ctx->arena_gen; hp[10] = (double)off0; hp[11] = (double)off1;
    FX_CUDA(cudaMemcpyAsync(ctx->peer_stage, hp, W * 8, cudaMemcpyHostToDevice, s));
    FX_TRY(fxcomm_allgather(ctx->comm, ctx->peer_stage, ctx->peer_stage + W, W, s));
    FX_CUDA(cudaMemcpyAsync(hp + 16, ctx->peer_stage + W, (size_t)G * W * 8, cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    bool all_ok = true, changed = false;
    for (int c = 0; c < G; ++c) {
        const double* r = hp + 16 + c * W;
        if (r[8] != 1.0) all_ok = false;
        if (memcmp(r, ctx->peer_handle[c], 64) != 0 || r[9] != ctx->peer_gen[c]) changed = true;
        ctx->peer_off[c][0] = (size_t)r[10];
        ctx->peer_off[c][1] = (size_t)r[11];
    }
    if (!all_ok) {                                         // same verdict on every rank
        fx_peer_close_all(ctx);
        ctx->peer_failed = true;
        return 0;
    }
    if (changed) {
        double fails = 0.0;
        for (int c = 0; c < G; ++c) {
            const double* r = hp + 16 + c * W;
            const bool diff = memcmp(r, ctx->peer_handle[c], 64) != 0 || r[9] != ctx->peer_gen[c];
            if (!diff) continue;
            if (c != ctx->rank) {
                if (ctx->peer_arena[c]) { cudaIpcCloseMemHandle(ctx->peer_arena[c]); cudaGetLastError(); ctx->peer_arena[c] = nullptr; }
                cudaIpcMemHandle_t ph;
                memcpy(&ph, r, 64);
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    fails += 1.0;
                } else {
                    ctx->peer_arena[c] = (char*)ptr;
                }
            }
            memcpy(ctx->peer_handle[c], r, 64);
            ctx->peer_gen[c] = r[9];
        }
        // every rank has dropped its stale mappings once this reduction completes
        hp[0] = fails;
        FX_CUDA(cudaMemcpyAsync(ctx->peer_stage, hp, 8, cudaMemcpyHostToDevice, s));
        FX_TRY(fxcomm_allreduce_sum(ctx->comm, ctx->peer_stage, 1, s));
        FX_CUDA(cudaMemcpyAsync(hp, ctx->peer_stage, 8, cudaMemcpyDeviceToHost, s));
        FX_CUDA(cudaStreamSynchronize(s));
        for (char* a : ctx->retired) cudaFree(a);
        ctx->retired.clear();
        if (hp[0] != 0.0) {
            fx_peer_close_all(ctx);
            ctx->peer_failed = true;
            return 0;
        }
    }
    ctx->peer_valid = true;
    return 0;
}

void* fx_arena_alloc(fx_ctx* ctx, size_t bytes) {
    bytes = fx_align256(bytes);
    if (ctx->arena_off + bytes > ctx->arena_cap) return nullptr;   // callers reserve first
    void* p = ctx->arena + ctx->arena_off;
    ctx->arena_off += bytes;
    return p;
}

cudaEvent_t FxTimer::get_event() {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}
void FxTimer::begin(double* acc, cudaStream_t on) {
    Span sp;
    sp.a = get_event();
    sp.b = get_event();
    sp.acc = acc;
    sp.s = on ? on : ctx->stream;
    cudaEventRecord(sp.a, sp.s);
    open_.push_back((int)spans.size());
    spans.push_back(sp);
}
void FxTimer::end() {
    if (open_.empty()) return;
    cudaEventRecord(spans[open_.back()].b, spans[open_.back()].s);
    open_.pop_back();
}
int FxTimer::resolve() {
    for (auto& sp : spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) *sp.acc += ms;
        else cudaGetLastError();
    }
    spans.clear();
    open_.clear();
    ctx->ev_used = 0;
    return 0;
}

static int ctx_common_init(int device, fx_ctx** out) {
    if (!out) { fx_set_error("fx_init: out is NULL"); return -2; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fx_set_error("fx_init: no CUDA device available (%s); this library has no CPU fallback",
                     cudaGetErrorString(e));
        return FX_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { fx_set_error("fx_init: device %d out of range", device); return -1; }
    FX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        fx_set_error("fx_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                     device, prop.major, prop.minor);
        return FX_ERR_CUDA;
    }
    fx_ctx* c = new fx_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    FX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    {
        // highest priority: its single-CTA kernels must get the first SM a K*V CTA vacates, not
        // queue behind the rest of that grid
        int least = 0, greatest = 0;
        FX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        FX_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, greatest));
    }
    FX_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    FX_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    FX_CUDA(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
    c->pinned_cap = 16384;      // 0..4095 solver scalars / peer exchange / agree flag, 8192.. small matrices (klda.cu)
    FX_CUDA(cudaMallocHost((void**)&c->pinned, c->pinned_cap * sizeof(double)));
    FX_TRY(peer_stage_alloc(c));                            // never allocate between collectives
    *out = c;
    return 0;
}

extern "C" int fx_init(int device, fx_ctx** out) { return ctx_common_init(device, out); }

// Single process, several GPUs: one sub-context and one host thread per device,
// NCCL communicators created in-process, peer access enabled directly.  A Julia caller makes ONE
// ccall per fit and never sees NCCL or ranks.
extern "C" int fx_init_multi(const int* devices, int ndev, fx_ctx** out) {
    if (!out) { fx_set_error("fx_init_multi: out is NULL"); return -3; }
    *out = nullptr;
    if (!devices) { fx_set_error("fx_init_multi: devices is NULL"); return -1; }
    if (ndev < 1 || ndev > fx_ctx::MAX_PEERS) { fx_set_error("fx_init_multi: ndev must lie in 1..%d", fx_ctx::MAX_PEERS); return -2; }
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) { fx_set_error("fx_init_multi: device %d listed twice", devices[i]); return -1; }
    fx_ctx* parent = new fx_ctx();
    parent->world = ndev;
    parent->ms = new FxMultiShared();
    parent->ms->n = ndev;
    int rc = 0;
    for (int i = 0; i < ndev && !rc; ++i) {
        fx_ctx* c = nullptr;
        rc = ctx_common_init(devices[i], &c);
        if (rc) break;
        c->rank = i; c->world = ndev; c->parent = parent;
        parent->sub.push_back(c);
    }
    if (!rc && ndev > 1) {
        FxComm* comms[fx_ctx::MAX_PEERS] = {};
        rc = fxcomm_create_all(devices, ndev, comms);
        if (!rc) for (int i = 0; i < ndev; ++i) parent->sub[i]->comm = comms[i];
    }
    if (rc) { fx_shutdown(parent); return rc; }
    parent->device = devices[0];
    parent->sm_count = parent->sub[0]->sm_count;
    *out = parent;
    return 0;
}

// Runs fn(sub-context, rank) on one host thread per device and joins; the first failure's code and
// message become the caller's.
int fx_multi_run(fx_ctx* parent, const std::function<int(fx_ctx*, int)>& fn) {
    const int G = (int)parent->sub.size();
    std::vector<int> rc(G, 0);
    std::vector<std::string> msg(G);
    std::vector<std::thread> th;
    for (int i = 0; i < G; ++i)
        th.emplace_back([&, i]() {
            cudaSetDevice(parent->sub[i]->device);
            rc[i] = fn(parent->sub[i], i);
            if (rc[i]) msg[i] = fx_last_error();
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < G; ++i)
        if (rc[i]) { fx_set_error("%s", msg[i].c_str()); return rc[i]; }
    return 0;
}

extern "C" int fx_comm_unique_id(void* id128) {
    if (!id128) { fx_set_error("fx_comm_unique_id: invalid argument 1 (!id128)"); return -1; }
    return fxcomm_unique_id(id128);
}

extern "C" int fx_init_distributed(int device, int rank, int world, const void* id128, fx_ctx** out) {
    if (world < 1 || rank < 0 || rank >= world) { fx_set_error("bad rank/world %d/%d", rank, world); return -2; }
    if (!id128 && world > 1) { fx_set_error("fx_init_distributed: invalid argument 4 (!id128 && world > 1)"); return -4; }
    FX_TRY(ctx_common_init(device, out));
    fx_ctx* c = *out;
    c->rank = rank;
    c->world = world;
    if (world > 1) {
        int r = fxcomm_create(rank, world, id128, &c->comm);
        if (r) { fx_shutdown(c); *out = nullptr; return r; }
    }
    return 0;
}

extern "C" int fx_set_stream(fx_ctx* ctx, void* cuda_stream) {
    if (!ctx) { fx_set_error("fx_set_stream: ctx is NULL"); return -1; }
    if (!ctx->sub.empty()) { fx_set_error("fx_set_stream: a multi-GPU context owns one stream per device"); return -1; }
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        FX_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return 0;
}

extern "C" int fx_ctx_info(const fx_ctx* ctx, int* device, int* rank, int* world, int* sm_count) {
    if (!ctx) { fx_set_error("fx_ctx_info: ctx is NULL"); return -1; }
    if (device) *device = ctx->device;
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (sm_count) *sm_count = ctx->sm_count;
    return 0;
}

extern "C" int fx_ctx_peer_access(const fx_ctx* ctx) {
    if (ctx && !ctx->sub.empty()) return ctx->sub[0]->peer_valid ? 1 : 0;
    return (ctx && ctx->peer_valid) ? 1 : 0;
}

extern "C" void fx_shutdown(fx_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->sub.empty() || ctx->ms) {
        for (fx_ctx* c : ctx->sub) {
            for (int k = 0; k < fx_ctx::MAX_PEERS; ++k) c->peer_arena[k] = nullptr;   // plain pointers, not IPC mappings
            c->parent = nullptr;
            fx_shutdown(c);
        }
        delete ctx->ms;
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    fx_peer_close_all(ctx);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->comm) fxcomm_destroy(ctx->comm);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->arena) cudaFree(ctx->arena);
    for (char* a : ctx->retired) cudaFree(a);
    if (ctx->peer_stage) cudaFree(ctx->peer_stage);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void fx_model_free(fx_model* m) {
    if (!m) return;
    if (!m->sub.empty()) {
        for (fx_model* r : m->sub) fx_model_free(r);
        delete m;
        return;
    }
    cudaSetDevice(m->device);
    cudaFree(m->Xp);
    cudaFree(m->norms);
    cudaFree(m->W);
    cudaFree(m->bias);
    cudaFree(m->wsum);
    cudaFree(m->colmean);
    delete m;
}

extern "C" int fx_model_info(const fx_model* m, int64_t* n, int64_t* d, int* k, int* kernel_id) {
    if (!m) { fx_set_error("fx_model_info: model is NULL"); return -1; }
    if (!m->sub.empty()) m = m->sub[0];
    if (n) *n = m->n;
    if (d) *d = m->d;
    if (k) *k = m->k;
    if (kernel_id) *kernel_id = m->kernel_id;
    return 0;
}
