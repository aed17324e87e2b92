// comm.cu -- NCCL over NVLink 5 / NVSwitch, resolved with dlopen so the library has no
// link-time dependency on it (a Julia / Python host never sees NCCL).  Only two collectives
// are used on the data path: all-gather of the n x b block vector and all-reduce of the small
// projection matrices / centring sums.  The reference has no distributed path at all.
#include <dlfcn.h>
#include <mutex>
#include "fx_common.cuh"

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int (*pfn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*pfn_CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int);
typedef int (*pfn_CommDestroy)(ncclComm_p);
typedef int (*pfn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef int (*pfn_AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t);
typedef const char* (*pfn_GetErrorString)(int);
typedef int (*pfn_Group)(void);
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    pfn_GetUniqueId GetUniqueId = nullptr;
    pfn_CommInitRank CommInitRank = nullptr;
    pfn_CommDestroy CommDestroy = nullptr;
    pfn_AllReduce AllReduce = nullptr;
    pfn_AllGather AllGather = nullptr;
    pfn_GetErrorString GetErrorString = nullptr;
    pfn_Group GroupStart = nullptr, GroupEnd = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_once;

void load_nccl() {
    // If the host process (e.g. torch) already loaded an NCCL, the SONAME lookup returns it.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    g_nccl.GetUniqueId = (pfn_GetUniqueId)dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (pfn_CommInitRank)dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.CommDestroy = (pfn_CommDestroy)dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.AllReduce = (pfn_AllReduce)dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.AllGather = (pfn_AllGather)dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.GetErrorString = (pfn_GetErrorString)dlsym(g_nccl.handle, "ncclGetErrorString");
    g_nccl.GroupStart = (pfn_Group)dlsym(g_nccl.handle, "ncclGroupStart");
    g_nccl.GroupEnd = (pfn_Group)dlsym(g_nccl.handle, "ncclGroupEnd");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
                g_nccl.AllGather && g_nccl.GetErrorString;
}

int need_nccl() {
    std::call_once(g_once, load_nccl);
    if (!g_nccl.ok) {
        fx_set_error("NCCL not available: dlopen(libnccl.so.2) failed (%s)", dlerror());
        return FX_ERR_NCCL;
    }
    return 0;
}

#define FX_NCCL(expr)                                                                        \
    do {                                                                                     \
        int _r = (expr);                                                                     \
        if (_r != 0) {                                                                       \
            fx_set_error("NCCL error %s at %s:%d", g_nccl.GetErrorString(_r), __FILE__,      \
                         __LINE__);                                                          \
            return FX_ERR_NCCL;                                                              \
        }                                                                                    \
    } while (0)
}  // namespace

struct FxComm {
    ncclComm_p comm = nullptr;
    int rank = 0, world = 1;
};

int fxcomm_unique_id(void* id128) {
    FX_TRY(need_nccl());
    ncclUniqueId_t id;
    FX_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int fxcomm_create(int rank, int world, const void* id128, FxComm** out) {
    FX_TRY(need_nccl());
    ncclUniqueId_t id;
    memcpy(&id, id128, sizeof(id));
    FxComm* c = new FxComm();
    c->rank = rank;
    c->world = world;
    int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        fx_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return FX_ERR_NCCL;
    }
    *out = c;
    return 0;
}

int fxcomm_create_all(const int* devices, int ndev, FxComm** out) {
    FX_TRY(need_nccl());
    if (!g_nccl.GroupStart || !g_nccl.GroupEnd) { fx_set_error("NCCL: ncclGroupStart/End not found"); return FX_ERR_NCCL; }
    ncclUniqueId_t id;
    FX_NCCL(g_nccl.GetUniqueId(&id));
    for (int i = 0; i < ndev; ++i) { out[i] = new FxComm(); out[i]->rank = i; out[i]->world = ndev; }
    FX_NCCL(g_nccl.GroupStart());
    int rc = 0;
    for (int i = 0; i < ndev && !rc; ++i) {
        if (cudaSetDevice(devices[i]) != cudaSuccess) { rc = -1; break; }
        rc = g_nccl.CommInitRank(&out[i]->comm, ndev, id, i);
    }
    const int re = g_nccl.GroupEnd();
    if (rc || re) {
        fx_set_error("ncclCommInitRank (single process, %d devices) failed: %s", ndev,
                     g_nccl.GetErrorString(rc > 0 ? rc : re));
        for (int i = 0; i < ndev; ++i) { delete out[i]; out[i] = nullptr; }
        return FX_ERR_NCCL;
    }
    return 0;
}

void fxcomm_destroy(FxComm* c) {
    if (!c) return;
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    delete c;
}

int fxcomm_allreduce_sum(FxComm* c, double* buf, size_t count, cudaStream_t s) {
    if (!c || c->world == 1 || count == 0) return 0;
    FX_NCCL(g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, c->comm, s));
    return 0;
}

int fxcomm_allgather(FxComm* c, const double* send, double* recv, size_t count_per_rank,
                     cudaStream_t s) {
    if (!c || c->world == 1) {
        if (send != recv && count_per_rank)
            FX_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    FX_NCCL(g_nccl.AllGather(send, recv, count_per_rank, kNcclFloat64, c->comm, s));
    return 0;
}
