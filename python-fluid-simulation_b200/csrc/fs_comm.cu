#include <dlfcn.h>

#include "fs_comm.cuh"

// minimal NCCL surface (ABI-stable since NCCL 2.x); declared here so no nccl.h is needed at build time
extern "C" {
typedef struct { char internal[128]; } fs_ncclUniqueId;
typedef int ncclResult_t_;
}

namespace {

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(fs_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, fs_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

// ncclDataType_t / ncclRedOp_t values (nccl.h): ncclUint8 = 1, ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0
constexpr int kNcclUint8 = 1, kNcclF32 = 7, kNcclF64 = 8, kNcclSum = 0;

int nccl_load() {
    if (g_nccl.lib) return FS_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fs::fail(FS_ERR_STATE, "cannot dlopen libnccl.so.2: %s", dlerror());
#define FS_SYM(field, name)                                                                   \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                               \
    if (!g_nccl.field) return fs::fail(FS_ERR_STATE, "libnccl is missing symbol %s", name);
    FS_SYM(GetUniqueId, "ncclGetUniqueId")
    FS_SYM(CommInitRank, "ncclCommInitRank")
    FS_SYM(CommDestroy, "ncclCommDestroy")
    FS_SYM(AllReduce, "ncclAllReduce")
    FS_SYM(Send, "ncclSend")
    FS_SYM(Recv, "ncclRecv")
    FS_SYM(GroupStart, "ncclGroupStart")
    FS_SYM(GroupEnd, "ncclGroupEnd")
    FS_SYM(GetErrorString, "ncclGetErrorString")
#undef FS_SYM
    g_nccl.lib = lib;
    return FS_OK;
}

#define FS_NCCL(expr)                                                                                       \
    do {                                                                                                    \
        int _r = (expr);                                                                                    \
        if (_r != 0) return fs::fail(FS_ERR_CUDA, "%s: %s", #expr, g_nccl.GetErrorString(_r));               \
    } while (0)

}  // namespace

namespace fs {

int comm_allreduce_sum_f64(fs_comm* c, double* p, int count, cudaStream_t s) {
    FS_NCCL(g_nccl.AllReduce(p, p, (size_t)count, kNcclF64, kNcclSum, c->nccl_comm, s));
    return FS_OK;
}

int comm_halo_exchange(fs_comm* c, int has_lo, int has_hi, int nplanes, const void* const* send_lo, void* const* recv_lo,
                       const void* const* send_hi, void* const* recv_hi, size_t count, int type, cudaStream_t s) {
    const int t = type == COMM_F32 ? kNcclF32 : (type == COMM_F64 ? kNcclF64 : kNcclUint8);
    if (!has_lo && !has_hi) return FS_OK;
    FS_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < nplanes; ++k) {
        if (has_lo) {
            FS_NCCL(g_nccl.Send(send_lo[k], count, t, c->rank - 1, c->nccl_comm, s));
            FS_NCCL(g_nccl.Recv(recv_lo[k], count, t, c->rank - 1, c->nccl_comm, s));
        }
        if (has_hi) {
            FS_NCCL(g_nccl.Send(send_hi[k], count, t, c->rank + 1, c->nccl_comm, s));
            FS_NCCL(g_nccl.Recv(recv_hi[k], count, t, c->rank + 1, c->nccl_comm, s));
        }
    }
    FS_NCCL(g_nccl.GroupEnd());
    return FS_OK;
}

}  // namespace fs

extern "C" {

int fs_comm_unique_id(void* out128) {
    if (!out128) return fs::fail(FS_ERR_ARG, "fs_comm_unique_id: null argument");
    FS_TRY(nccl_load());
    fs_ncclUniqueId id;
    FS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return FS_OK;
}

int fs_comm_create(fs_comm** out, int rank, int nranks, const void* id128) {
    if (!out || !id128) return fs::fail(FS_ERR_ARG, "fs_comm_create: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fs::fail(FS_ERR_ARG, "fs_comm_create: bad rank / nranks");
    FS_TRY(nccl_load());
    fs_ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    fs_comm* c = new fs_comm();
    c->rank = rank;
    c->nranks = nranks;
    int r = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (r != 0) { delete c; return fs::fail(FS_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    *out = c;
    return FS_OK;
}

void fs_comm_destroy(fs_comm* c) {
    if (!c) return;
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    delete c;
}

// ---- CUDA-IPC shared device buffers (peer-mapped over NVLink) ----------------------------------------
void* fs_shared_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { fs::fail(FS_ERR_CUDA, "fs_shared_alloc: cudaMalloc failed"); return nullptr; }
    if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); fs::fail(FS_ERR_CUDA, "fs_shared_alloc: cudaMemset failed"); return nullptr; }
    return p;
}

void fs_shared_free(void* p) {
    if (p) cudaFree(p);
}

int fs_shared_get_handle(void* p, void* out64) {
    if (!p || !out64) return fs::fail(FS_ERR_ARG, "fs_shared_get_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    FS_CUDA(cudaIpcGetMemHandle(&hd, p));
    memcpy(out64, &hd, 64);
    return FS_OK;
}

void* fs_shared_open(const void* handle64) {
    if (!handle64) { fs::fail(FS_ERR_ARG, "fs_shared_open: null argument"); return nullptr; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { fs::fail(FS_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return nullptr; }
    return p;
}

void fs_shared_close(void* mapped) {
    if (mapped) cudaIpcCloseMemHandle(mapped);
}

int fs_comm_rank(const fs_comm* c) { return c ? c->rank : -1; }
int fs_comm_size(const fs_comm* c) { return c ? c->nranks : -1; }

}  // extern "C"
