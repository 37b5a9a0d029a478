// Library-wide state and the small non-template pieces of the CG control.
#include <stdlib.h>

#include "fs_common.cuh"

namespace fs {

thread_local char g_err[512] = "";
long long g_launches = 0;

int coop_max_blocks(const void* fn, int threads, size_t dyn_smem) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, dyn_smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return sms * per_sm;
}

__global__ void cg_state_init_kernel(CgState* st, double tol2, long long max_iter, int dist) {
    st->delta = 0.0;
    st->delta_old = 0.0;
    st->dq = 0.0;
    st->alpha = 0.0;
    st->beta = 0.0;
    st->tol2 = tol2;
    st->delta0 = 0.0;
    st->iter = 0;
    st->max_iter = max_iter;
    st->done = 0;
    st->dist = dist;
    st->red = 0.0;
    for (int i = 0; i < 4; ++i) st->counter[i] = 0;
    st->sr_first = 1;
    st->sr_parity = 0;
}

__global__ void cg_finish_kernel(CgState* st, int which) {
    if (st->done) return;
    const double s = st->red;
    if (which == 0) {
        st->delta = s;
        st->delta0 = s;
        st->delta_old = s;
        if (s < st->tol2) st->done = 1;
    } else {
        st->delta_old = st->delta;
        st->delta = s;
        st->iter += 1;
        if (s < st->tol2) st->done = 1;
        else if (st->iter >= st->max_iter || !(s == s)) st->done = 2;
    }
}

// bench hook: let already-initialised state run for an unbounded number of iterations (no convergence stop)
__global__ void cg_state_unlimit_kernel(CgState* st) {
    st->done = 0;
    st->tol2 = -1.0;
    st->max_iter = 0x7fffffffffffffffLL;
}

// ---- active-segment list -------------------------------------------------------------------------------------
__device__ __forceinline__ bool seg_listed(const uint8_t* __restrict__ act, long long s, long long nseg_total, int per_seg_flags) {
    if (per_seg_flags) return s < nseg_total && act[s] != 0;
    return seg_flag(act, s, nseg_total);
}

__global__ void seg_count_kernel(const uint8_t* act, long long nseg_total, int* block_count, int per_seg_flags) {
    const long long s = (long long)blockIdx.x * kSegBlock + threadIdx.x;
    const int n = __syncthreads_count(seg_listed(act, s, nseg_total, per_seg_flags) ? 1 : 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = n;
}

__global__ void seg_scan_kernel(int* block_count, int nblocks, int* nseg_out) {
    __shared__ int wtot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const int v = i < nblocks ? block_count[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int t = wtot[lane];
            int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            wtot[lane] = ti - t;                 // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int c = carry;
        if (i < nblocks) block_count[i] = c + wtot[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wtot[w] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nseg_out = carry;
}

__global__ void seg_write_kernel(const uint8_t* act, long long nseg_total, const int* block_off, int* list, int per_seg_flags) {
    __shared__ int wsum[kSegBlock / 32];
    const long long s = (long long)blockIdx.x * kSegBlock + threadIdx.x;
    const bool f = seg_listed(act, s, nseg_total, per_seg_flags);
    const unsigned int m = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = block_off[blockIdx.x];
    for (int k = 0; k < w; ++k) off += wsum[k];
    if (f) list[off + __popc(m & ((1u << lane) - 1u))] = (int)s;
}

int SegList::init(long long npts, void* list_dev, void* scratch_dev) {
    nseg_total = (npts + kSegPts - 1) / kSegPts;
    nblocks = (int)((nseg_total + kSegBlock - 1) / kSegBlock);
    list = (int*)list_dev;
    block_off = (int*)scratch_dev;
    nseg_dev = block_off + nblocks + 1;
    nseg = 0;
    if (!nseg_pinned) FS_CUDA(cudaHostAlloc((void**)&nseg_pinned, sizeof(int), cudaHostAllocDefault));
    return FS_OK;
}

void SegList::destroy() {
    if (nseg_pinned) cudaFreeHost(nseg_pinned);
    nseg_pinned = nullptr;
    if (ready) cudaEventDestroy(ready);
    ready = nullptr;
    pending = false;
}

int SegList::enqueue(const uint8_t* act, cudaStream_t s, int per_seg_flags) {
    if (!ready) FS_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    seg_count_kernel<<<nblocks, kSegBlock, 0, s>>>(act, nseg_total, block_off, per_seg_flags);
    FS_LAUNCH_CHECK();
    seg_scan_kernel<<<1, 1024, 0, s>>>(block_off, nblocks, nseg_dev);
    FS_LAUNCH_CHECK();
    seg_write_kernel<<<nblocks, kSegBlock, 0, s>>>(act, nseg_total, block_off, list, per_seg_flags);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaMemcpyAsync(nseg_pinned, nseg_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    FS_CUDA(cudaEventRecord(ready, s));
    pending = true;
    return FS_OK;
}

int SegList::finish() {
    if (!pending) return FS_OK;
    FS_CUDA(cudaEventSynchronize(ready));
    pending = false;
    nseg = *nseg_pinned;
    return FS_OK;
}

int SegList::build(const uint8_t* act, cudaStream_t s) {
    FS_TRY(enqueue(act, s));
    return finish();
}

bool IterGraph::enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("FLUIDSOLVER_B200_GRAPH");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

int CgHost::init() {
    FS_CUDA(cudaHostAlloc((void**)&st_pinned, 2 * sizeof(CgState), cudaHostAllocDefault));
    memset(st_pinned, 0, 2 * sizeof(CgState));
    for (int i = 0; i < 2; ++i) FS_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    return FS_OK;
}

void CgHost::destroy() {
    if (st_pinned) cudaFreeHost(st_pinned);
    st_pinned = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (ev[i]) cudaEventDestroy(ev[i]);
        ev[i] = nullptr;
    }
}

// tuning switches: explicit setting (fs_set_option) > environment variable > built-in default (-1)
static int g_opt[OPT_COUNT] = {-2, -2, -2, -2};
static const char* const kOptEnv[OPT_COUNT] = {"FLUIDSOLVER_B200_RESIDENT", "FLUIDSOLVER_B200_K1BLOCK", "FLUIDSOLVER_B200_K1TILE", "FLUIDSOLVER_B200_SPARSE_SETUP"};
static const char* const kOptName[OPT_COUNT] = {"resident_form", "k1_block", "k1_tile", "sparse_setup"};

static int g_opt_epoch = 0;
int tuning_epoch() { return g_opt_epoch; }

int tuning(int which) {
    if (which < 0 || which >= OPT_COUNT) return -1;
    if (g_opt[which] == -2) {
        const char* e = getenv(kOptEnv[which]);
        g_opt[which] = (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : -1;
    }
    return g_opt[which];
}

}  // namespace fs

extern "C" {

int fs_set_option(const char* name, int value) {
    if (!name) return fs::fail(FS_ERR_ARG, "fs_set_option: null name");
    for (int k = 0; k < fs::OPT_COUNT; ++k)
        if (strcmp(name, fs::kOptName[k]) == 0) { fs::g_opt[k] = value < 0 ? -1 : value; ++fs::g_opt_epoch; return FS_OK; }
    return fs::fail(FS_ERR_ARG, "fs_set_option: unknown option '%s'", name);
}

int fs_abi_version(void) { return FS_ABI_VERSION; }
const char* fs_last_error(void) { return fs::g_err; }
int64_t fs_launch_count(void) { return fs::g_launches; }

}  // extern "C"
