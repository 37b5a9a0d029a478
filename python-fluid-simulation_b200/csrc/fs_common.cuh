// Shared device/host helpers for the B200 (sm_100a) implicit-solver kernels:
// error plumbing for the C ABI, deterministic block/grid reductions, and the fused CG
// vector kernels (K2: x,r update + r.r ; K3: d update) that replace the CuPy expression
// chains of the reference's solve() loops (ViscosityCGSolver3D.py:592-610,
// PressureCGSolver3D.py:211-221).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fluidsolver_b200.h"

namespace fs {

// ---------------------------------------------------------------------------------------------
// error handling (C ABI never throws)
// ---------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern long long g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define FS_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) return ::fs::fail(FS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define FS_LAUNCH_CHECK()                                                                   \
    do {                                                                                   \
        ++::fs::g_launches;                                                                \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) return ::fs::fail(FS_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
    } while (0)

#define FS_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s < 0) return _s;       \
    } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Largest grid a cooperative launch of `fn` with `threads` threads per block can have on the CURRENT context: the SM count
// the context exposes (smaller than kSMs under MPS active-thread percentages, green contexts, other sm_100 SKUs) times
// the resident blocks per SM.  0 on error.
int coop_max_blocks(const void* fn, int threads, size_t dyn_smem = 0);

// tuning switches (fs_set_option / FLUIDSOLVER_B200_* environment variables): -1 = not set, use the built-in default
enum { OPT_RESIDENT_FORM = 0, OPT_K1BLOCK = 1, OPT_K1TILE = 2, OPT_SPARSE_SETUP = 3, OPT_COUNT = 4 };
int tuning(int which);
int tuning_epoch();      // bumped by every fs_set_option call: captured iteration graphs are keyed on it

// ---------------------------------------------------------------------------------------------
// device-resident CG control block: alpha/beta/convergence live on the GPU so the iteration
// needs no host round trip (the reference does two blocking .item() reads per iteration).
// ---------------------------------------------------------------------------------------------
struct CgState {
    double delta;      // current  r.r
    double delta_old;  // previous r.r
    double dq;         // d.q of the current iteration
    double alpha, beta;
    double tol2;       // tol*tol
    double delta0;
    long long iter;
    long long max_iter;
    int done;          // 0 running, 1 converged, 2 iteration budget exhausted
    int dist;          // 1: multi-GPU — reducing kernels leave their LOCAL sum in `red`; cg_finish_kernel applies it after the allreduce
    double red;
    unsigned int counter[4];  // last-block tickets (one per reducing kernel type)
    // single-reduction (Chronopoulos-Gear) CG: `first` = no search direction yet (beta = 0, alpha = gamma/(w.r))
    int sr_first;
    int sr_parity;     // which w buffer the last K1s wrote (multi-GPU: iteration parity; single GPU: always 0)
};

// ---------------------------------------------------------------------------------------------
// Multi-GPU peer block (device memory).  Pointers are CUDA-IPC mappings of the other ranks' buffers
// (all GPUs of one NVSwitch box), so kernels exchange data with plain st.global / ld.global:
//   * q_lo / q_hi : the neighbours' halo planes of q — K1 stores its boundary rows there directly;
//   * mbox[r]     : rank r's scalar mailbox — the all-reduce of d.q and r.r is done by the last
//                   block of K1 / K2 itself (LL-style 8-byte words carrying 4 data bytes + a 4-byte
//                   sequence flag, so no fence/flag round trip is needed on the reader side).
// The reduction order is rank 0..N-1 on every rank, so all ranks obtain bit-identical sums and take
// identical convergence decisions.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kMboxWords = 2 /*kinds*/ * 2 /*parity*/ * kMaxRanks * 2 /*words*/;

struct PeerInfo {
    int rank, nranks;
    int has_lo, has_hi;
    char* q_lo[3];                         // low neighbour's plane X'-2 of q, per component (peer-mapped)
    char* q_hi[3];                         // high neighbour's plane 0 of q
    unsigned long long* mbox[kMaxRanks];   // every rank's mailbox (own entry is the local pointer)
    unsigned int seq[2];                   // persistent sequence numbers of the two reductions (never reset)
    int error;                             // set to 1 if a peer did not answer within the spin budget
    // halo rows inside each component array of `comp_len` elements (they mirror rows OWNED by a neighbour and must not
    // be counted in r.r): [0, halo_lo_end) and [halo_hi_begin, halo_hi_end)
    long long comp_len, halo_lo_end, halo_hi_begin, halo_hi_end;
};

// The fields the per-point / per-chunk code paths need, passed to kernels BY VALUE (constant bank) so that no global
// load sits in the inner loops; the PeerInfo pointer is only dereferenced in the reduction tail.
struct PeerHot {
    int has_lo, has_hi;
    char* q_lo[3];
    char* q_hi[3];
    // the same planes of the neighbours' w = A r buffers (single-reduction CG), [parity*3 + component].  With ONE cross-GPU
    // synchronisation per iteration a fast rank is already pushing the boundary rows of iteration k+1 while a slow neighbour
    // still reads those of iteration k in its update phase, so w is double-buffered by iteration parity.
    char* w_lo[6];
    char* w_hi[6];
    long long comp_len, halo_lo_end, halo_hi_begin, halo_hi_end;
    long long hb[6], he[6];   // the same halo rows as [begin,end) intervals of the flat 3-component index (empty if unused)
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All 32 lanes of ONE warp call this with the same `local`; returns the global sum (same bits on every rank).
__device__ __forceinline__ double peer_allreduce_warp(double local, PeerInfo* P, int kind) {
    const int lane = threadIdx.x & 31;
    unsigned int seq = 0;
    if (lane == 0) { seq = P->seq[kind] + 1; P->seq[kind] = seq; }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int n = P->nranks;
    const int slot = ((kind * 2 + (int)(seq & 1u)) * kMaxRanks);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
    const unsigned long long w0 = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)seq << 32) | (bits >> 32);
    __threadfence_system();                      // everything this rank wrote for its peers is visible first
    if (lane < n) {                              // lane r publishes to rank r's mailbox (including our own)
        unsigned long long* dst = P->mbox[lane] + (slot + P->rank) * 2;
        st_sys_u64(dst, w0);
        st_sys_u64(dst + 1, w1);
    }
    double v = 0.0;
    bool ok = true;
    if (lane < n) {                              // lane r collects rank r's contribution from OUR mailbox
        const unsigned long long* src = P->mbox[P->rank] + (slot + lane) * 2;
        unsigned long long a, b;
        const long long t0 = clock64();
        for (;;) {
            a = ld_sys_u64(src);
            b = ld_sys_u64(src + 1);
            if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
            if (clock64() - t0 > 20000000000LL) { ok = false; break; }     // ~10 s: a peer died; bail out instead of hanging
        }
        v = __longlong_as_double((long long)(((b & 0xffffffffull) << 32) | (a & 0xffffffffull)));
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) P->error = 1;
        return __longlong_as_double(0x7ff8000000000000LL);
    }
    double sum = 0.0;
    for (int r = 0; r < n; ++r) sum += __shfl_sync(0xffffffffu, v, r);        // fixed rank order
    __threadfence_system();
    return sum;
}

// Two scalars in ONE mailbox round (single-reduction CG): kind 0 carries `a`, kind 1 carries `b`; both sequence numbers
// advance together.  All 32 lanes of one warp call this; results replace a and b (same bits on every rank).
__device__ __forceinline__ void peer_allreduce2_warp(double& a, double& b, PeerInfo* P) {
    const int lane = threadIdx.x & 31;
    unsigned int seq0 = 0, seq1 = 0;
    if (lane == 0) { seq0 = P->seq[0] + 1; P->seq[0] = seq0; seq1 = P->seq[1] + 1; P->seq[1] = seq1; }
    seq0 = __shfl_sync(0xffffffffu, seq0, 0);
    seq1 = __shfl_sync(0xffffffffu, seq1, 0);
    const int n = P->nranks;
    const int kind = lane >> 4, rk = lane & 15;              // lanes 0..15: kind 0, lanes 16..31: kind 1
    const unsigned int seq = kind ? seq1 : seq0;
    const int slot = ((kind * 2 + (int)(seq & 1u)) * kMaxRanks);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(kind ? b : a);
    const unsigned long long w0 = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)seq << 32) | (bits >> 32);
    __threadfence_system();
    if (rk < n) {
        unsigned long long* dst = P->mbox[rk] + (slot + P->rank) * 2;
        st_sys_u64(dst, w0);
        st_sys_u64(dst + 1, w1);
    }
    double v = 0.0;
    bool ok = true;
    if (rk < n) {
        const unsigned long long* src = P->mbox[P->rank] + (slot + rk) * 2;
        unsigned long long x, y;
        const long long t0 = clock64();
        for (;;) {
            x = ld_sys_u64(src);
            y = ld_sys_u64(src + 1);
            if ((unsigned int)(x >> 32) == seq && (unsigned int)(y >> 32) == seq) break;
            if (clock64() - t0 > 20000000000LL) { ok = false; break; }
        }
        v = __longlong_as_double((long long)(((y & 0xffffffffull) << 32) | (x & 0xffffffffull)));
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) P->error = 1;
        a = b = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    double sa = 0.0, sb = 0.0;
    for (int r = 0; r < n; ++r) { sa += __shfl_sync(0xffffffffu, v, r); sb += __shfl_sync(0xffffffffu, v, 16 + r); }   // fixed rank order
    __threadfence_system();
    a = sa; b = sb;
}

// ---------------------------------------------------------------------------------------------
// reductions: warp shuffle -> shared -> one partial per block -> last block sums the partials
// in a fixed order (deterministic run to run; no floating-point atomics).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// all threads of the block must call; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm /*>=32 doubles*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sm reuse
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nwarps ? sm[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Grid-wide sum with a "last block finishes" epilogue.  `fin(total)` runs in thread 0 of the last
// block to arrive, after every block's partial is visible.  blockDim must be 1-D.
template <class Fin>
__device__ __forceinline__ void grid_sum_finish(double v, double* partials, unsigned int* counter, Fin fin,
                                                PeerInfo* peers = nullptr, int kind = 0, bool wrote_peer = false) {
    __shared__ double sm[32];
    __shared__ bool is_last;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    v = block_sum(v, sm);
    if (threadIdx.x == 0) {
        partials[bid] = v;
        if (wrote_peer) __threadfence_system(); else __threadfence();   // system scope only if this block stored into a peer GPU
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) s += __ldcg(partials + i);
        s = block_sum(s, sm);
        if (peers && threadIdx.x < 32) {          // fused all-reduce over the NVSwitch peers (warp 0 of the last block)
            s = __shfl_sync(0xffffffffu, s, 0);
            s = peer_allreduce_warp(s, peers, kind);
        }
        if (threadIdx.x == 0) {
            *counter = 0;
            fin(s);
        }
    }
}

// Two sums in one pass (single-reduction CG: r.r and w.r travel together).  partials: 2 * nblocks doubles.
template <class Fin>
__device__ __forceinline__ void grid_sum2_finish(double v1, double v2, double* partials, unsigned int* counter, Fin fin,
                                                 PeerInfo* peers = nullptr, bool wrote_peer = false) {
    __shared__ double sm[32];
    __shared__ bool is_last;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    v1 = block_sum(v1, sm);
    v2 = block_sum(v2, sm);
    if (threadIdx.x == 0) {
        partials[bid] = v1;
        partials[nblocks + bid] = v2;
        if (wrote_peer) __threadfence_system(); else __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s1 = 0.0, s2 = 0.0;
        for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) { s1 += __ldcg(partials + i); s2 += __ldcg(partials + nblocks + i); }
        s1 = block_sum(s1, sm);
        s2 = block_sum(s2, sm);
        if (peers && threadIdx.x < 32) {          // both scalars cross NVLink in the same mailbox round (kinds 0 and 1)
            s1 = __shfl_sync(0xffffffffu, s1, 0);
            s2 = __shfl_sync(0xffffffffu, s2, 0);
            peer_allreduce2_warp(s1, s2, peers);
        }
        if (threadIdx.x == 0) {
            *counter = 0;
            fin(s1, s2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// vector types for 16-byte accesses
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ void v_get(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void v_get(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 v_make(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 v_make(const double* o) { return make_double2(o[0], o[1]); }

constexpr int kVecThreads = 256;
constexpr int kVecBlocksPerSM = 8;
constexpr int kVecGrid = kSMs * kVecBlocksPerSM;  // persistent-style grid for streaming kernels

// Streaming access helper: VEC elements per 16-byte access when the arrays are 16-byte aligned
// (VEC = Vec16<T>::N), scalar otherwise (VEC = 1).  The scalar tail n % VEC is handled by the same loop body.
template <typename T, int VEC> struct Chunk {
    T a[VEC];
    __device__ __forceinline__ void load(const T* p, long long i) {
        if constexpr (VEC == 1) a[0] = p[i];
        else v_get(reinterpret_cast<const typename Vec16<T>::type*>(p)[i], a);
    }
    __device__ __forceinline__ void store(T* p, long long i) const {
        if constexpr (VEC == 1) p[i] = a[0];
        else reinterpret_cast<typename Vec16<T>::type*>(p)[i] = v_make(a);
    }
};

#define FS_STREAM_SETUP(n, VEC)                                             \
    const long long _nv = (n) / (VEC);                                      \
    const long long _stride = (long long)gridDim.x * blockDim.x;            \
    const long long _t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;

// K2:  alpha = delta/dq ; x += alpha d ; r -= alpha q ; delta' = r.r ; convergence bookkeeping.
// (ViscosityCGSolver3D.py:594-606 / PressureCGSolver3D.py:211-219)
// DIST (multi-GPU, fused transport): halo rows are updated like any other row (their q was stored by the owner), but
// only rows owned by this rank enter r.r; the all-reduce over the peers runs in the tail of this kernel.
template <typename T, int VEC, bool DIST>
__global__ void __launch_bounds__(kVecThreads) cg_update_xr_kernel(long long n, T* __restrict__ x, T* __restrict__ r,
                                                                   const T* __restrict__ d, const T* __restrict__ q,
                                                                   CgState* st, double* partials, int freeze, PeerInfo* peers, PeerHot hot) {
    if (*(volatile int*)&st->done) return;
    const double alpha_d = st->delta / st->dq;
    const T alpha = (T)alpha_d;
    double acc = 0.0;
    FS_STREAM_SETUP(n, VEC)
    for (long long i = _t0; i < _nv; i += _stride) {
        Chunk<T, VEC> xv, rv, dv, qv;
        xv.load(x, i); rv.load(r, i); dv.load(d, i); qv.load(q, i);
        bool own = true;
        if (DIST) {                                   // a 16-byte chunk never straddles a plane (plane size % 4 == 0)
            const long long e = i * VEC;
#pragma unroll
            for (int k = 0; k < 6; ++k) own = own && !(e >= hot.hb[k] && e < hot.he[k]);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            xv.a[k] = xv.a[k] + alpha * dv.a[k];
            rv.a[k] = rv.a[k] - alpha * qv.a[k];
            if (own) acc += (double)rv.a[k] * (double)rv.a[k];
        }
        xv.store(x, i); rv.store(r, i);
    }
    for (long long i = _nv * VEC + _t0; i < n; i += _stride) {   // scalar tail (n % VEC elements; never a halo row)
        x[i] = x[i] + alpha * d[i];
        const T rr = r[i] - alpha * q[i];
        r[i] = rr;
        acc += (double)rr * (double)rr;
    }
    grid_sum_finish(acc, partials, &st->counter[1], [=](double s) {
        if (freeze) return;                      // profiling hook: keep alpha/delta fixed across repeated launches
        st->alpha = alpha_d;
        if (st->dist) { st->red = s; return; }
        st->delta_old = st->delta;
        st->delta = s;
        st->iter += 1;
        if (s < st->tol2) st->done = 1;
        else if (st->iter >= st->max_iter || !(s == s)) st->done = 2;  // NaN: the reference would spin to max_iter
    }, (DIST && !freeze) ? peers : nullptr, 1, false);
}

// K3:  beta = delta/delta_old ; d = r + beta d      (ViscosityCGSolver3D.py:607-610)
template <typename T, int VEC>
__global__ void __launch_bounds__(kVecThreads) cg_update_d_kernel(long long n, T* __restrict__ d, const T* __restrict__ r, CgState* st) {
    if (*(volatile int*)&st->done) return;
    const double beta_d = st->delta / st->delta_old;
    const T beta = (T)beta_d;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->beta = beta_d;
    FS_STREAM_SETUP(n, VEC)
    long long i = _t0;
    for (; i + _stride < _nv; i += 2 * _stride) {     // two chunks per trip: four 16-byte loads in flight per thread
        Chunk<T, VEC> d0, r0, d1, r1;
        d0.load(d, i); r0.load(r, i); d1.load(d, i + _stride); r1.load(r, i + _stride);
#pragma unroll
        for (int k = 0; k < VEC; ++k) { d0.a[k] = r0.a[k] + beta * d0.a[k]; d1.a[k] = r1.a[k] + beta * d1.a[k]; }
        d0.store(d, i); d1.store(d, i + _stride);
    }
    for (; i < _nv; i += _stride) {
        Chunk<T, VEC> dv, rv;
        dv.load(d, i); rv.load(r, i);
#pragma unroll
        for (int k = 0; k < VEC; ++k) dv.a[k] = rv.a[k] + beta * dv.a[k];
        dv.store(d, i);
    }
    for (long long i = _nv * VEC + _t0; i < n; i += _stride) d[i] = r[i] + beta * d[i];
}

// start of a solve:  d = b - q ; r = d ; delta0 = r.r   (ViscosityCGSolver3D.py:577-587)
template <typename T, int VEC>
__global__ void __launch_bounds__(kVecThreads) cg_residual_init_kernel(long long n, const T* __restrict__ b, const T* __restrict__ q,
                                                                       T* __restrict__ d, T* __restrict__ r, CgState* st, double* partials, PeerInfo* peers) {
    double acc = 0.0;
    FS_STREAM_SETUP(n, VEC)
    for (long long i = _t0; i < _nv; i += _stride) {
        Chunk<T, VEC> bv, qv;
        bv.load(b, i); qv.load(q, i);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            bv.a[k] = bv.a[k] - qv.a[k];
            acc += (double)bv.a[k] * (double)bv.a[k];
        }
        bv.store(d, i); bv.store(r, i);
    }
    for (long long i = _nv * VEC + _t0; i < n; i += _stride) {
        const T dd = b[i] - q[i];
        d[i] = dd; r[i] = dd;
        acc += (double)dd * (double)dd;
    }
    grid_sum_finish(acc, partials, &st->counter[2], [=](double s) {
        if (st->dist) { st->red = s; return; }
        st->delta = s;
        st->delta0 = s;
        st->delta_old = s;
        if (s < st->tol2) st->done = 1;           // `if not self.delta < tol ** 2:` skips the loop
    }, peers, 1);
}

// ---------------------------------------------------------------------------------------------
// Active-segment lists.  Solver vectors are exactly zero (and x is never changed) on every row the
// operator does not compute, so the CG kernels only need to visit the rows that are computed.  The
// lattice / cell array is cut into segments of kSegPts consecutive points; a segment is ACTIVE when
// any of its points carries a computed row.  The sorted list of active segments is built once per
// solve (count -> scan -> write, deterministic order) and K1/K2/K3 walk the list instead of the
// whole array: traffic and time scale with the active set, not with the grid.
// ---------------------------------------------------------------------------------------------
constexpr int kSegPts = 32;
constexpr int kSegBlock = 256;            // segments handled per block by the list-building kernels
constexpr unsigned int kActCompute = 0x07u;   // bits 0..2: row of component c is computed by this rank
constexpr unsigned int kActHalo = 0x70u;      // bits 4..6: row is computed by a neighbour slab and mirrored here (multi-GPU)

__device__ __forceinline__ bool seg_flag(const uint8_t* __restrict__ act, long long s, long long nseg_total) {
    if (s >= nseg_total) return false;
    const uint4* p = reinterpret_cast<const uint4*>(act + s * kSegPts);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    return ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) & 0x77777777u) != 0u;
}

// per_seg_flags = 0: `act` holds one activity byte per POINT (a segment is listed if any of its 32 bytes has a row bit);
//               = 1: `act` holds one flag byte per SEGMENT (listed if non-zero)
__global__ void seg_count_kernel(const uint8_t* act, long long nseg_total, int* block_count, int per_seg_flags);
// exclusive scan of block_count[0..nblocks) in place; total -> *nseg_out
__global__ void seg_scan_kernel(int* block_count, int nblocks, int* nseg_out);
__global__ void seg_write_kernel(const uint8_t* act, long long nseg_total, const int* block_off, int* list, int per_seg_flags);

struct SegList {
    int* list = nullptr;        // device: active segment ids, ascending
    int* block_off = nullptr;   // device: nblocks ints (scratch of the scan)
    int* nseg_dev = nullptr;    // device: number of active segments
    int* nseg_pinned = nullptr; // host mirror
    long long nseg_total = 0;   // segments in the whole array
    int nblocks = 0;
    int nseg = 0;               // host copy, valid after build()
    static size_t list_bytes(long long npts) { return (size_t)((npts + kSegPts - 1) / kSegPts) * sizeof(int); }
    static size_t scratch_bytes(long long npts) {
        const long long ns = (npts + kSegPts - 1) / kSegPts;
        return (size_t)((ns + kSegBlock - 1) / kSegBlock + 8) * sizeof(int);
    }
    int init(long long npts, void* list_dev, void* scratch_dev);
    void destroy();
    // act: one byte per point, readable up to 32*nseg_total bytes.  build() = enqueue() + finish().
    // enqueue() launches the three list kernels and the async read-back of the count; finish() blocks until the count
    // is on the host.  Independent work enqueued in between hides the host round trip.
    int enqueue(const uint8_t* act, cudaStream_t s, int per_seg_flags = 0);
    int finish();
    int build(const uint8_t* act, cudaStream_t s);
    cudaEvent_t ready = nullptr;
    bool pending = false;
};

// grid size for a list-walking kernel: enough CTAs for the list, at most `cap`; the list length is rounded up to a
// power of two first so that the launch configuration (and with it the captured iteration graph) changes rarely
inline int seg_grid(int nseg, int segs_per_block, int cap) {
    long long n = 1;
    while (n < nseg) n <<= 1;
    long long b = (n + segs_per_block - 1) / segs_per_block;
    if (b < 1) b = 1;
    return (int)(b < cap ? b : cap);
}
inline long long seg_level(int nseg) {
    long long n = 1;
    while (n < nseg) n <<= 1;
    return n;
}

// two consecutive elements per lane (16-byte accesses for fp64, 8-byte for fp32): a half-warp covers one segment
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// K2 on the active set (same arithmetic as cg_update_xr_kernel).  NCOMP component arrays of `comp_stride` elements
// follow each other in x,r,d,q; a point index p of the list addresses element c*comp_stride + p of component c.  A
// half-warp takes one segment and moves all NCOMP components of it at once (4*NCOMP 16-byte loads in flight per lane).
// Returns this thread's share of r.r (owned rows only when DIST).  No __restrict__/__ldg on the vectors: the same body
// runs inside the persistent whole-iteration kernel, where other CTAs rewrite them between grid barriers.
template <typename T, int NCOMP, bool DIST>
__device__ __forceinline__ double cg_update_xr_seg_body(long long comp_stride, long long npts, const int* __restrict__ seg, int nseg,
                                                        T* x, T* r, const T* d, const T* q, T alpha, const PeerHot& hot) {
    using V = typename Vec2<T>::type;
    const int hl = threadIdx.x & 15;
    const long long hw0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long nhw = ((long long)gridDim.x * blockDim.x) >> 4;
    double acc = 0.0;
    long long j = hw0;
    int sg = j < nseg ? __ldg(seg + j) : 0;
    while (j < nseg) {
        const long long p = (long long)sg * kSegPts + 2 * hl;
        j += nhw;
        sg = j < nseg ? __ldg(seg + j) : 0;           // next segment id requested before this one's data
        if (p >= npts) continue;                      // tail of the last segment (npts is even)
        V xv[NCOMP], rv[NCOMP], dv[NCOMP], qv[NCOMP];
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + p;
            xv[c] = *reinterpret_cast<const V*>(x + e);
            rv[c] = *reinterpret_cast<const V*>(r + e);
            dv[c] = *reinterpret_cast<const V*>(d + e);
            qv[c] = *reinterpret_cast<const V*>(q + e);
        }
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + p;
            bool own = true;
            if (DIST) {
#pragma unroll
                for (int k = 0; k < 6; ++k) own = own && !(e >= hot.hb[k] && e < hot.he[k]);
            }
            xv[c].x = xv[c].x + alpha * dv[c].x; xv[c].y = xv[c].y + alpha * dv[c].y;
            rv[c].x = rv[c].x - alpha * qv[c].x; rv[c].y = rv[c].y - alpha * qv[c].y;
            if (own) acc += (double)rv[c].x * (double)rv[c].x + (double)rv[c].y * (double)rv[c].y;
            *reinterpret_cast<V*>(x + e) = xv[c];
            *reinterpret_cast<V*>(r + e) = rv[c];
        }
    }
    return acc;
}

// bookkeeping after the r.r reduction (thread 0 of the finishing block)
__device__ __forceinline__ void cg_after_rr(CgState* st, double alpha_d, double s) {
    st->alpha = alpha_d;
    if (st->dist) { st->red = s; return; }
    st->delta_old = st->delta;
    st->delta = s;
    st->iter += 1;
    if (s < st->tol2) st->done = 1;
    else if (st->iter >= st->max_iter || !(s == s)) st->done = 2;  // NaN: the reference would spin to max_iter
}

template <typename T, int NCOMP, bool DIST>
__global__ void __launch_bounds__(kVecThreads) cg_update_xr_seg_kernel(long long comp_stride, long long npts,
                                                                       const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                       T* x, T* r, const T* d, const T* q, CgState* st, double* partials, int freeze,
                                                                       PeerInfo* peers, PeerHot hot) {
    if (*(volatile int*)&st->done) return;
    const double alpha_d = st->delta / st->dq;
    const double acc = cg_update_xr_seg_body<T, NCOMP, DIST>(comp_stride, npts, seg, *nseg_p, x, r, d, q, (T)alpha_d, hot);
    grid_sum_finish(acc, partials, &st->counter[1], [=](double s) {
        if (freeze) return;
        cg_after_rr(st, alpha_d, s);
    }, (DIST && !freeze) ? peers : nullptr, 1, false);
}

// K3 on the active set
template <typename T, int NCOMP>
__device__ __forceinline__ void cg_update_d_seg_body(long long comp_stride, long long npts, const int* __restrict__ seg, int nseg,
                                                     T* d, const T* r, T beta) {
    using V = typename Vec2<T>::type;
    const int hl = threadIdx.x & 15;
    const long long hw0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long nhw = ((long long)gridDim.x * blockDim.x) >> 4;
    long long j = hw0;
    int sg = j < nseg ? __ldg(seg + j) : 0;
    while (j < nseg) {
        const long long p = (long long)sg * kSegPts + 2 * hl;
        j += nhw;
        sg = j < nseg ? __ldg(seg + j) : 0;
        if (p >= npts) continue;
        V dv[NCOMP], rv[NCOMP];
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + p;
            dv[c] = *reinterpret_cast<const V*>(d + e);
            rv[c] = *reinterpret_cast<const V*>(r + e);
        }
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + p;
            dv[c].x = rv[c].x + beta * dv[c].x; dv[c].y = rv[c].y + beta * dv[c].y;
            *reinterpret_cast<V*>(d + e) = dv[c];
        }
    }
}

template <typename T, int NCOMP>
__global__ void __launch_bounds__(kVecThreads) cg_update_d_seg_kernel(long long comp_stride, long long npts,
                                                                      const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                      T* d, const T* r, CgState* st) {
    if (*(volatile int*)&st->done) return;
    const double beta_d = st->delta / st->delta_old;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->beta = beta_d;
    cg_update_d_seg_body<T, NCOMP>(comp_stride, npts, seg, *nseg_p, d, r, (T)beta_d);
}

// ---------------------------------------------------------------------------------------------
// Single-reduction CG (Chronopoulos & Gear 1989): the same Krylov iterates as the reference's loop
// (ViscosityCGSolver3D.py:588-610) with BOTH dot products of an iteration taken at one point, so an
// iteration needs one grid-wide (and, multi-GPU, one cross-GPU) reduction instead of two:
//     w = A r ;  gamma = r.r ;  dl = w.r                      (K1s: operator apply fused with both dots)
//     stop if gamma < tol^2      (the reference's test "after each r update", evaluated on the same r)
//     beta = gamma/gamma_old (0 first) ;  alpha = gamma / (dl - beta*gamma/alpha_old)
//     p = r + beta p ;  s = w + beta s  (= A p) ;  x += alpha p ;  r -= alpha s      (K2s: one fused pass)
// p lives in the D vector, s in Q (it IS the reference's q = A d), w in the second direction buffer.
// Words per point and iteration are unchanged (13 + 27 instead of 13 + 18 + 9).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cg_sr_scalars(double gamma, double dl, double gamma_old, double alpha_old, bool first,
                                              double& alpha, double& beta) {
    beta = first ? 0.0 : gamma / gamma_old;
    alpha = first ? gamma / dl : gamma / (dl - beta * gamma / alpha_old);
}

// bookkeeping after the (r.r, w.r) reduction of K1s (thread 0 of the finishing block)
__device__ __forceinline__ void cg_sr_after_dots(CgState* st, double gamma, double dl, int parity = 0) {
    st->sr_parity = parity;
    st->delta = gamma;
    if (gamma < st->tol2) { st->done = 1; return; }
    if (st->iter >= st->max_iter || !(gamma == gamma)) { st->done = 2; return; }   // NaN: the reference would spin to max_iter
    double alpha, beta;
    cg_sr_scalars(gamma, dl, st->delta_old, st->alpha, st->sr_first != 0, alpha, beta);
    st->alpha = alpha; st->beta = beta; st->dq = dl; st->delta_old = gamma; st->sr_first = 0;
}

template <typename T, int NCOMP>
__device__ __forceinline__ void cg_update_sr_seg_body(long long comp_stride, long long npts, const int* __restrict__ seg, int nseg,
                                                      T* x, T* r, T* p, T* sv, const T* w, T alpha, T beta) {
    using V = typename Vec2<T>::type;
    const int hl = threadIdx.x & 15;
    const long long hw0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long nhw = ((long long)gridDim.x * blockDim.x) >> 4;
    long long j = hw0;
    int sg = j < nseg ? __ldg(seg + j) : 0;
    while (j < nseg) {
        const long long pt = (long long)sg * kSegPts + 2 * hl;
        j += nhw;
        sg = j < nseg ? __ldg(seg + j) : 0;
        if (pt >= npts) continue;
        V xv[NCOMP], rv[NCOMP], pv[NCOMP], qv[NCOMP], wv[NCOMP];
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + pt;
            xv[c] = *reinterpret_cast<const V*>(x + e);
            rv[c] = *reinterpret_cast<const V*>(r + e);
            pv[c] = *reinterpret_cast<const V*>(p + e);
            qv[c] = *reinterpret_cast<const V*>(sv + e);
            wv[c] = *reinterpret_cast<const V*>(w + e);
        }
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            const long long e = (long long)c * comp_stride + pt;
            pv[c].x = rv[c].x + beta * pv[c].x; pv[c].y = rv[c].y + beta * pv[c].y;
            qv[c].x = wv[c].x + beta * qv[c].x; qv[c].y = wv[c].y + beta * qv[c].y;
            xv[c].x = xv[c].x + alpha * pv[c].x; xv[c].y = xv[c].y + alpha * pv[c].y;
            rv[c].x = rv[c].x - alpha * qv[c].x; rv[c].y = rv[c].y - alpha * qv[c].y;
            *reinterpret_cast<V*>(p + e) = pv[c];
            *reinterpret_cast<V*>(sv + e) = qv[c];
            *reinterpret_cast<V*>(x + e) = xv[c];
            *reinterpret_cast<V*>(r + e) = rv[c];
        }
    }
}

template <typename T, int NCOMP>
__global__ void __launch_bounds__(kVecThreads) cg_update_sr_seg_kernel(long long comp_stride, long long npts,
                                                                       const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                       T* x, T* r, T* p, T* sv, const T* w /*[2][NCOMP][comp_stride]*/, CgState* st, int freeze) {
    if (*(volatile int*)&st->done) return;
    const double alpha_d = st->alpha, beta_d = st->beta;
    w += (long long)st->sr_parity * NCOMP * comp_stride;           // the buffer the preceding K1s wrote
    cg_update_sr_seg_body<T, NCOMP>(comp_stride, npts, seg, *nseg_p, x, r, p, sv, w, (T)alpha_d, (T)beta_d);
    if (!freeze && blockIdx.x == 0 && threadIdx.x == 0) st->iter += 1;     // read by the next K1s (stream order)
}

// ---------------------------------------------------------------------------------------------
// Grid-wide synchronisation for the persistent whole-iteration kernels (cooperative launch: every
// CTA is resident).  One monotonically increasing arrival counter (reset by the host before each
// launch): barrier k is passed when the counter reaches k*gridDim.x, so the last arriver's atomic
// itself releases everybody — no second flag, no reset on the critical path.
//   grid_sync        plain barrier
//   grid_allreduce   sum of one double over the grid, returned to EVERY thread: each block leaves
//                    its partial (double-buffered by barrier parity), passes the barrier and then
//                    adds up all partials itself in a fixed order, so all blocks obtain the same
//                    bits and keep the CG scalars in registers — no global state on the critical path.
// Arrival is an acq_rel atomic (releases this block's writes — bar.sync before it makes that
// cumulative over the block); waiters spin on an acquire load.
// ---------------------------------------------------------------------------------------------
// (A two-level arrival — 16-CTA groups on separate lines, the last arriver of a group arriving at a top counter — was
// measured and rejected: the second dependent atomic round trip costs more than the serialisation of 148 arrivals on one
// line saves; plain barrier 1.21 -> 1.84 us, reducing barrier 2.72 -> 3.58 us on B200.)
// (Also measured and rejected: a counter-free reduction in which every block posts its partial sums as flag-in-data words
// (4 data bytes + sequence number) and thread t of EVERY block polls block t's words — no atomic, no separate read of the
// partials.  148 blocks x 148 pollers hammer the ~40 L2 lines of the slots: reducing barrier 2.56 -> 4.78 us, and the
// following plain barrier 1.17 -> 1.90 us, on the 256^3 benchmark scene.)
struct GridBar { unsigned int count; unsigned int pad; unsigned long long result_ll[8]; };

__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
    unsigned int r;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int r;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

struct GridSync {
    GridBar* bar;
    unsigned int passed;     // barriers passed so far in this launch (identical in every thread of the grid)
    __device__ __forceinline__ void arrive_and_wait(bool system_scope) {   // thread 0 of the block only
        ++passed;
        const unsigned int target = passed * gridDim.x;
        if (system_scope) fence_acq_rel_sys();                             // stores into a peer GPU need system scope
        if (atom_add_acq_rel_gpu(&bar->count, 1u) + 1u < target) {
            while (ld_acquire_gpu(&bar->count) < target) { }
        }
    }
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) arrive_and_wait(false);
        else ++passed;
        __syncthreads();
    }
};

// every thread of the block gets the block total; fixed order => identical bits in every block that sums the same values
__device__ __forceinline__ double block_sum_all(double v, double* sm /*>=32 doubles*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = lane < nwarps ? sm[lane] : 0.0;
    return warp_sum(t);
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Multi-GPU part of grid_allreduce: `local` is this rank's sum (known to every block).  Warp 0 of block 0 publishes it
// to every rank's mailbox (its own included), collects the N rank sums from the local mailbox, adds them in rank order
// and republishes the total in a local LL-style word pair that the other blocks poll: only N lanes of one warp spin at
// system scope on the mailbox line the NVLink writes land in (every block polling it made the reduction slower with
// every added rank), the other 147 blocks spin on a private line at GPU scope.  Mailbox and result words carry 4 data
// bytes + a 4-byte sequence number, so no separate flag round trip is needed; one acquire fence after the spin makes
// the q halo rows a neighbour stored before publishing visible.  seq = sequence number of this reduction (the same in
// every block; the caller keeps the running count).  Bounded spin: a dead peer yields NaN and P->error = 1.
__device__ __forceinline__ double peer_allreduce_block(double local, PeerInfo* P, int kind, unsigned int seq, double* sm_bcast,
                                                       unsigned long long* result_ll /*[2 kinds][2 parities][2 words], local*/) {
    const int lane = threadIdx.x & 31;
    unsigned long long* res = result_ll + ((kind & 1) * 2 + (int)(seq & 1u)) * 2;     // per kind and parity: tags of the two kinds coincide
    if (threadIdx.x < 32) {
        double sum = 0.0;
        if (blockIdx.x == 0) {
            const int n = P->nranks, me = P->rank;
            const int slot = ((kind * 2 + (int)(seq & 1u)) * kMaxRanks);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
            const unsigned long long w0 = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
            const unsigned long long w1 = ((unsigned long long)seq << 32) | (bits >> 32);
            fence_acq_rel_sys();                       // everything this rank did (peer stores, halo reads) is ordered before
            if (lane < n) {
                unsigned long long* dst = P->mbox[lane] + (slot + me) * 2;
                st_relaxed_sys_u64(dst, w0);
                st_relaxed_sys_u64(dst + 1, w1);
            }
            double v = 0.0;
            bool ok = true;
            if (lane < n) {
                const unsigned long long* src = P->mbox[me] + (slot + lane) * 2;
                unsigned long long a, b;
                const long long t0 = clock64();
                for (;;) {
                    a = ld_relaxed_sys_u64(src);
                    b = ld_relaxed_sys_u64(src + 1);
                    if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
                    if (clock64() - t0 > 20000000000LL) { ok = false; break; }     // ~10 s: a peer died; bail out instead of hanging
                }
                v = __longlong_as_double((long long)(((b & 0xffffffffull) << 32) | (a & 0xffffffffull)));
            }
            fence_acq_rel_sys();                       // acquire what the peers stored before publishing; release it to the local blocks
            if (!__all_sync(0xffffffffu, ok)) {
                if (lane == 0) P->error = 1;
                sum = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                for (int r = 0; r < n; ++r) sum += __shfl_sync(0xffffffffu, v, r);        // fixed rank order
            }
            if (lane == 0) {
                const unsigned long long sb = (unsigned long long)__double_as_longlong(sum);
                st_relaxed_gpu_u64(res, ((unsigned long long)seq << 32) | (sb & 0xffffffffull));
                st_relaxed_gpu_u64(res + 1, ((unsigned long long)seq << 32) | (sb >> 32));
            }
        } else if (lane == 0) {
            unsigned long long a, b;
            const long long t0 = clock64();
            for (;;) {
                a = ld_relaxed_gpu_u64(res);
                b = ld_relaxed_gpu_u64(res + 1);
                if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
                if (clock64() - t0 > 40000000000LL) { a = 0; b = 0x7ff80000ull; break; }   // block 0 gave up: NaN
            }
            fence_acq_rel_gpu();
            sum = __longlong_as_double((long long)(((b & 0xffffffffull) << 32) | (a & 0xffffffffull)));
        }
        if (lane == 0) *sm_bcast = sum;
    }
    __syncthreads();
    const double out = *sm_bcast;
    __syncthreads();
    return out;
}

// seq: when `peers` is given, the sequence number of this cross-GPU reduction (see peer_allreduce_block)
__device__ __forceinline__ double grid_allreduce(double v, double* partials /*2*gridDim.x*/, GridSync& gs,
                                                 PeerInfo* peers = nullptr, int kind = 0, bool wrote_peer = false, unsigned int seq = 0) {
    __shared__ double sm[32];
    __shared__ double s_glob;
    const unsigned int nblocks = gridDim.x;
    double* slot = partials + (size_t)(gs.passed & 1u) * nblocks;
    v = block_sum(v, sm);
    if (threadIdx.x == 0) {
        slot[blockIdx.x] = v;
        gs.arrive_and_wait(wrote_peer);
    } else {
        ++gs.passed;
    }
    __syncthreads();
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) s += __ldcg(slot + i);
    s = block_sum_all(s, sm);
    if (peers) s = peer_allreduce_block(s, peers, kind, seq, &s_glob, gs.bar->result_ll);
    return s;
}

// Two-scalar variant of peer_allreduce_block / grid_allreduce (single-reduction CG): both values cross NVLink in one
// mailbox round (kind 0 and kind 1 slots, sequence numbers seq0 / seq1) and come back through the two local LL word pairs.
__device__ __forceinline__ void peer_allreduce2_block(double& a, double& b, PeerInfo* P, unsigned int seq0, unsigned int seq1,
                                                      double* sm_bcast /*[2]*/, unsigned long long* result_ll) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const int kind = lane >> 4, rk = lane & 15;
        const unsigned int seq = kind ? seq1 : seq0;
        unsigned long long* res = result_ll + (kind * 2 + (int)(seq & 1u)) * 2;
        double sum = 0.0;
        if (blockIdx.x == 0) {
            const int n = P->nranks, me = P->rank;
            const int slot = ((kind * 2 + (int)(seq & 1u)) * kMaxRanks);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(kind ? b : a);
            const unsigned long long w0 = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
            const unsigned long long w1 = ((unsigned long long)seq << 32) | (bits >> 32);
            fence_acq_rel_sys();
            if (rk < n) {
                unsigned long long* dst = P->mbox[rk] + (slot + me) * 2;
                st_relaxed_sys_u64(dst, w0);
                st_relaxed_sys_u64(dst + 1, w1);
            }
            double v = 0.0;
            bool ok = true;
            if (rk < n) {
                const unsigned long long* src = P->mbox[me] + (slot + rk) * 2;
                unsigned long long x, y;
                const long long t0 = clock64();
                for (;;) {
                    x = ld_relaxed_sys_u64(src);
                    y = ld_relaxed_sys_u64(src + 1);
                    if ((unsigned int)(x >> 32) == seq && (unsigned int)(y >> 32) == seq) break;
                    if (clock64() - t0 > 20000000000LL) { ok = false; break; }
                }
                v = __longlong_as_double((long long)(((y & 0xffffffffull) << 32) | (x & 0xffffffffull)));
            }
            fence_acq_rel_sys();
            const bool all_ok = __all_sync(0xffffffffu, ok);
            double sa = 0.0, sb = 0.0;
            for (int r = 0; r < n; ++r) { sa += __shfl_sync(0xffffffffu, v, r); sb += __shfl_sync(0xffffffffu, v, 16 + r); }
            if (!all_ok) {
                if (lane == 0) P->error = 1;
                sa = sb = __longlong_as_double(0x7ff8000000000000LL);
            }
            sum = kind ? sb : sa;
            if (rk == 0) {
                const unsigned long long sbits = (unsigned long long)__double_as_longlong(sum);
                st_relaxed_gpu_u64(res, ((unsigned long long)seq << 32) | (sbits & 0xffffffffull));
                st_relaxed_gpu_u64(res + 1, ((unsigned long long)seq << 32) | (sbits >> 32));
            }
        } else if (rk == 0) {
            unsigned long long x, y;
            const long long t0 = clock64();
            for (;;) {
                x = ld_relaxed_gpu_u64(res);
                y = ld_relaxed_gpu_u64(res + 1);
                if ((unsigned int)(x >> 32) == seq && (unsigned int)(y >> 32) == seq) break;
                if (clock64() - t0 > 40000000000LL) { x = 0; y = 0x7ff80000ull; break; }
            }
            fence_acq_rel_gpu();
            sum = __longlong_as_double((long long)(((y & 0xffffffffull) << 32) | (x & 0xffffffffull)));
        }
        if (rk == 0) sm_bcast[kind] = sum;
    }
    __syncthreads();
    a = sm_bcast[0];
    b = sm_bcast[1];
    __syncthreads();
}

// two values through ONE pair of block barriers; totals valid in thread 0 (block_sum2) / in every thread (block_sum_all2)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sm /*>=64 doubles*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { sm[warp] = a; sm[32 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        a = warp_sum(lane < nwarps ? sm[lane] : 0.0);
        b = warp_sum(lane < nwarps ? sm[32 + lane] : 0.0);
    }
}
__device__ __forceinline__ void block_sum_all2(double& a, double& b, double* sm /*>=64 doubles*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { sm[warp] = a; sm[32 + warp] = b; }
    __syncthreads();
    a = warp_sum(lane < nwarps ? sm[lane] : 0.0);
    b = warp_sum(lane < nwarps ? sm[32 + lane] : 0.0);
}

// partials: 4 * gridDim.x doubles (two values, double-buffered by barrier parity)
__device__ __forceinline__ void grid_allreduce2(double& v1, double& v2, double* partials, GridSync& gs,
                                                PeerInfo* peers = nullptr, bool wrote_peer = false, unsigned int seq0 = 0, unsigned int seq1 = 0) {
    __shared__ double sm[64];
    __shared__ double s_glob[2];
    const unsigned int nblocks = gridDim.x;
    double* slot = partials + (size_t)(gs.passed & 1u) * 2 * nblocks;
    block_sum2(v1, v2, sm);
    if (threadIdx.x == 0) {
        slot[blockIdx.x] = v1;
        slot[nblocks + blockIdx.x] = v2;
        gs.arrive_and_wait(wrote_peer);
    } else {
        ++gs.passed;
    }
    __syncthreads();
    double s1 = 0.0, s2 = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += blockDim.x) { s1 += __ldcg(slot + i); s2 += __ldcg(slot + nblocks + i); }
    block_sum_all2(s1, s2, sm);
    if (peers) peer_allreduce2_block(s1, s2, peers, seq0, seq1, s_glob, gs.bar->result_ll);
    v1 = s1; v2 = s2;
}

constexpr int kPersistThreads = 512;      // one CTA per SM (128 registers per thread for the operator body)

constexpr int kSegsPerVecBlock = kVecThreads / 16;   // one half-warp per segment

template <typename T, int NCOMP>
int cg_launch_update_xr_seg(long long comp_stride, long long npts, const SegList& sl, T* x, T* r, const T* d, const T* q,
                            CgState* st, double* partials, cudaStream_t s, int freeze = 0, PeerInfo* peers = nullptr,
                            const PeerHot* hotp = nullptr) {
    PeerHot hot;
    memset(&hot, 0, sizeof(hot));
    if (hotp) hot = *hotp;
    const int grid = seg_grid(sl.nseg, kSegsPerVecBlock, kVecGrid);
    if (peers) cg_update_xr_seg_kernel<T, NCOMP, true><<<grid, kVecThreads, 0, s>>>(comp_stride, npts, sl.list, sl.nseg_dev, x, r, d, q, st, partials, freeze, peers, hot);
    else cg_update_xr_seg_kernel<T, NCOMP, false><<<grid, kVecThreads, 0, s>>>(comp_stride, npts, sl.list, sl.nseg_dev, x, r, d, q, st, partials, freeze, nullptr, hot);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

template <typename T, int NCOMP>
int cg_launch_update_d_seg(long long comp_stride, long long npts, const SegList& sl, T* d, const T* r, CgState* st, cudaStream_t s) {
    const int grid = seg_grid(sl.nseg, kSegsPerVecBlock, kVecGrid);
    cg_update_d_seg_kernel<T, NCOMP><<<grid, kVecThreads, 0, s>>>(comp_stride, npts, sl.list, sl.nseg_dev, d, r, st);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
inline int vec_grid(long long n, int vec) {
    long long b = (n / vec + kVecThreads - 1) / kVecThreads;
    if (b < 1) b = 1;
    return (int)(b < kVecGrid ? b : kVecGrid);
}

template <typename T>
int cg_launch_update_xr(long long n, T* x, T* r, const T* d, const T* q, CgState* st, double* partials, cudaStream_t s, int freeze = 0,
                        PeerInfo* peers = nullptr, const PeerHot* hotp = nullptr) {
    PeerHot hot;
    memset(&hot, 0, sizeof(hot));
    if (hotp) hot = *hotp;
    constexpr int N = Vec16<T>::N;
    const bool vec = aligned16(x) && aligned16(r) && aligned16(d) && aligned16(q);
    if (peers) {
        if (vec) cg_update_xr_kernel<T, N, true><<<vec_grid(n, N), kVecThreads, 0, s>>>(n, x, r, d, q, st, partials, freeze, peers, hot);
        else cg_update_xr_kernel<T, 1, true><<<vec_grid(n, 1), kVecThreads, 0, s>>>(n, x, r, d, q, st, partials, freeze, peers, hot);
    } else {
        if (vec) cg_update_xr_kernel<T, N, false><<<vec_grid(n, N), kVecThreads, 0, s>>>(n, x, r, d, q, st, partials, freeze, nullptr, hot);
        else cg_update_xr_kernel<T, 1, false><<<vec_grid(n, 1), kVecThreads, 0, s>>>(n, x, r, d, q, st, partials, freeze, nullptr, hot);
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

template <typename T>
int cg_launch_update_d(long long n, T* d, const T* r, CgState* st, cudaStream_t s) {
    constexpr int N = Vec16<T>::N;
    if (aligned16(d) && aligned16(r))
        cg_update_d_kernel<T, N><<<vec_grid(n, N), kVecThreads, 0, s>>>(n, d, r, st);
    else
        cg_update_d_kernel<T, 1><<<vec_grid(n, 1), kVecThreads, 0, s>>>(n, d, r, st);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

template <typename T>
int cg_launch_residual_init(long long n, const T* b, const T* q, T* d, T* r, CgState* st, double* partials, cudaStream_t s,
                            PeerInfo* peers = nullptr) {
    constexpr int N = Vec16<T>::N;
    if (aligned16(b) && aligned16(q) && aligned16(d) && aligned16(r))
        cg_residual_init_kernel<T, N><<<vec_grid(n, N), kVecThreads, 0, s>>>(n, b, q, d, r, st, partials, peers);
    else
        cg_residual_init_kernel<T, 1><<<vec_grid(n, 1), kVecThreads, 0, s>>>(n, b, q, d, r, st, partials, peers);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

__global__ void cg_state_init_kernel(CgState* st, double tol2, long long max_iter, int dist = 0);
__global__ void cg_state_unlimit_kernel(CgState* st);
// multi-GPU: bookkeeping that the reducing kernels skip, run after `red` has been allreduced.
// which = 0: start of a solve (delta0);  which = 1: after K2 (delta', iteration count, convergence flags)
__global__ void cg_finish_kernel(CgState* st, int which);

// host-side CG control shared by all solvers ------------------------------------------------
struct CgHost {
    CgState* st_dev = nullptr;      // in workspace
    double* partials_dev = nullptr; // in workspace, >= max grid size of any reducing kernel
    CgState* st_pinned = nullptr;   // cudaHostAlloc, 2 slots
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int init();
    void destroy();
};

// A batch of CG iterations captured once into a CUDA graph (3 kernel nodes per iteration) and replayed: removes the
// per-launch gaps that dominate when an iteration is only ~100 us of work (multi-GPU slabs, 64^3, the 2-D configs).
// Capture happens on an internal non-blocking stream because the caller's stream is usually the legacy default
// stream, which cannot be captured; the instantiated graph is then launched into the caller's stream.
struct IterGraph {
    cudaGraphExec_t exec = nullptr;
    cudaStream_t cap = nullptr;
    double key = 0.0;
    long long key2 = 0;     // launch-configuration level (active-set size class); a change re-captures the graph
    int iters = 0;
    bool valid = false;
    static bool enabled();

    template <class EnqueueOne>
    int ensure(double key_, long long key2_, int iters_, EnqueueOne one) {
        if (valid && key == key_ && key2 == key2_ && iters == iters_) return FS_OK;
        if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
        valid = false;
        if (!cap) FS_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        const long long launches_before = g_launches;
        FS_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        int status = FS_OK;
        for (int k = 0; k < iters_ && status >= 0; ++k) status = one(cap);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(cap, &graph);
        per_launch = g_launches - launches_before;
        g_launches = launches_before;
        if (status < 0) { if (graph) cudaGraphDestroy(graph); return status; }
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { exec = nullptr; return fail(FS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
        key = key_; key2 = key2_; iters = iters_; valid = true;
        return FS_OK;
    }
    int launch(cudaStream_t s) {
        FS_CUDA(cudaGraphLaunch(exec, s));
        g_launches += per_launch;
        return FS_OK;
    }
    void destroy() {
        if (exec) cudaGraphExecDestroy(exec);
        if (cap) cudaStreamDestroy(cap);
        exec = nullptr; cap = nullptr; valid = false;
    }
    long long per_launch = 0;
};

constexpr int kCgBatch = 16;
constexpr int kCgBatchPersistent = 64;   // iterations per launch of a persistent whole-iteration kernel (it stops early when done)
constexpr int kCgBatchResident = 256;    // ... of the shared-memory resident kernels: every launch first loads / finally stores the resident state

// Enqueue `n` iterations: whole batches of kCgBatch through the graph, the remainder launch by launch.
template <class EnqueueOne>
int cg_enqueue_iterations(IterGraph& g, bool use_graph, double key, long long n, EnqueueOne one, cudaStream_t s, long long key2 = 0) {
    if (use_graph && IterGraph::enabled() && n >= kCgBatch) {
        FS_TRY(g.ensure(key, key2, kCgBatch, one));
        while (n >= kCgBatch) { FS_TRY(g.launch(s)); n -= kCgBatch; }
    }
    for (long long k = 0; k < n; ++k) FS_TRY(one(s));
    return FS_OK;
}

// Runs `batch_fn(stream, nb)` (enqueue nb iterations of K1,K2,K3) until the device flags completion.
// Blocks the calling thread; fills stats.  Returns FS_OK / FS_NOT_CONVERGED / <0.
template <class BatchFn>
int cg_drive(CgHost& c, BatchFn batch_fn, long long max_iter, fs_cg_stats* stats, cudaStream_t s, int batch = kCgBatch) {
    // state already initialised and residual_init enqueued by the caller
    int slot = 0;
    bool pending[2] = {false, false};
    long long enq = 0;
    int status = FS_OK;
    for (;;) {
        long long nb = batch;
        if (enq + nb > max_iter) nb = max_iter - enq;
        FS_TRY(batch_fn(s, nb));
        enq += nb;
        FS_CUDA(cudaMemcpyAsync(&c.st_pinned[slot], c.st_dev, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        FS_CUDA(cudaEventRecord(c.ev[slot], s));
        pending[slot] = true;
        // look at the previous batch while this one runs
        int prev = slot ^ 1;
        bool finished = false;
        if (pending[prev]) {
            FS_CUDA(cudaEventSynchronize(c.ev[prev]));
            pending[prev] = false;
            if (c.st_pinned[prev].done) finished = true;
        }
        if (finished || enq >= max_iter) {
            FS_CUDA(cudaEventSynchronize(c.ev[slot]));
            pending[slot] = false;
            const CgState& h = c.st_pinned[slot];
            if (stats) {
                stats->iterations = h.iter;
                stats->delta = h.delta;
                stats->alpha = h.alpha;
                stats->beta = h.beta;
                stats->delta0 = h.delta0;
                stats->converged = (h.done == 1);
                stats->reserved = 0;
            }
            status = (h.done == 1) ? FS_OK : FS_NOT_CONVERGED;
            break;
        }
        slot ^= 1;
    }
    return status;
}

}  // namespace fs
