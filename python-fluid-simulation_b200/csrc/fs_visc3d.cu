// 3-D variational viscosity solver on a packed MAC lattice (sm_100a).
//
// Replaces ViscosityCGSolver3D.py (kernels :8-470, launchers :472-530, solve :566-613).
//
// HBM layout (all SoA, z contiguous).  One padded lattice  X=nx+1, Y=ny+1, Zp=roundup(nz+1,4)  is
// shared by every array so a single flat index i=(x*Y+y)*Zp+z addresses the u,v,w faces, the cell
// centre and the three low edges that belong to lattice point (x,y,z); neighbours are i+-1, i+-Zp,
// i+-Y*Zp for every array alike.
//   coef[0..2]  Vu,Vv,Vw   face liquid-volume fraction; NaN on rows the operator never computes
//                          (solid faces, the frozen boundary layer, lattice padding)
//   coef[3]     Vc         cell-centre volume
//   coef[4..6]  Exy,Exz,Eyz edge-centred volumes
//   mask[0..2]  fluid flag of each face (sphi >= 0), bytes; used by the masked apply / RHS /
//                          extrapolation / write-back only — never inside the CG loop
//   vec[v][c]   X,R,D,Q,B  solver vectors, 3 components each, NL elements per component
#include <type_traits>

#include "fs_comm.cuh"
#include "fs_common.cuh"
#include "fs_visc_rows.cuh"

namespace fs {

struct Lat3 {
    int nx, ny, nz;
    int X, Y, Zp;
    long long sx, sy, NL;
    int u_xhi;   // last computed x-plane of u rows: nx-1, or nx-2 when a higher slab owns plane nx-1 (multi-GPU)
};

template <typename T> struct Visc3Dev {
    Lat3 L;
    const T* coef[7];
    const uint8_t* mask[3];
    const uint8_t* act;      // per lattice point: bit c set <=> row c is computed (same information as the NaN tags, 1 byte)
};

__device__ __forceinline__ void lat_decode(const Lat3& L, long long i, int& x, int& y, int& z) {
    z = (int)(i % L.Zp);
    long long t = i / L.Zp;
    y = (int)(t % L.Y);
    x = (int)(t / L.Y);
}

// shape of component c's MAC array
__device__ __forceinline__ void comp_shape(const Lat3& L, int c, int& s0, int& s1, int& s2) {
    s0 = L.nx + (c == 0);
    s1 = L.ny + (c == 1);
    s2 = L.nz + (c == 2);
}

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// pack: fine-grid (2n+1)^3 fp64 sphi / lvol  ->  lattice coefficient planes + face masks
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_pack_kernel(Lat3 L, const double* __restrict__ sphi, const double* __restrict__ lvol,
                                                               double vol_norm, T* __restrict__ coef /*[7][NL]*/, uint8_t* __restrict__ mask /*[3][NL]*/,
                                                               uint8_t* __restrict__ act /*[NL + 32]*/) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    int x, y, z;
    lat_decode(L, i, x, y, z);
    const long long fz = 1, fy = 2LL * L.nz + 1, fx = fy * (2LL * L.ny + 1);
    const long long f0 = 2LL * x * fx + 2LL * y * fy + 2LL * z * fz;  // fine node (2x,2y,2z)
    const bool ix = x < L.nx, iy = y < L.ny, iz = z < L.nz, inz = z <= L.nz;
    const T nan = (T)__longlong_as_double(0x7ff8000000000000LL);
    auto vol = [&](long long off) { return (T)(lvol[f0 + off] / vol_norm); };
    unsigned int abits = 0;
    // faces: fine parities (0,1,1) (1,0,1) (1,1,0)
    {
        const bool in = iy && iz;
        bool fluid = false;
        T v = nan;
        if (in) {
            fluid = sphi[f0 + fy + fz] >= 0.0;
            const bool interior = x >= 1 && x <= L.u_xhi && y >= 1 && y <= L.ny - 2 && z >= 1 && z <= L.nz - 2;
            if (fluid && interior) { v = vol(fy + fz); if (v == v) abits |= 1u; }
        }
        coef[0 * L.NL + i] = v;
        mask[0 * L.NL + i] = fluid;
    }
    {
        const bool in = ix && iz;
        bool fluid = false;
        T v = nan;
        if (in) {
            fluid = sphi[f0 + fx + fz] >= 0.0;
            const bool interior = x >= 1 && x <= L.nx - 2 && y >= 1 && y <= L.ny - 1 && z >= 1 && z <= L.nz - 2;
            if (fluid && interior) { v = vol(fx + fz); if (v == v) abits |= 2u; }
        }
        coef[1 * L.NL + i] = v;
        mask[1 * L.NL + i] = fluid;
    }
    {
        const bool in = ix && iy && inz;
        bool fluid = false;
        T v = nan;
        if (in) {
            fluid = sphi[f0 + fx + fy] >= 0.0;
            const bool interior = x >= 1 && x <= L.nx - 2 && y >= 1 && y <= L.ny - 2 && z >= 1 && z <= L.nz - 1;
            if (fluid && interior) { v = vol(fx + fy); if (v == v) abits |= 4u; }
        }
        coef[2 * L.NL + i] = v;
        mask[2 * L.NL + i] = fluid;
    }
    coef[3 * L.NL + i] = (ix && iy && iz) ? vol(fx + fy + fz) : T(0);   // cell centre (1,1,1)
    coef[4 * L.NL + i] = (iz) ? vol(fz) : T(0);                         // Exy: (0,0,1)
    coef[5 * L.NL + i] = (iy && inz) ? vol(fy) : T(0);                  // Exz: (0,1,0)
    coef[6 * L.NL + i] = (ix && inz) ? vol(fx) : T(0);                  // Eyz: (1,0,0)
    act[i] = (uint8_t)abits;
}

// ---------------------------------------------------------------------------------------------
// load / store between the caller's dense MAC arrays and lattice vectors
// ---------------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(kThreads) visc3d_load_kernel(Lat3 L, const S* __restrict__ a0, const S* __restrict__ a1, const S* __restrict__ a2,
                                                               T* __restrict__ vec /*[3][NL]*/) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    int x, y, z;
    lat_decode(L, i, x, y, z);
    const S* src[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int s0, s1, s2;
        comp_shape(L, c, s0, s1, s2);
        T v = T(0);
        if (x < s0 && y < s1 && z < s2) v = (T)src[c][((long long)x * s1 + y) * s2 + z];
        vec[c * L.NL + i] = v;
    }
}

template <typename T, typename S>
__global__ void __launch_bounds__(kThreads) visc3d_store_kernel(Lat3 L, const T* __restrict__ vec, const uint8_t* __restrict__ mask,
                                                                S* __restrict__ a0, S* __restrict__ a1, S* __restrict__ a2, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    int x, y, z;
    lat_decode(L, i, x, y, z);
    S* dst[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int s0, s1, s2;
        comp_shape(L, c, s0, s1, s2);
        if (!(x < s0 && y < s1 && z < s2)) continue;
        bool w;
        if (mode == FS_STORE_ALL) w = true;
        else if (mode == FS_STORE_INTERIOR) w = x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2 && z >= 1 && z <= s2 - 2;
        else  // apply_viscosity_kernel :461-470 — one index range 1..g-1 for all three components
            w = x >= 1 && x <= L.nx - 1 && y >= 1 && y <= L.ny - 1 && z >= 1 && z <= L.nz - 1 && mask[c * L.NL + i];
        if (w) dst[c][((long long)x * s1 + y) * s2 + z] = (S)vec[c * L.NL + i];
    }
}

// ---------------------------------------------------------------------------------------------
// extrapolation sweep (extrapolate_kernel :8-39), all three components per launch, ping-pong
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_extrapolate_kernel(Lat3 L, const T* __restrict__ vin, const uint8_t* __restrict__ valin,
                                                                      T* __restrict__ vout, uint8_t* __restrict__ valout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    int x, y, z;
    lat_decode(L, i, x, y, z);
    const long long st[3] = {L.sx, L.sy, 1};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int s0, s1, s2;
        comp_shape(L, c, s0, s1, s2);
        const T* v = vin + c * L.NL;
        const uint8_t* va = valin + c * L.NL;
        T out = v[i];
        uint8_t ov = va[i];
        const bool interior = x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2 && z >= 1 && z <= s2 - 2;
        if (interior && !ov) {
            T val = T(0);
            int count = 0;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {   // +x,-x,+y,-y,+z,-z  (:19-36)
                if (va[i + st[ax]]) { val += v[i + st[ax]]; ++count; }
                if (va[i - st[ax]]) { val += v[i - st[ax]]; ++count; }
            }
            if (count > 0) { out = val / (T)count; ov = 1; }
        }
        vout[c * L.NL + i] = out;
        valout[c * L.NL + i] = ov;
    }
}

// ---------------------------------------------------------------------------------------------
// masked apply / RHS (used once per solve and by the module-level matvecmul / initialize_solver)
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) visc3d_general_kernel(Visc3Dev<T> P, T s, T s2, const T* __restrict__ src, T* __restrict__ dst) {
    const Lat3& L = P.L;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    int x, y, z;
    lat_decode(L, i, x, y, z);
    const long long st[3] = {L.sx, L.sy, 1};
    const long long NL = L.NL;
    const uint8_t* const* mask = P.mask;
    // neighbour value under the mode's mask: apply keeps fluid neighbours, RHS keeps solid ones
    auto nb = [&](int comp, long long j) -> T {
        const bool fluid = mask[comp][j] != 0;
        const bool keep = (MODE == ROW_APPLY) ? fluid : !fluid;
        return keep ? src[comp * NL + j] : T(0);
    };
    auto row = [&](auto Atag) {
        constexpr int A = decltype(Atag)::value;
        int s0, s1, s2_;
        comp_shape(L, A, s0, s1, s2_);
        const int xhi = (A == 0) ? L.u_xhi : s0 - 2;
        const bool interior = x >= 1 && x <= xhi && y >= 1 && y <= s1 - 2 && z >= 1 && z <= s2_ - 2;
        if (!interior) return;                    // boundary layer (and rows owned by another slab): never written (:251)
        T out = T(0);
        if (mask[A][i]) {                         // solid rows -> 0 (:255-258)
            const T center = P.coef[A][i];
            const T own = src[A * NL + i];
            out = visc_row<T, 3, A, true, MODE>(P.coef, i, st, center, own, s, s2, nb);
        }
        dst[A * NL + i] = out;
    };
    row(std::integral_constant<int, 0>{});
    row(std::integral_constant<int, 1>{});
    row(std::integral_constant<int, 2>{});
}

// ---------------------------------------------------------------------------------------------
// K1: CG-loop apply fused with d.q.  Inside the loop d is exactly zero on every row that is not
// computed (solid / boundary / padding / other slab), so neighbour masks are not needed (SURVEY A-1)
// and the row flag rides in the NaN tag of the face volume.
//
// Persistent grid (kK1BlocksPerSM CTAs per SM, grid-stride over the lattice, z contiguous across the
// warp).  Per lattice point: the three NaN tags are loaded first; a warp whose 32 points carry no
// computed row skips the body (the reference's `if sphi < 0: return`, made warp-uniform).  Otherwise
// the body is branch-free — all 27 neighbour values and 16 coefficients are requested before the first
// use, one memory-latency period per point instead of one per row — and rows that are not computed
// select 0.  Loads of discarded lanes may fall outside the lattice; the workspace carries guard bands
// of one plane + one row + one element around the coefficient and vector regions for exactly that.
// One block reduction at the very end (fixed-order, deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int kK1Threads = 256;
// CTAs per SM: the fp64 body keeps 43 loaded values (86 registers) in flight, so it gets 128 registers per thread
// (2 CTAs/SM); with an 80-register cap (3 CTAs/SM) ptxas split the loads into dependent phases and the dense-scene
// K1 ran at 0.556 ms instead of 0.357 ms.  fp32 needs half the registers.
template <typename T> struct K1Occ { static constexpr int value = 2; };
template <> struct K1Occ<float> { static constexpr int value = 3; };

template <typename T, bool DIST>
__global__ void __launch_bounds__(kK1Threads, K1Occ<T>::value) visc3d_apply_dot_kernel(Visc3Dev<T> P, T s, T s2, const T* __restrict__ d, T* __restrict__ q,
                                                                                       CgState* st_, double* partials, PeerInfo* peers, PeerHot hot) {
    if (*(volatile int*)&st_->done) return;
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const long long st[3] = {L.sx, L.sy, 1};
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long NLw = (NL + 31) & ~31LL;       // whole warps take part in every trip (ballot below)
    const T nan = (T)__longlong_as_double(0x7ff8000000000000LL);
    double acc = 0.0;
    bool wrote_peer = false;
    auto nb = [&](int comp, long long j) -> T { return __ldg(d + comp * NL + j); };
    // The activity byte of the NEXT trip is requested before this trip's work: in solid regions a trip is a single 1-byte
    // load, and without the prefetch each warp would have one memory round trip in flight.
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int a_n = i < NL ? __ldg(P.act + i) : 0u;
    for (; i < NLw; i += stride) {
        const bool in = i < NL;
        const unsigned int a = a_n;
        {
            const long long i2 = i + stride;
            a_n = i2 < NL ? __ldg(P.act + i2) : 0u;
        }
        const bool au = a & 1u, av = a & 2u, aw = a & 4u;
        T ou = T(0), ov = T(0), ow = T(0);
        if (__any_sync(0xffffffffu, a != 0u)) {
            const long long j = in ? i : (NL - 1);
            const T cu = __ldg(P.coef[0] + j), cv = __ldg(P.coef[1] + j), cw = __ldg(P.coef[2] + j);
            const T du = nb(0, j), dv = nb(1, j), dw = nb(2, j);
            const T ru = visc_row<T, 3, 0, false, ROW_APPLY>(P.coef, j, st, cu, du, s, s2, nb);
            const T rv = visc_row<T, 3, 1, false, ROW_APPLY>(P.coef, j, st, cv, dv, s, s2, nb);
            const T rw = visc_row<T, 3, 2, false, ROW_APPLY>(P.coef, j, st, cw, dw, s, s2, nb);
            if (au) { ou = ru; acc += (double)du * (double)ru; }
            if (av) { ov = rv; acc += (double)dv * (double)rv; }
            if (aw) { ow = rw; acc += (double)dw * (double)rw; }
        }
        // multi-GPU: the halo planes of q (0 and X-2) are written by the NEIGHBOURS' K1 over NVLink — never by this rank
        const bool halo_plane = DIST && ((hot.has_lo && i < L.sx) ||
                                         (hot.has_hi && i >= (long long)(L.X - 2) * L.sx && i < (long long)(L.X - 1) * L.sx));
        // Rows that are not computed hold q == 0 since the masked apply at the start of the solve (which writes 0 on every
        // interior non-fluid row) and are never written afterwards, so only computed rows are stored — in solid regions
        // K1 therefore moves one byte per lattice point.
        if (in && !halo_plane && a != 0u) {
            if (au) q[i] = ou;
            if (av) q[NL + i] = ov;
            if (aw) q[2 * NL + i] = ow;
            if (DIST) {
                // multi-GPU: my first / last owned planes are the neighbours' halo planes of q — store them straight
                // into the peers' memory over NVLink; the all-reduce in this kernel's tail publishes them.
                const long long lo0 = L.sx, hi0 = (long long)(L.X - 3) * L.sx;
                if (hot.has_lo && i >= lo0 && i < lo0 + L.sx) {
                    const long long o = i - lo0;
                    if (au) reinterpret_cast<T*>(hot.q_lo[0])[o] = ou;
                    if (av) reinterpret_cast<T*>(hot.q_lo[1])[o] = ov;
                    if (aw) reinterpret_cast<T*>(hot.q_lo[2])[o] = ow;
                    wrote_peer = true;
                }
                if (hot.has_hi && i >= hi0 && i < hi0 + L.sx) {
                    const long long o = i - hi0;
                    if (au) reinterpret_cast<T*>(hot.q_hi[0])[o] = ou;
                    if (av) reinterpret_cast<T*>(hot.q_hi[1])[o] = ov;
                    if (aw) reinterpret_cast<T*>(hot.q_hi[2])[o] = ow;
                    wrote_peer = true;
                }
            }
        }
    }
    const bool block_wrote_peer = DIST ? (__syncthreads_or(wrote_peer ? 1 : 0) != 0) : false;
    grid_sum_finish(acc, partials, &st_->counter[0], [=](double sum) { st_->dq = sum; }, DIST ? peers : nullptr, 0, block_wrote_peer);
}

}  // namespace fs

// =================================================================================================
// host side / C ABI
// =================================================================================================
using namespace fs;

struct fs_visc3d {
    Lat3 L;
    int dtype;
    size_t esz;
    char* ws;
    size_t ws_bytes;
    char* coef;      // [7][NL] T
    char* vecs;      // [5][3][NL] T
    uint8_t* mask;   // [3][NL]
    uint8_t* valid;  // [2][3][NL]
    uint8_t* act;    // [NL] computed-row bits
    double* partials;
    CgState* st;
    CgHost cg;
    int grid_pts;    // blocks for one-thread-per-lattice-point kernels
    bool packed;
    fs_comm* comm;   // multi-GPU: this handle is one x-slab (extended by one cell towards each neighbour)
    int has_lo, has_hi;
    IterGraph graph; // captured batch of iterations (single GPU and fused multi-GPU transport)
    PeerHot hot;     // by-value copy of the per-point fields of `peers`
    PeerInfo* peers; // device copy; non-null = collectives fused into K1/K2 over peer memory (NVLink), else NCCL per iteration
};

static Lat3 make_lat3(int nx, int ny, int nz) {
    Lat3 L;
    L.nx = nx; L.ny = ny; L.nz = nz;
    L.X = nx + 1; L.Y = ny + 1; L.Zp = (nz + 1 + 3) / 4 * 4;
    L.sy = L.Zp; L.sx = (long long)L.Y * L.Zp; L.NL = L.sx * L.X;
    L.u_xhi = nx - 1;
    return L;
}

struct Visc3Layout { size_t coef, vecs, mask, valid, act, partials, st, total; int grid_pts; };

static Visc3Layout visc3_layout(const Lat3& L, size_t esz) {
    Visc3Layout o;
    size_t p = 0;
    o.grid_pts = (int)((L.NL + kThreads - 1) / kThreads);
    // guard bands (one plane + one row + one element) so that the branch-free K1 may issue neighbour loads for
    // lanes whose rows are discarded, even at the first / last lattice planes
    const size_t guard = align_up((size_t)(L.sx + L.sy + 1) * esz, 256);
    p += guard;
    o.coef = p; p = align_up(p + 7 * L.NL * esz, 256) + guard;
    p += guard;
    o.vecs = p; p = align_up(p + 15 * L.NL * esz, 256) + guard;
    o.mask = p; p = align_up(p + 3 * L.NL, 256);
    o.valid = p; p = align_up(p + 6 * L.NL, 256);
    o.act = p; p = align_up(p + L.NL + 64, 256);
    size_t np = (size_t)(o.grid_pts > kVecGrid ? o.grid_pts : kVecGrid);
    o.partials = p; p = align_up(p + np * sizeof(double), 256);
    o.st = p; p = align_up(p + sizeof(CgState), 256);
    o.total = p;
    return o;
}

template <typename T> static Visc3Dev<T> dev_view(const fs_visc3d* h) {
    Visc3Dev<T> P;
    P.L = h->L;
    for (int k = 0; k < 7; ++k) P.coef[k] = reinterpret_cast<const T*>(h->coef) + k * h->L.NL;
    for (int k = 0; k < 3; ++k) P.mask[k] = h->mask + k * h->L.NL;
    P.act = h->act;
    return P;
}

template <typename T> static T* vec_ptr(const fs_visc3d* h, int v) { return reinterpret_cast<T*>(h->vecs) + (long long)v * 3 * h->L.NL; }

#define FS_DISPATCH(h, ...)                                      \
    do {                                                         \
        if ((h)->dtype == FS_F32) { using T = float; __VA_ARGS__; } \
        else { using T = double; __VA_ARGS__; }                  \
    } while (0)

// Halo exchange of the three component planes of one lattice array family (solver vector or validity bytes).
// Local planes: 0 = halo from the low neighbour, 1 = first owned; X-3 = last owned, X-2 = halo from the high neighbour
// (plane X-1 only exists because the extended grid is a standalone lattice; nothing owned reads it).
static int visc3d_halo(fs_visc3d* h, char* base /*[3][NL] elements of `esz` bytes*/, size_t esz, int type, cudaStream_t s) {
    if (!h->comm || (!h->has_lo && !h->has_hi)) return FS_OK;
    const Lat3& L = h->L;
    const void* send_lo[3]; void* recv_lo[3]; const void* send_hi[3]; void* recv_hi[3];
    for (int c = 0; c < 3; ++c) {
        char* comp = base + (size_t)c * L.NL * esz;
        recv_lo[c] = comp;
        send_lo[c] = comp + (size_t)1 * L.sx * esz;
        send_hi[c] = comp + (size_t)(L.X - 3) * L.sx * esz;
        recv_hi[c] = comp + (size_t)(L.X - 2) * L.sx * esz;
    }
    return comm_halo_exchange(h->comm, h->has_lo, h->has_hi, 3, send_lo, recv_lo, send_hi, recv_hi, (size_t)L.sx, type, s);
}

static int visc3d_halo_vec(fs_visc3d* h, int vec, cudaStream_t s) {
    return visc3d_halo(h, h->vecs + (size_t)vec * 3 * h->L.NL * h->esz, h->esz, h->dtype == FS_F32 ? COMM_F32 : COMM_F64, s);
}

extern "C" {

int fs_visc3d_set_slab(fs_visc3d* h, fs_comm* comm, int has_lo, int has_hi) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if ((has_lo || has_hi) && !comm) return fail(FS_ERR_ARG, "fs_visc3d_set_slab: neighbours need a communicator");
    if (h->L.nx < 2 + (has_lo ? 1 : 0) + (has_hi ? 1 : 0)) return fail(FS_ERR_ARG, "fs_visc3d_set_slab: slab too thin");
    h->comm = comm;
    h->has_lo = has_lo ? 1 : 0;
    h->has_hi = has_hi ? 1 : 0;
    h->L.u_xhi = h->L.nx - 1 - h->has_hi;
    h->packed = false;
    h->graph.valid = false;
    return FS_OK;
}

int fs_visc3d_set_peers(fs_visc3d* h, void* lo_ws, int lo_nx, void* hi_ws, int hi_nx, void* const* mailboxes) {
    if (!h || !mailboxes) return fail(FS_ERR_ARG, "fs_visc3d_set_peers: null argument");
    if (!h->comm) return fail(FS_ERR_STATE, "fs_visc3d_set_peers: call fs_visc3d_set_slab first");
    if ((h->has_lo && !lo_ws) || (h->has_hi && !hi_ws)) return fail(FS_ERR_ARG, "fs_visc3d_set_peers: missing neighbour workspace mapping");
    if (h->comm->nranks > kMaxRanks) return fail(FS_ERR_ARG, "fs_visc3d_set_peers: more ranks than the mailbox supports");
    PeerInfo pi;
    memset(&pi, 0, sizeof(pi));
    pi.rank = h->comm->rank; pi.nranks = h->comm->nranks;
    pi.has_lo = h->has_lo; pi.has_hi = h->has_hi;
    for (int r = 0; r < pi.nranks; ++r) {
        if (!mailboxes[r]) return fail(FS_ERR_ARG, "fs_visc3d_set_peers: null mailbox");
        pi.mbox[r] = (unsigned long long*)mailboxes[r];
    }
    const size_t esz = h->esz;
    if (h->has_lo) {
        Lat3 Ln = make_lat3(lo_nx, h->L.ny, h->L.nz);
        Visc3Layout ln = visc3_layout(Ln, esz);
        for (int c = 0; c < 3; ++c)
            pi.q_lo[c] = (char*)lo_ws + ln.vecs + (((size_t)FS_VEC_Q * 3 + c) * Ln.NL + (size_t)(Ln.X - 2) * Ln.sx) * esz;
    }
    if (h->has_hi) {
        Lat3 Ln = make_lat3(hi_nx, h->L.ny, h->L.nz);
        Visc3Layout ln = visc3_layout(Ln, esz);
        for (int c = 0; c < 3; ++c)
            pi.q_hi[c] = (char*)hi_ws + ln.vecs + (((size_t)FS_VEC_Q * 3 + c) * Ln.NL) * esz;
    }
    pi.comp_len = h->L.NL;
    pi.halo_lo_end = h->has_lo ? h->L.sx : 0;
    pi.halo_hi_begin = h->has_hi ? (long long)(h->L.X - 2) * h->L.sx : 0;
    pi.halo_hi_end = h->has_hi ? (long long)(h->L.X - 1) * h->L.sx : 0;
    h->hot.has_lo = pi.has_lo; h->hot.has_hi = pi.has_hi;
    for (int c = 0; c < 3; ++c) { h->hot.q_lo[c] = pi.q_lo[c]; h->hot.q_hi[c] = pi.q_hi[c]; }
    h->hot.comp_len = pi.comp_len; h->hot.halo_lo_end = pi.halo_lo_end;
    h->hot.halo_hi_begin = pi.halo_hi_begin; h->hot.halo_hi_end = pi.halo_hi_end;
    for (int c = 0; c < 3; ++c) {
        h->hot.hb[2 * c] = c * pi.comp_len;                        h->hot.he[2 * c] = c * pi.comp_len + pi.halo_lo_end;
        h->hot.hb[2 * c + 1] = c * pi.comp_len + pi.halo_hi_begin; h->hot.he[2 * c + 1] = c * pi.comp_len + pi.halo_hi_end;
    }
    if (!h->peers) FS_CUDA(cudaMalloc((void**)&h->peers, sizeof(PeerInfo)));
    FS_CUDA(cudaMemcpy(h->peers, &pi, sizeof(pi), cudaMemcpyHostToDevice));
    h->graph.valid = false;
    return FS_OK;
}

int fs_visc3d_peer_error(fs_visc3d* h) {
    if (!h || !h->peers) return 0;
    PeerInfo pi;
    if (cudaMemcpy(&pi, h->peers, sizeof(pi), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return pi.error;
}

size_t fs_visc3d_workspace_bytes(int nx, int ny, int nz, int dtype) {
    if (nx < 1 || ny < 1 || nz < 1 || (dtype != FS_F32 && dtype != FS_F64)) return 0;
    Lat3 L = make_lat3(nx, ny, nz);
    return visc3_layout(L, dtype == FS_F32 ? 4 : 8).total;
}

int fs_visc3d_create(fs_visc3d** out, int nx, int ny, int nz, int dtype, void* ws, size_t ws_bytes) {
    if (!out || !ws) return fail(FS_ERR_ARG, "fs_visc3d_create: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_visc3d_create: grid resolution must be >= 1");
    if (dtype != FS_F32 && dtype != FS_F64) return fail(FS_ERR_ARG, "fs_visc3d_create: dtype must be FS_F32 or FS_F64");
    if ((uintptr_t)ws % 256) return fail(FS_ERR_ARG, "fs_visc3d_create: workspace must be 256-byte aligned");
    fs_visc3d* h = new fs_visc3d();
    h->L = make_lat3(nx, ny, nz);
    h->dtype = dtype;
    h->esz = dtype == FS_F32 ? 4 : 8;
    Visc3Layout lay = visc3_layout(h->L, h->esz);
    if (ws_bytes < lay.total) { delete h; return fail(FS_ERR_ARG, "fs_visc3d_create: workspace too small"); }
    h->ws = (char*)ws; h->ws_bytes = ws_bytes;
    h->coef = h->ws + lay.coef; h->vecs = h->ws + lay.vecs;
    h->mask = (uint8_t*)(h->ws + lay.mask); h->valid = (uint8_t*)(h->ws + lay.valid); h->act = (uint8_t*)(h->ws + lay.act);
    h->partials = (double*)(h->ws + lay.partials); h->st = (CgState*)(h->ws + lay.st);
    h->grid_pts = lay.grid_pts;
    h->packed = false;
    h->comm = nullptr; h->has_lo = 0; h->has_hi = 0; h->peers = nullptr;
    memset(&h->hot, 0, sizeof(h->hot));
    int s = h->cg.init();
    if (s < 0) { delete h; return s; }
    h->cg.st_dev = h->st; h->cg.partials_dev = h->partials;
    cudaError_t e = cudaMemset(ws, 0, lay.total);     // cp.zeros semantics for every solver vector
    if (e != cudaSuccess) { h->cg.destroy(); delete h; return fail(FS_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    *out = h;
    return FS_OK;
}

void fs_visc3d_destroy(fs_visc3d* h) {
    if (!h) return;
    if (h->peers) cudaFree(h->peers);
    h->graph.destroy();
    h->cg.destroy();
    delete h;
}

int fs_visc3d_lattice(const fs_visc3d* h, int* X, int* Y, int* Zp, int64_t* NL) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (X) *X = h->L.X;
    if (Y) *Y = h->L.Y;
    if (Zp) *Zp = h->L.Zp;
    if (NL) *NL = h->L.NL;
    return FS_OK;
}

void* fs_visc3d_vector_ptr(const fs_visc3d* h, int vec, int comp) {
    if (!h || vec < 0 || vec >= FS_NUM_VECS || comp < 0 || comp > 2) return nullptr;
    return h->vecs + ((long long)vec * 3 + comp) * h->L.NL * h->esz;
}

int fs_visc3d_pack(fs_visc3d* h, const double* sphi, const double* lvol, double vol_norm, void* stream) {
    if (!h || !sphi || !lvol) return fail(FS_ERR_ARG, "fs_visc3d_pack: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    FS_DISPATCH(h, visc3d_pack_kernel<T><<<h->grid_pts, kThreads, 0, s>>>(h->L, sphi, lvol, vol_norm, reinterpret_cast<T*>(h->coef), h->mask, h->act));
    FS_LAUNCH_CHECK();
    h->packed = true;
    return FS_OK;
}

int fs_visc3d_load(fs_visc3d* h, int vec, const void* vx, const void* vy, const void* vz, int src_dtype, void* stream) {
    if (!h || !vx || !vy || !vz) return fail(FS_ERR_ARG, "fs_visc3d_load: null argument");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc3d_load: bad vector id");
    cudaStream_t s = (cudaStream_t)stream;
    if (src_dtype == FS_F32) {
        FS_DISPATCH(h, visc3d_load_kernel<T, float><<<h->grid_pts, kThreads, 0, s>>>(h->L, (const float*)vx, (const float*)vy, (const float*)vz, vec_ptr<T>(h, vec)));
    } else if (src_dtype == FS_F64) {
        FS_DISPATCH(h, visc3d_load_kernel<T, double><<<h->grid_pts, kThreads, 0, s>>>(h->L, (const double*)vx, (const double*)vy, (const double*)vz, vec_ptr<T>(h, vec)));
    } else return fail(FS_ERR_ARG, "fs_visc3d_load: bad dtype");
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_visc3d_store(fs_visc3d* h, int vec, void* vx, void* vy, void* vz, int dst_dtype, int mode, void* stream) {
    if (!h || !vx || !vy || !vz) return fail(FS_ERR_ARG, "fs_visc3d_store: null argument");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc3d_store: bad vector id");
    if (mode < FS_STORE_ALL || mode > FS_STORE_FLUID) return fail(FS_ERR_ARG, "fs_visc3d_store: bad mode");
    if (mode == FS_STORE_FLUID && !h->packed) return fail(FS_ERR_STATE, "fs_visc3d_store: FS_STORE_FLUID needs fs_visc3d_pack first");
    cudaStream_t s = (cudaStream_t)stream;
    if (dst_dtype == FS_F32) {
        FS_DISPATCH(h, visc3d_store_kernel<T, float><<<h->grid_pts, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->mask, (float*)vx, (float*)vy, (float*)vz, mode));
    } else if (dst_dtype == FS_F64) {
        FS_DISPATCH(h, visc3d_store_kernel<T, double><<<h->grid_pts, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->mask, (double*)vx, (double*)vy, (double*)vz, mode));
    } else return fail(FS_ERR_ARG, "fs_visc3d_store: bad dtype");
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_visc3d_extrapolate(fs_visc3d* h, int vec, int sweeps, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_extrapolate: call fs_visc3d_pack first");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc3d_extrapolate: bad vector id");
    if (sweeps <= 0) return FS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int scratch = (vec == FS_VEC_R) ? FS_VEC_D : FS_VEC_R;   // both are fully rewritten at CG start
    const long long NL3 = 3 * h->L.NL;
    FS_CUDA(cudaMemcpyAsync(h->valid, h->mask, NL3, cudaMemcpyDeviceToDevice, s));   // valid0 = (sphi >= 0)  (:479-481)
    int cur = 0;
    for (int k = 0; k < sweeps; ++k) {
        FS_DISPATCH(h, visc3d_extrapolate_kernel<T><<<h->grid_pts, kThreads, 0, s>>>(
            h->L, vec_ptr<T>(h, cur == 0 ? vec : scratch), h->valid + (size_t)cur * NL3,
            vec_ptr<T>(h, cur == 0 ? scratch : vec), h->valid + (size_t)(cur ^ 1) * NL3));
        FS_LAUNCH_CHECK();
        cur ^= 1;
        if (h->comm) {   // the sweep is Jacobi over the GLOBAL grid: refresh the halo planes of the new values and flags
            FS_TRY(visc3d_halo_vec(h, cur == 0 ? vec : scratch, s));
            FS_TRY(visc3d_halo(h, (char*)(h->valid + (size_t)cur * NL3), 1, COMM_U8, s));
        }
    }
    if (cur == 1) {  // result sits in scratch
        FS_CUDA(cudaMemcpyAsync(h->vecs + (size_t)vec * NL3 * h->esz, h->vecs + (size_t)scratch * NL3 * h->esz, NL3 * h->esz, cudaMemcpyDeviceToDevice, s));
    }
    return FS_OK;
}

static int visc3d_general(fs_visc3d* h, double scale, double mu, int src, int dst, int mode, cudaStream_t s) {
    if (!h->packed) return fail(FS_ERR_STATE, "viscosity operator used before fs_visc3d_pack");
    if (src < 0 || src >= FS_NUM_VECS || dst < 0 || dst >= FS_NUM_VECS || src == dst) return fail(FS_ERR_ARG, "bad src/dst vector ids");
    const double sm = scale * mu;
    if (mode == ROW_APPLY) {
        FS_DISPATCH(h, visc3d_general_kernel<T, ROW_APPLY><<<h->grid_pts, kThreads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, src), vec_ptr<T>(h, dst)));
    } else {
        FS_DISPATCH(h, visc3d_general_kernel<T, ROW_RHS><<<h->grid_pts, kThreads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, src), vec_ptr<T>(h, dst)));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_visc3d_rhs(fs_visc3d* h, double scale, double mu, int src_vec, int dst_vec, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    return visc3d_general(h, scale, mu, src_vec, dst_vec, ROW_RHS, (cudaStream_t)stream);
}

int fs_visc3d_apply(fs_visc3d* h, double scale, double mu, int src_vec, int dst_vec, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    return visc3d_general(h, scale, mu, src_vec, dst_vec, ROW_APPLY, (cudaStream_t)stream);
}

static int visc3d_k1(fs_visc3d* h, double sm, cudaStream_t s) {
    long long want = (h->L.NL + kK1Threads - 1) / kK1Threads;
    const long long cap = (long long)kSMs * (h->dtype == FS_F32 ? K1Occ<float>::value : K1Occ<double>::value);
    const int grid = (int)(want < cap ? want : cap);
    if (h->peers) {
        FS_DISPATCH(h, visc3d_apply_dot_kernel<T, true><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->st, h->partials, h->peers, h->hot));
    } else {
        FS_DISPATCH(h, visc3d_apply_dot_kernel<T, false><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->st, h->partials, nullptr, h->hot));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}
static int visc3d_k2(fs_visc3d* h, cudaStream_t s, int freeze = 0) {
    const long long n = 3 * h->L.NL;
    FS_DISPATCH(h, FS_TRY(cg_launch_update_xr<T>(n, vec_ptr<T>(h, FS_VEC_X), vec_ptr<T>(h, FS_VEC_R), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->st, h->partials, s, freeze, h->peers, h->peers ? &h->hot : nullptr)));
    return FS_OK;
}
static int visc3d_k3(fs_visc3d* h, cudaStream_t s) {
    const long long n = 3 * h->L.NL;
    FS_DISPATCH(h, FS_TRY(cg_launch_update_d<T>(n, vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_R), h->st, s)));
    return FS_OK;
}

static int visc3d_iteration(fs_visc3d* h, double sm, cudaStream_t s) {
    if (h->peers) {
        // fused path: K1 pushes its boundary q planes into the neighbours and all-reduces d.q in its tail; K2 updates the
        // halo rows of r with them and all-reduces r.r in its tail; K3 keeps the halo rows of d current.  No other traffic.
        FS_TRY(visc3d_k1(h, sm, s));
        FS_TRY(visc3d_k2(h, s));
        FS_TRY(visc3d_k3(h, s));
        return FS_OK;
    }
    if (h->comm) FS_TRY(visc3d_halo_vec(h, FS_VEC_D, s));                       // neighbours' d planes for the stencil
    FS_TRY(visc3d_k1(h, sm, s));
    if (h->comm) FS_TRY(comm_allreduce_sum_f64(h->comm, &h->st->dq, 1, s));     // d.q over all slabs
    FS_TRY(visc3d_k2(h, s));
    if (h->comm) {
        FS_TRY(comm_allreduce_sum_f64(h->comm, &h->st->red, 1, s));             // r.r over all slabs
        cg_finish_kernel<<<1, 1, 0, s>>>(h->st, 1);
        FS_LAUNCH_CHECK();
    }
    FS_TRY(visc3d_k3(h, s));
    return FS_OK;
}

// NCCL calls are not captured: graphs are used on a single GPU and with the fused peer-memory transport only
static int visc3d_iterations(fs_visc3d* h, double sm, long long n, cudaStream_t s) {
    const bool graph_ok = !(h->comm && !h->peers);
    return cg_enqueue_iterations(h->graph, graph_ok, sm, n, [&](cudaStream_t ss) { return visc3d_iteration(h, sm, ss); }, s);
}

static int visc3d_cg_begin(fs_visc3d* h, double scale, double mu, double tol, int64_t max_iter, cudaStream_t s) {
    const long long n = 3 * h->L.NL;
    cg_state_init_kernel<<<1, 1, 0, s>>>(h->st, tol * tol, (long long)max_iter, (h->comm && !h->peers) ? 1 : 0);
    FS_LAUNCH_CHECK();
    FS_TRY(visc3d_general(h, scale, mu, FS_VEC_X, FS_VEC_Q, ROW_APPLY, s));   // q = A x   (:575)
    if (h->peers) {
        // fused transport: the halo planes of q still hold what the neighbours stored during the previous solve; they must
        // read as zero (rows not computed here) when r0 = b - q and delta0 are formed
        for (int c = 0; c < 3; ++c) {
            char* comp = h->vecs + ((size_t)FS_VEC_Q * 3 + c) * h->L.NL * h->esz;
            if (h->has_lo) FS_CUDA(cudaMemsetAsync(comp, 0, (size_t)h->L.sx * h->esz, s));
            if (h->has_hi) FS_CUDA(cudaMemsetAsync(comp + (size_t)(h->L.X - 2) * h->L.sx * h->esz, 0, (size_t)h->L.sx * h->esz, s));
        }
    }
    FS_DISPATCH(h, FS_TRY(cg_launch_residual_init<T>(n, vec_ptr<T>(h, FS_VEC_B), vec_ptr<T>(h, FS_VEC_Q), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_R), h->st, h->partials, s, h->peers)));
    if (h->comm && !h->peers) {
        FS_TRY(comm_allreduce_sum_f64(h->comm, &h->st->red, 1, s));
        cg_finish_kernel<<<1, 1, 0, s>>>(h->st, 0);
        FS_LAUNCH_CHECK();
    }
    if (h->peers) {   // one-time: halo rows of r and d (the fused iteration keeps them current from here on)
        FS_TRY(visc3d_halo_vec(h, FS_VEC_R, s));
        FS_TRY(visc3d_halo_vec(h, FS_VEC_D, s));
    }
    return FS_OK;
}

int fs_visc3d_cg(fs_visc3d* h, double scale, double mu, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (max_iter < 0) return fail(FS_ERR_ARG, "fs_visc3d_cg: max_iter < 0");
    cudaStream_t s = (cudaStream_t)stream;
    FS_TRY(visc3d_cg_begin(h, scale, mu, tol, max_iter, s));
    const double sm = scale * mu;
    return cg_drive(h->cg, [&](cudaStream_t ss, long long nb) { return visc3d_iterations(h, sm, nb, ss); }, (long long)max_iter, stats, s);
}

int fs_visc3d_cg_enqueue(fs_visc3d* h, double scale, double mu, int64_t n, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_cg_enqueue before fs_visc3d_pack");
    const double sm = scale * mu;
    cg_state_unlimit_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->st);
    FS_LAUNCH_CHECK();
    return visc3d_iterations(h, sm, n, (cudaStream_t)stream);
}

int fs_visc3d_kernel_enqueue(fs_visc3d* h, int which, double scale, double mu, int64_t n, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_kernel_enqueue before fs_visc3d_pack");
    if (which < 1 || which > 3) return fail(FS_ERR_ARG, "fs_visc3d_kernel_enqueue: which must be 1, 2 or 3");
    cudaStream_t s = (cudaStream_t)stream;
    const double sm = scale * mu;
    cg_state_unlimit_kernel<<<1, 1, 0, s>>>(h->st);
    FS_LAUNCH_CHECK();
    for (int64_t k = 0; k < n; ++k) {
        if (which == 1) FS_TRY(visc3d_k1(h, sm, s));
        else if (which == 2) FS_TRY(visc3d_k2(h, s, 1));
        else FS_TRY(visc3d_k3(h, s));
    }
    return FS_OK;
}

int fs_visc3d_read_stats(fs_visc3d* h, fs_cg_stats* stats, void* stream) {
    if (!h || !stats) return fail(FS_ERR_ARG, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    FS_CUDA(cudaMemcpyAsync(&h->cg.st_pinned[0], h->st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    FS_CUDA(cudaStreamSynchronize(s));
    const CgState& c = h->cg.st_pinned[0];
    stats->iterations = c.iter; stats->delta = c.delta; stats->alpha = c.alpha; stats->beta = c.beta;
    stats->delta0 = c.delta0; stats->converged = (c.done == 1); stats->reserved = 0;
    return FS_OK;
}

int fs_visc3d_solve(fs_visc3d* h, double dt, double mu, double rho, double cell_vol,
                    void* vx, void* vy, void* vz, int vel_dtype, const double* sphi, const double* lvol,
                    double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    const double scale = dt / cell_vol / rho;                                   // :567
    FS_TRY(fs_visc3d_pack(h, sphi, lvol, cell_vol * 0.125, stream));            // :568
    FS_TRY(fs_visc3d_load(h, FS_VEC_X, vx, vy, vz, vel_dtype, stream));         // :569-571
    FS_TRY(fs_visc3d_extrapolate(h, FS_VEC_X, 3, stream));                      // :573
    FS_TRY(fs_visc3d_rhs(h, scale, mu, FS_VEC_X, FS_VEC_B, stream));            // :574
    int status = fs_visc3d_cg(h, scale, mu, tol, max_iter, stats, stream);      // :575-612
    if (status != FS_OK) return status;                                         // the reference raises before write-back
    FS_TRY(fs_visc3d_store(h, FS_VEC_X, vx, vy, vz, vel_dtype, FS_STORE_FLUID, stream));   // :613
    FS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FS_OK;
}

}  // extern "C"
