// 3-D variational viscosity solver on a packed MAC lattice (sm_100a).
//
// Replaces ViscosityCGSolver3D.py (kernels :8-470, launchers :472-530, solve :566-613).
//
// HBM layout (all SoA, z contiguous).  One padded lattice  X=nx+1, Y=ny+1, Zp=roundup(nz+1,4)  is
// shared by every array so a single flat index i=(x*Y+y)*Zp+z addresses the u,v,w faces, the cell
// centre and the three low edges that belong to lattice point (x,y,z); neighbours are i+-1, i+-Zp,
// i+-Y*Zp for every array alike.
//   coef[0..2]  Vu,Vv,Vw   face liquid-volume fraction (NaN where the row is never computed)
//   coef[3]     Vc         cell-centre volume
//   coef[4..6]  Exy,Exz,Eyz edge-centred volumes
//   mask[0..2]  fluid flag of each face (sphi >= 0), bytes; used by the masked apply / RHS /
//                          extrapolation / write-back only — never inside the CG loop
//   act         1 byte per lattice point: bit c = row of component c is computed (bits 4..6: by a neighbour slab)
//   seg list    sorted ids of the active 32-point segments (any act bit set): what K1/K2/K3 walk
//   vec[v][c]   X,R,D,Q,B  solver vectors, 3 components each, NL elements per component;
//                          R,D,Q,B are exactly zero outside the active segments (invariant kept by pack)
//
// One solve:  pack (+ activity map, segment list) -> load -> 3 in-place extrapolation sweeps -> begin (b, q=Ax, r=d=b-q,
// delta0 on the active set) -> CG iterations (one persistent cooperative kernel, or K1/K2/K3 from a CUDA graph for
// HBM-sized active sets) -> write-back of the active rows.
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
    int has_lo, has_hi;   // multi-GPU: plane 0 / plane nx-1 mirrors rows owned by the lower / higher slab
    // x-window of the once-per-solve passes (pack, load, extrapolation): the caller's arrays hold the cells [wlo, wcells) of
    // the grid only — lattice planes wlo..whi are (re)built from them.  Whole grid: wlo = 0, whi = nx, wcells = nx.
    // (gathered multi-GPU solve: every rank packs its own window of the GLOBAL lattice; chunked host uploads)
    int wlo, whi, wcells;
};

template <typename T> struct Visc3Dev {
    Lat3 L;
    const T* coef[7];
    const T* cs[7];          // pre-scaled coefficients of the CG-loop apply (visc3d_scale_kernel): diag_u, diag_v, diag_w, 2s*Vc, s*Exy, s*Exz, s*Eyz
    const uint8_t* mask[3];
    const uint8_t* act;      // per lattice point: bit c set <=> row c is computed (same information as the NaN tags, 1 byte)
};

__device__ __forceinline__ void lat_decode(const Lat3& L, long long i, int& x, int& y, int& z) {
    if (L.NL < 0x7fffffffLL) {                       // 32-bit divisions (the common case) are several times cheaper
        const unsigned int u = (unsigned int)i, zp = (unsigned int)L.Zp, yy = (unsigned int)L.Y;
        const unsigned int t = u / zp;
        z = (int)(u - t * zp);
        const unsigned int xx = t / yy;
        y = (int)(t - xx * yy);
        x = (int)xx;
        return;
    }
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
// pack: fine-grid (2n+1)^3 fp64 sphi / lvol  ->  lattice coefficient planes + face masks + activity map.
// One block per lattice row (x,y); threads run along z (no per-thread index division).
//
// Activity map (1 byte per lattice point): bit c = row of component c is COMPUTED by this rank — fluid face
// (sphi >= 0), interior, owned — and, with FS_ACTIVE_NONZERO, at least one of its seven coefficients (face volume, two
// cell-centre volumes, four edge volumes = the fine node of the face and its six fine-grid neighbours) is non-zero.
// A row whose coefficients are all zero is an all-zero row AND column of the operator: b, q, r, d are exactly 0 there
// for the whole solve and x never changes (faces far from any liquid), so dropping it changes no result bit.
// Bits 4..6: the row is computed by a neighbour slab and mirrored here (multi-GPU halo planes).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512, 3) visc3d_pack_kernel(Lat3 L, const double* __restrict__ sphi, const double* __restrict__ lvol,
                                                          double vol_norm, T* __restrict__ coef /*[7][NL]*/, uint8_t* __restrict__ mask /*[3][NL]*/,
                                                          uint8_t* __restrict__ act /*[NL + 64]*/, uint8_t* __restrict__ rowflag /*[X*Y]*/,
                                                          uint8_t* __restrict__ rownz /*[X*Y] in/out: the row's coefficient planes hold a non-zero*/,
                                                          int nonzero_only) {
    const int row = blockIdx.x + L.wlo * L.Y;        // x*Y + y
    int any_valid = 0;                               // does this lattice row hold any fluid face? (extrapolation sweep 1 skips far rows)
    const int x = row / L.Y, y = row - x * L.Y;
    const long long fz = 1, fy = 2LL * L.nz + 1, fx = fy * (2LL * L.ny + 1);
    const long long frow = 2LL * (x - L.wlo) * fx + 2LL * y * fy;     // the fine grids start at fine plane 2*wlo
    const bool ix = x < L.wcells, iy = y < L.ny;
    // windowed pack: the activity test of a row reads fine planes 2x-1 .. 2x+2; on the first plane of a window that does not
    // start at the grid boundary, and on the closing plane of one that does not end there, they are not all present — those
    // planes get no activity bits from this rank (their owner supplies them)
    const bool act_ok = !((L.wlo > 0 && x == L.wlo) || (L.wcells < L.nx && x >= L.wcells));
    // rows of these planes are computed by a neighbour slab and mirrored here (K2/K3 keep r and d current on them)
    const bool halo_x = (L.has_lo && x == 0) || (L.has_hi && x == L.nx - 1);
    // Liquid-free lattice rows (all ten fine-grid volumes of every point zero: most of a typical scene) whose seven
    // coefficient planes ALREADY hold zeros from the previous pack are not rewritten: a third of this kernel's traffic.
    // Needs the row-wide verdict before the stores, i.e. one z per thread (rows longer than the block are always written).
    const bool can_skip = L.Zp <= (int)blockDim.x;
    const bool was_nz = rownz[row] != 0;
    for (int z0 = 0; z0 < L.Zp; z0 += blockDim.x) {
        const int z = z0 + threadIdx.x;
        const bool live = z < L.Zp;                  // (all threads stay in the loop: block-wide vote below)
        const long long i = (long long)row * L.Zp + (live ? z : 0);
        const long long f0 = frow + 2LL * (live ? z : 0);         // fine node (2x,2y,2z)
        const bool iz = live && z < L.nz, inz = live && z <= L.nz;
        // All ten fine-grid values of this lattice point are requested first (one memory latency), then normalised.
        // vol = lvol / vol_norm (:568): the fp64 division (~25 instructions) is skipped for the very common exact zeros, whose
        // quotient is the same zero; the empty volatile asm keeps ptxas from if-converting the branch (it otherwise evaluates
        // all 16 divisions of a lattice point unconditionally and selects: 45 % of the kernel's instructions).
        auto norm = [&](double l) -> T {
            if (l != 0.0) {
                asm volatile("" : "+d"(l));
                l = l / vol_norm;
            }
            return (T)l;
        };
        auto vol = [&](long long off) -> T { return norm(lvol[f0 + off]); };
        auto nz = [](T v) { return v != T(0); };     // NaN counts as non-zero: it must propagate like in the reference
        const bool in_c = ix && iy && iz, in_u = iy && iz, in_v = ix && iz, in_w = ix && iy && inz;
        const double l_vc = in_c ? lvol[f0 + fx + fy + fz] : 0.0;    // cell centre (1,1,1)
        const double l_exy = iz ? lvol[f0 + fz] : 0.0;               // Exy: (0,0,1)
        const double l_exz = (iy && inz) ? lvol[f0 + fy] : 0.0;      // Exz: (0,1,0)
        const double l_eyz = (ix && inz) ? lvol[f0 + fx] : 0.0;      // Eyz: (1,0,0)
        const double l_u = in_u ? lvol[f0 + fy + fz] : 0.0;          // faces: fine parities (0,1,1) (1,0,1) (1,1,0); they share
        const double l_v = in_v ? lvol[f0 + fx + fz] : 0.0;          // their fine rows (and DRAM sectors) with Exz, Eyz, Vc
        const double l_w = in_w ? lvol[f0 + fx + fy] : 0.0;
        const double s_u = in_u ? sphi[f0 + fy + fz] : -1.0;
        const double s_v = in_v ? sphi[f0 + fx + fz] : -1.0;
        const double s_w = in_w ? sphi[f0 + fx + fy] : -1.0;
        const T vc = norm(l_vc), exy = norm(l_exy), exz = norm(l_exz), eyz = norm(l_eyz);
        T vface[3];
        unsigned int abits = 0;
        {
            const bool fluid = in_u && s_u >= 0.0;
            T v = T(0);
            const bool yz = y >= 1 && y <= L.ny - 2 && z >= 1 && z <= L.nz - 2;
            if (fluid && yz && x >= 1 && x <= L.u_xhi && act_ok) {    // (act_ok: fine plane 2x-1 is present)
                v = norm(l_u);
                bool on = true;
                if (nonzero_only) on = nz(v) || nz(vc) || nz(exy) || nz(exz) || nz(vol(fy + fz - fx)) || nz(vol(fy + fz + fy)) || nz(vol(fy + fz + fz));
                if (on) abits |= 1u;
            }
            if (fluid && halo_x && yz) abits |= 0x10u;
            vface[0] = v;
            if (live) mask[0 * L.NL + i] = fluid;
        }
        {
            const bool fluid = in_v && s_v >= 0.0;
            T v = T(0);
            const bool yz = y >= 1 && y <= L.ny - 1 && z >= 1 && z <= L.nz - 2;
            if (fluid && yz && x >= 1 && x <= L.nx - 2) {
                v = norm(l_v);
                bool on = true;
                if (nonzero_only) on = nz(v) || nz(vc) || nz(exy) || nz(eyz) || nz(vol(fx + fz + fx)) || nz(vol(fx + fz - fy)) || nz(vol(fx + fz + fz));
                if (on) abits |= 2u;
            }
            if (fluid && halo_x && yz) abits |= 0x20u;
            vface[1] = v;
            if (live) mask[1 * L.NL + i] = fluid;
        }
        {
            const bool fluid = in_w && s_w >= 0.0;
            T v = T(0);
            const bool yz = y >= 1 && y <= L.ny - 2 && z >= 1 && z <= L.nz - 1;
            if (fluid && yz && x >= 1 && x <= L.nx - 2) {
                v = norm(l_w);
                bool on = true;
                if (nonzero_only) on = nz(v) || nz(vc) || nz(exz) || nz(eyz) || nz(vol(fx + fy + fx)) || nz(vol(fx + fy + fy)) || nz(vol(fx + fy - fz));
                if (on) abits |= 4u;
            }
            if (fluid && halo_x && yz) abits |= 0x40u;
            vface[2] = v;
            if (live) mask[2 * L.NL + i] = fluid;
        }
        any_valid |= (in_u && s_u >= 0.0) || (in_v && s_v >= 0.0) || (in_w && s_w >= 0.0);
        const int nz_pt = nz(vface[0]) || nz(vface[1]) || nz(vface[2]) || nz(vc) || nz(exy) || nz(exz) || nz(eyz);
        bool write = true;
        if (can_skip) {
            const int nz_row = __syncthreads_or(nz_pt);
            write = nz_row || was_nz;                // zero now and zero before: the planes already hold this row's zeros
            if (threadIdx.x == 0) rownz[row] = (uint8_t)(nz_row != 0);
        }
        if (live && write) {
            coef[0 * L.NL + i] = vface[0];
            coef[1 * L.NL + i] = vface[1];
            coef[2 * L.NL + i] = vface[2];
            coef[3 * L.NL + i] = vc;
            coef[4 * L.NL + i] = exy;
            coef[5 * L.NL + i] = exz;
            coef[6 * L.NL + i] = eyz;
        }
        if (live) act[i] = (uint8_t)(act_ok ? abits : 0u);
    }
    if (!can_skip && threadIdx.x == 0) rownz[row] = 1;
    any_valid = __syncthreads_or(any_valid);
    if (threadIdx.x == 0) rowflag[row] = (uint8_t)(any_valid != 0);
}

// ---------------------------------------------------------------------------------------------
// load / store between the caller's dense MAC arrays and lattice vectors
// ---------------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(512) visc3d_load_kernel(Lat3 L, const S* __restrict__ a0, const S* __restrict__ a1, const S* __restrict__ a2,
                                                          T* __restrict__ vec /*[3][NL]*/) {
    const int row = blockIdx.x + L.wlo * L.Y;        // x*Y + y, one block per lattice row of the window
    const int x = row / L.Y, y = row - x * L.Y;
    const S* src[3] = {a0, a1, a2};
    for (int z = threadIdx.x; z < L.Zp; z += blockDim.x) {
        const long long i = (long long)row * L.Zp + z;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int s0, s1, s2;
            comp_shape(L, c, s0, s1, s2);
            const int xs = c == 0 ? L.whi + 1 : L.wcells;          // planes present in the caller's (windowed) array: [wlo, xs)
            T v = T(0);
            if (x < xs && y < s1 && z < s2) v = (T)src[c][((long long)(x - L.wlo) * s1 + y) * s2 + z];
            vec[c * L.NL + i] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sparse set-up (fs_visc3d_solve on a single GPU).  The CG reads the start vector only on the active rows and their
// stencil neighbours; after three extrapolation sweeps those values depend on nothing farther than 4 lattice points
// (Chebyshev) from an active row.  So instead of loading and sweeping the whole lattice, the solve marks the segments
// within 5 rows / planes and 8 points of an active segment (a superset of that neighbourhood, one spare layer),
// loads the caller's velocities there only, and runs sweep 1 over those segments only (sweeps 2 and 3 follow the work
// lists as before).  The validity bytes are still initialised everywhere, so every sweep takes exactly the decisions of
// the dense pass wherever its inputs are complete: by induction the state after sweep k equals the dense one within
// 5 - k layers of the active rows — in particular on everything the solve reads.  Outside the marked region the lattice
// vector x keeps whatever an earlier solve left there (nothing reads it; apply_viscosity only writes active rows back).
// ---------------------------------------------------------------------------------------------
constexpr int kSparseSetupDefault = 1;    // fs_visc3d_solve loads / extrapolates around the active set only unless "sparse_setup" / FLUIDSOLVER_B200_SPARSE_SETUP = 0
constexpr int kRegionLayers = 5;          // rows / planes around an active segment (4 are needed)
constexpr int kRegionPoints = 8;          // points along z (4 are needed)

__global__ void __launch_bounds__(256) visc3d_mark_region_kernel(Lat3 L, const int* __restrict__ seg, const int* __restrict__ nseg_p, long long nseg_total,
                                                                 uint8_t* __restrict__ flags) {
    constexpr int side = 2 * kRegionLayers + 1;
    const long long n = (long long)*nseg_p * side * side;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / (side * side);
        const int o = (int)(e - k * (side * side));
        const int dx = o / side - kRegionLayers, dy = o % side - kRegionLayers;
        const long long base = (long long)__ldg(seg + k) * kSegPts + dx * L.sx + dy * L.sy;
        long long lo = base - kRegionPoints, hi = base + kSegPts - 1 + kRegionPoints;
        const long long w_lo = (long long)L.wlo * L.sx, w_hi = (long long)(L.whi + 1) * L.sx - 1;     // the planes this handle loads (whole lattice: 0 .. NL-1)
        if (hi < w_lo || lo > w_hi) continue;       // (rows that fall off a plane wrap into its neighbour: a harmless superset)
        if (lo < w_lo) lo = w_lo;
        if (hi > w_hi) hi = w_hi;
        for (long long t = lo / kSegPts; t <= hi / kSegPts && t < nseg_total; ++t) flags[t] = 1;
    }
}

// load (visc3d_load_kernel) on a list of segments: a warp per segment
template <typename T, typename S>
__global__ void __launch_bounds__(kThreads) visc3d_load_region_kernel(Lat3 L, const S* __restrict__ a0, const S* __restrict__ a1, const S* __restrict__ a2,
                                                                      T* __restrict__ vec /*[3][NL]*/, const int* __restrict__ seg, const int* __restrict__ nseg_p) {
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const S* src[3] = {a0, a1, a2};
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)__ldg(seg + k) * kSegPts + lane;
        if (i >= L.NL) continue;
        int x, y, z;
        lat_decode(L, i, x, y, z);
        if (x < L.wlo || x > L.whi) continue;       // (a segment that straddles the first / last plane of an x-window)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int s0, s1, s2;
            comp_shape(L, c, s0, s1, s2);
            const int xs = c == 0 ? L.whi + 1 : L.wcells;          // planes present in the caller's (windowed) array: [wlo, xs)
            T v = T(0);
            if (x < xs && y < s1 && z < s2) v = (T)src[c][((long long)(x - L.wlo) * s1 + y) * s2 + z];
            vec[c * L.NL + i] = v;
        }
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
// extrapolation sweep (extrapolate_kernel :8-39), all three components per launch, IN PLACE.
// The reference ping-pongs full copies of the three velocity arrays and validity masks per sweep
// (Jacobi).  Here the validity byte carries a generation: 0 = invalid, 1 = valid from the start
// (sphi >= 0), k+1 = filled by sweep k.  Sweep k treats a neighbour as valid iff 1 <= byte <= k, so
// a face that is being filled concurrently (byte 0 or k+1) is never read in the same sweep, and a
// value is only ever written to a face nobody reads during that sweep: identical to Jacobi, without
// copies — a sweep reads the validity bytes and touches values only next to the fluid/solid interface.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bytes_all_nonzero(uint32_t w) {
    return ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u) == 0x80808080u;
}

// Work lists of the extrapolation: the faces filled by sweep k are the only places next to which sweep k+1 can fill
// anything, so sweep 1 (a full pass) records them and the later sweeps only look around them.
struct ExtrapWork {
    unsigned int* list[2];      // flat face ids c*NL + i, ping-pong between sweeps
    unsigned int* count;        // [0..1] entries of list[k], [2] overflow flag
    unsigned int cap;           // 0 = lists disabled (full passes only)
};

__device__ __forceinline__ void extrap_push(const ExtrapWork& W, int which, unsigned int face) {
    if (W.cap == 0u) return;
    // warp-aggregated: the lanes that reach this point together take ONE ticket range (hundreds of thousands of faces are
    // recorded per sweep; one atomic per lane on a single counter serialised the list sweeps at ~50 us each)
    const unsigned int m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(W.count + which, (unsigned int)__popc(m));
    base = __shfl_sync(m, base, leader);
    const unsigned int k = base + __popc(m & ((1u << lane) - 1u));
    if (k < W.cap) (which ? W.list[1] : W.list[0])[k] = face;      // (no dynamic index into the by-value struct: that is a stack copy)
    else W.count[2] = 1u;                          // overflow: the next sweep falls back to a full pass
}

// fill face (c,i) if it is invalid, interior and has a neighbour that was valid before this sweep (:12-38)
template <typename T>
__device__ __forceinline__ bool extrap_try_fill(const Lat3& L, T* v_all, uint8_t* valid_all, int c, long long i, int x, int y, int z, unsigned int sw) {
    int s0, s1, s2;
    comp_shape(L, c, s0, s1, s2);
    if (!(x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2 && z >= 1 && z <= s2 - 2)) return false;
    T* v = v_all + c * L.NL;
    uint8_t* va = valid_all + c * L.NL;
    if (va[i] != 0) return false;
    auto ok = [&](unsigned int g) { return g >= 1u && g <= sw; };
    T val = T(0);
    int count = 0;                                 // +x,-x,+y,-y,+z,-z
    if (ok(va[i + L.sx])) { val += v[i + L.sx]; ++count; }
    if (ok(va[i - L.sx])) { val += v[i - L.sx]; ++count; }
    if (ok(va[i + L.sy])) { val += v[i + L.sy]; ++count; }
    if (ok(va[i - L.sy])) { val += v[i - L.sy]; ++count; }
    if (ok(va[i + 1])) { val += v[i + 1]; ++count; }
    if (ok(va[i - 1])) { val += v[i - 1]; ++count; }
    if (count == 0) return false;
    v[i] = val / (T)count;
    va[i] = (uint8_t)(sw + 1u);
    return true;
}

// Full pass.  One thread per 4 z-consecutive lattice points (Zp % 4 == 0): the validity bytes are read as words, a group
// whose twelve bytes are all valid (fluid regions) exits after three loads, and the index decode and the neighbour
// words are only touched where something may have to be filled.  Grid-stride, so the same kernel serves as the
// fall-back of the list-driven sweeps (`only_if_overflow`).
// one 4-point group of a sweep (see the kernels below)
template <typename T>
__device__ __forceinline__ void extrap_group(const Lat3& L, T* v_all, uint8_t* valid_all, int sweep, const ExtrapWork& W, int push_to,
                                             const uint8_t* __restrict__ rowflag, long long g) {
    const int nrows = L.X * L.Y;
    const unsigned int sw = (unsigned int)sweep;
    const uint32_t swv = sw * 0x01010101u;           // (sweep <= 250: fits a byte)
    const long long i4 = g * 4;
    if (rowflag) {
        // sweep 1 can only fill next to an originally valid face: a group whose lattice row and the four rows around it
        // (y+-1, x+-1) hold no fluid face at all (deep inside the solid, outside the container) is skipped on five bytes
        const int row = (int)(L.NL < 0x7fffffffLL ? (unsigned int)i4 / (unsigned int)L.Zp : i4 / L.Zp);
        const int ra = row > 0 ? row - 1 : row, rb = row + 1 < nrows ? row + 1 : row;
        const int rc = row >= L.Y ? row - L.Y : row, rd = row + L.Y < nrows ? row + L.Y : row;
        if (!(__ldg(rowflag + row) | __ldg(rowflag + ra) | __ldg(rowflag + rb) | __ldg(rowflag + rc) | __ldg(rowflag + rd))) return;
    }
    uint32_t own[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) own[c] = *reinterpret_cast<const uint32_t*>(valid_all + c * L.NL + i4);
    if (bytes_all_nonzero(own[0]) && bytes_all_nonzero(own[1]) && bytes_all_nonzero(own[2])) return;
    int x, y, z0;
    lat_decode(L, i4, x, y, z0);
    // the generation words of the four x/y neighbour groups and the two z-adjacent bytes, requested for all three
    // components before any of them is looked at (one memory latency instead of three)
    uint32_t nb[3][4];
    unsigned int zl[3], zr[3];
    bool need[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int s0, s1, s2;
        comp_shape(L, c, s0, s1, s2);
        need[c] = !bytes_all_nonzero(own[c]) && x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2;
        const uint8_t* va = valid_all + c * L.NL + i4;
        nb[c][0] = need[c] ? *reinterpret_cast<const uint32_t*>(va + L.sx) : 0u;
        nb[c][1] = need[c] ? *reinterpret_cast<const uint32_t*>(va - L.sx) : 0u;
        nb[c][2] = need[c] ? *reinterpret_cast<const uint32_t*>(va + L.sy) : 0u;
        nb[c][3] = need[c] ? *reinterpret_cast<const uint32_t*>(va - L.sy) : 0u;
        zl[c] = (need[c] && z0 > 0) ? va[-1] : 0u;
        zr[c] = (need[c] && z0 + 4 < L.Zp) ? va[4] : 0u;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (!need[c]) continue;
        // deep inside the solid every neighbour generation is 0: nothing can be filled (most groups that get here)
        if ((nb[c][0] | nb[c][1] | nb[c][2] | nb[c][3] | own[c] | zl[c] | zr[c]) == 0u) continue;
        int s0, s1, s2;
        comp_shape(L, c, s0, s1, s2);
        T* v = v_all + c * L.NL;
        uint8_t* va = valid_all + c * L.NL;
        // Which of the four faces can be filled is decided on whole words (byte-wise SIMD compares: 0xff where the
        // generation g satisfies 1 <= g <= sweep), so a warp that only grazes the fluid/solid interface — nearly every warp
        // that gets here — pays one short loop trip per candidate instead of the fully unrolled 4-face body.
        auto okw = [&](uint32_t w) { return __vcmpgeu4(w, 0x01010101u) & __vcmpleu4(w, swv); };
        auto okb = [&](unsigned int gg) { return gg >= 1u && gg <= sw; };
        const uint32_t oxp = okw(nb[c][0]), oxm = okw(nb[c][1]), oyp = okw(nb[c][2]), oym = okw(nb[c][3]);
        const uint32_t oown = okw(own[c]);
        const uint32_t ozp = (oown >> 8) | (okb(zr[c]) ? 0xff000000u : 0u);       // byte k: face k+1 of the group, or the next group's first
        const uint32_t ozm = (oown << 8) | (okb(zl[c]) ? 0x000000ffu : 0u);       // byte k: face k-1, or the previous group's last
        uint32_t cand = __vcmpeq4(own[c], 0u) & (oxp | oxm | oyp | oym | ozp | ozm);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!(z0 + k >= 1 && z0 + k <= s2 - 2)) cand &= ~(0xffu << (8 * k));
        while (cand) {
            const int sh = (__ffs((int)cand) - 1) & ~7;
            cand &= ~(0xffu << sh);
            const long long i = i4 + (sh >> 3);
            T val = T(0);
            int count = 0;                                // +x,-x,+y,-y,+z,-z  (:19-36)
            if ((oxp >> sh) & 1u) { val += v[i + L.sx]; ++count; }
            if ((oxm >> sh) & 1u) { val += v[i - L.sx]; ++count; }
            if ((oyp >> sh) & 1u) { val += v[i + L.sy]; ++count; }
            if ((oym >> sh) & 1u) { val += v[i - L.sy]; ++count; }
            if ((ozp >> sh) & 1u) { val += v[i + 1]; ++count; }
            if ((ozm >> sh) & 1u) { val += v[i - 1]; ++count; }
            v[i] = val / (T)count;
            va[i] = (uint8_t)(sweep + 1);
            if (push_to >= 0) extrap_push(W, push_to, (unsigned int)(c * L.NL + i));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_extrapolate_kernel(Lat3 L, T* v_all, uint8_t* valid_all, int sweep /*1-based*/, ExtrapWork W,
                                                                      int push_to /*list to record fills in, -1 = none*/, int only_if_overflow,
                                                                      const uint8_t* __restrict__ rowflag /*sweep 1 only, or null*/,
                                                                      long long g_begin, long long g_end /*4-point groups of the swept x-planes*/) {
    if (only_if_overflow && (W.cap == 0u || W.count[2] == 0u)) return;
    for (long long g = g_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; g < g_end; g += (long long)gridDim.x * blockDim.x)
        extrap_group<T>(L, v_all, valid_all, sweep, W, push_to, rowflag, g);
}

// The same pass restricted to the 4-point groups of a list of 32-point segments (sparse set-up: the neighbourhood of the
// active set, see visc3d_mark_region_kernel).
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_extrapolate_region_kernel(Lat3 L, T* v_all, uint8_t* valid_all, int sweep, ExtrapWork W, int push_to,
                                                                             const uint8_t* __restrict__ rowflag, const int* __restrict__ seg,
                                                                             const int* __restrict__ nseg_p, long long g_begin, long long g_end) {
    const long long n8 = (long long)*nseg_p * (kSegPts / 4);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += (long long)gridDim.x * blockDim.x) {
        const long long g = (long long)__ldg(seg + (e >> 3)) * (kSegPts / 4) + (e & 7);
        if (g < g_begin || g >= g_end) continue;    // (the swept x-planes: the whole lattice, or an x-window minus its edge planes)
        extrap_group<T>(L, v_all, valid_all, sweep, W, push_to, rowflag, g);
    }
}

// List-driven sweep k >= 2: only the six same-component neighbours of the faces filled by sweep k-1 can be filled now.
// Two entries may target the same face: both compute the same value from the same (final) inputs, so the race is benign.
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_extrapolate_list_kernel(Lat3 L, T* v_all, uint8_t* valid_all, int sweep, ExtrapWork W,
                                                                           int from, int push_to) {
    if (W.count[2] != 0u) return;                  // overflowed: the full-pass fall-back does this sweep
    const unsigned int n = W.count[from];
    const unsigned int sw = (unsigned int)sweep;
    // one thread per (recorded face, direction): the six probes of a face are independent memory round trips
    const unsigned long long n6 = 6ull * n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e6 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e6 < n6; e6 += stride) {
        const unsigned int e = (unsigned int)(e6 / 6ull), k = (unsigned int)(e6 - 6ull * e);
        const unsigned int face = (from ? W.list[1] : W.list[0])[e];
        const int c = (int)(face / (unsigned int)L.NL);
        const long long i = (long long)(face - (unsigned int)c * (unsigned int)L.NL);
        const long long step = k < 2u ? L.sx : k < 4u ? L.sy : 1;
        const long long j = (k & 1u) ? i - step : i + step;
        if (j < 0 || j >= L.NL) continue;
        if (valid_all[c * L.NL + j] != 0) continue;
        int x, y, z;
        lat_decode(L, j, x, y, z);
        if (extrap_try_fill<T>(L, v_all, valid_all, c, j, x, y, z, sw) && push_to >= 0)
            extrap_push(W, push_to, (unsigned int)(c * L.NL + j));
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
// Start of a solve on the active set: b = RHS(x), q = A x (both with the reference's neighbour masks and
// association, ViscosityCGSolver3D.py:574-575), d = r = b - q, delta0 = r.r (:577-587) — one pass over the active
// segments.  Rows outside the active set hold b = q = r = d = 0 (invariant kept by visc3d_clear_kernel).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_begin_kernel(Visc3Dev<T> P, T s, T s2, const T* __restrict__ xv, T* __restrict__ b, T* __restrict__ q,
                                                                T* __restrict__ r, T* __restrict__ d, const int* __restrict__ seg,
                                                                const int* __restrict__ nseg_p, CgState* st_, double* partials, PeerInfo* peers) {
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const long long st[3] = {L.sx, L.sy, 1};
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const uint8_t* const* mask = P.mask;
    auto nb_apply = [&](int comp, long long j) -> T { return mask[comp][j] ? xv[comp * NL + j] : T(0); };
    auto nb_rhs = [&](int comp, long long j) -> T { return mask[comp][j] ? T(0) : xv[comp * NL + j]; };
    double acc = 0.0;
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)__ldg(seg + k) * kSegPts + lane;
        if (i >= NL) continue;
        const unsigned int a = (unsigned int)P.act[i] & kActCompute;
        if (a == 0u) continue;
        auto row = [&](auto Atag) {
            constexpr int A = decltype(Atag)::value;
            if (!(a & (1u << A))) return;
            const T center = P.coef[A][i];
            const T own = xv[A * NL + i];
            const T bb = visc_row<T, 3, A, true, ROW_RHS>(P.coef, i, st, center, own, s, s2, nb_rhs);
            const T qq = visc_row<T, 3, A, true, ROW_APPLY>(P.coef, i, st, center, own, s, s2, nb_apply);
            const T rr = bb - qq;
            b[A * NL + i] = bb; q[A * NL + i] = qq; r[A * NL + i] = rr; d[A * NL + i] = rr;
            acc += (double)rr * (double)rr;
        };
        row(std::integral_constant<int, 0>{});
        row(std::integral_constant<int, 1>{});
        row(std::integral_constant<int, 2>{});
    }
    grid_sum_finish(acc, partials, &st_->counter[2], [=](double sum) {
        if (st_->dist) { st_->red = sum; return; }
        st_->delta = sum;
        st_->delta0 = sum;
        st_->delta_old = sum;
        if (sum < st_->tol2) st_->done = 1;           // `if not self.delta < tol ** 2:` skips the loop
    }, peers, 1);
}

// zero r, d, q, b on the segments of the (previous) active list
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_clear_kernel(long long NL, T* __restrict__ vecs /*[5][3][NL]*/, T* __restrict__ d2 /*[2][3][NL]*/,
                                                                const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                uint8_t* __restrict__ act_clear /*gathered mode: also forget the activity bits, else null*/) {
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)__ldg(seg + k) * kSegPts + lane;
        if (i >= NL) continue;
#pragma unroll
        for (int v = FS_VEC_R; v <= FS_VEC_B; ++v)
#pragma unroll
            for (int c = 0; c < 3; ++c) vecs[((long long)v * 3 + c) * NL + i] = T(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) d2[(long long)c * NL + i] = T(0);        // both parities of w
        if (act_clear) act_clear[i] = 0;
    }
}

// apply_viscosity (:458-470) restricted to the rows the solve can have changed (the computed rows of the active set).
// Every other fluid face still holds the value that was loaded from the very same caller arrays.
template <typename T, typename S>
__global__ void __launch_bounds__(kThreads) visc3d_store_active_kernel(Lat3 L, const T* __restrict__ vec, const uint8_t* __restrict__ act,
                                                                       const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                       S* __restrict__ a0, S* __restrict__ a1, S* __restrict__ a2,
                                                                       int own_lo, int own_hi /*x-planes [own_lo, own_hi) are written; the arrays start at plane L.wlo*/) {
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    S* dst[3] = {a0, a1, a2};
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)__ldg(seg + k) * kSegPts + lane;
        if (i >= L.NL) continue;
        const unsigned int a = (unsigned int)act[i] & kActCompute;
        if (a == 0u) continue;
        int x, y, z;
        lat_decode(L, i, x, y, z);
        if (x < own_lo || x >= own_hi) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!(a & (1u << c))) continue;
            int s0, s1, s2;
            comp_shape(L, c, s0, s1, s2);
            dst[c][((long long)(x - L.wlo) * s1 + y) * s2 + z] = (S)vec[c * L.NL + i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pre-scaled coefficients for the CG-loop apply, once per solve, on the active segments: the row diagonals (reference
// association: diag = vol + scale*mu*(2*hi_x + 2*lo_x + hi_y + ...), :268) and the products 2*scale*mu*Vc, scale*mu*E the
// reference forms first in every off-diagonal term.  An active row also reads volumes that sit at neighbour points
// (i-e_A for Vc, i+e_ax for the edges), possibly in segments that are not active themselves, so each lane writes the scaled
// values of its own point AND of the nine neighbour slots its rows read; duplicates store identical values.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_scale_kernel(Visc3Dev<T> P, T s, T s2, T* __restrict__ cs /*[7][NL]*/,
                                                                const int* __restrict__ seg, const int* __restrict__ nseg_p) {
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const long long st[3] = {L.sx, L.sy, 1};
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)__ldg(seg + k) * kSegPts + lane;
        if (i >= NL) continue;
        auto put = [&](int plane, long long j, T c) {
            if (j >= 0 && j < NL) cs[plane * NL + j] = c * __ldg(P.coef[plane] + j);
        };
        put(3, i, s2); put(3, i - st[0], s2); put(3, i - st[1], s2); put(3, i - st[2], s2);      // 2s*Vc at i, i-e_x, i-e_y, i-e_z
        put(4, i, s); put(4, i + st[1], s); put(4, i + st[0], s);                                 // s*Exy at i, i+e_y, i+e_x
        put(5, i, s); put(5, i + st[2], s); put(5, i + st[0], s);                                 // s*Exz at i, i+e_z, i+e_x
        put(6, i, s); put(6, i + st[2], s); put(6, i + st[1], s);                                 // s*Eyz at i, i+e_z, i+e_y
        const unsigned int a = (unsigned int)P.act[i] & kActCompute;
        auto diag = [&](auto Atag) {
            constexpr int A = decltype(Atag)::value;
            if (!(a & (1u << A))) return;
            T sum = T(0);
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                T hi, lo;
                if (ax == A) { hi = __ldg(P.coef[3] + i); lo = __ldg(P.coef[3] + i - st[A]); }
                else { const T* E = P.coef[3 + A + ax]; hi = __ldg(E + i + st[ax]); lo = __ldg(E + i); }
                const T h = (ax == A) ? Ar<true>::mul(T(2), hi) : hi;
                const T l = (ax == A) ? Ar<true>::mul(T(2), lo) : lo;
                sum = (ax == 0) ? Ar<true>::add(h, l) : Ar<true>::add(Ar<true>::add(sum, h), l);
            }
            cs[A * NL + i] = Ar<true>::add(__ldg(P.coef[A] + i), Ar<true>::mul(s, sum));
        };
        diag(std::integral_constant<int, 0>{});
        diag(std::integral_constant<int, 1>{});
        diag(std::integral_constant<int, 2>{});
    }
}

// ---------------------------------------------------------------------------------------------
// K1: CG-loop apply fused with d.q.  Inside the loop d is exactly zero on every row that is not
// computed (solid / boundary / padding / other slab / all-zero row), so neighbour masks are not
// needed (SURVEY A-1).
//
// Persistent grid (kK1BlocksPerSM CTAs per SM).  Each warp walks the sorted list of ACTIVE segments
// (32 consecutive lattice points, z contiguous across the warp) — the reference's `if sphi < 0:
// return` turned into "never launched": solid and liquid-free regions cost nothing.  The body is
// branch-free: all 27 neighbour values, 16 coefficients and the activity byte of a point are
// requested before the first use (one memory-latency period per segment; the next segment id is
// prefetched one trip ahead), and rows that are not computed select 0 and are not stored — they
// hold q == 0 since the masked apply at the start of the solve.  Loads of discarded lanes may fall
// outside the lattice; the workspace carries guard bands of one plane + one row + one element
// around the coefficient and vector regions for exactly that.  One block reduction at the very end
// (fixed order, deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int kK1Threads = 256;
constexpr int kK1TileDefault = 0;         // shared-memory tiled K1s on dense lattices when "k1_tile" / FLUIDSOLVER_B200_K1TILE is not set
constexpr int kK1BlockDefault = 4;        // consecutive trips per CTA block in the stand-alone K1s on HBM-sized lists (0 = interleaved)
constexpr int kK1SegsPerBlock = kK1Threads / 32;
// CTAs per SM: the fp64 body keeps 43 loaded values (86 registers) in flight, so it gets 128 registers per thread
// (2 CTAs/SM); with an 80-register cap (3 CTAs/SM) ptxas split the loads into dependent phases and the dense-scene
// K1 ran at 0.556 ms instead of 0.357 ms.  fp32 needs half the registers.
template <typename T> struct K1Occ { static constexpr int value = 2; };
template <> struct K1Occ<float> { static constexpr int value = 3; };

// One pass over the active segments: q = A d on computed rows, returns this thread's share of d.q.
// COHERENT = false: d is read through the read-only path (stand-alone K1: nothing writes d during the launch);
// COHERENT = true : plain loads (persistent kernel: other CTAs rewrote d before the last grid barrier).
// SR (single-reduction CG): d is the residual r, q the buffer w = A r; acc2 additionally collects r.r over the computed
// rows, and the boundary rows go to the neighbours' w planes instead of their q planes.
template <typename T, bool DIST, bool COHERENT, bool SR = false>
__device__ __forceinline__ double visc3d_apply_dot_body(const Visc3Dev<T>& P, T s, T s2, const T* d, T* q, const int* __restrict__ seg, int nseg,
                                                        const PeerHot& hot, bool& wrote_peer, double* acc2_out = nullptr, int par = 0, int blk = 0) {
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const long long st[3] = {L.sx, L.sy, 1};
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0, acc2 = 0.0;
    auto nb = [&](int comp, long long j) -> T { return COHERENT ? d[comp * NL + j] : __ldg(d + comp * NL + j); };
    // Warp -> list position of trip q.  blk == 0: interleaved over the whole grid (warp g takes g, g + nw, ...).
    // blk > 0 (HBM-sized lists): a CTA takes `blk` consecutive trips of blockDim/32 CONSECUTIVE list entries — consecutive
    // lattice rows — before it jumps ahead by the grid: the y-neighbour rows of one trip are the own rows of the next, so
    // they are served by L1 instead of L2 (the stand-alone apply runs at the L2 throughput cap, not at the HBM one), while
    // all CTAs still sweep the same few x-planes together (x-neighbour reuse through L2, DRAM traffic unchanged).
    const int wl = threadIdx.x >> 5, wc = blockDim.x >> 5;
    const long long span = (long long)blk * wc;
    auto kof = [&](long long trip) -> long long {
        if (blk <= 0) return w0 + trip * nw;
        const long long m = trip / blk, t = trip - m * blk;
        return (m * gridDim.x + blockIdx.x) * span + t * wc + wl;
    };
    long long k = kof(0);
    int sg_n = k < nseg ? __ldg(seg + k) : 0;
    for (long long trip = 0; k < nseg; ++trip) {
        const long long i = (long long)sg_n * kSegPts + lane;
        k = kof(trip + 1);                         // (strictly increasing in trip: the loop ends at the first position past the list)
        sg_n = k < nseg ? __ldg(seg + k) : 0;
        const bool in = i < NL;
        const long long j = in ? i : (NL - 1);
        const unsigned int a = in ? ((unsigned int)__ldg(P.act + i) & kActCompute) : 0u;
        const bool au = a & 1u, av = a & 2u, aw = a & 4u;
        const T cu = __ldg(P.cs[0] + j), cv = __ldg(P.cs[1] + j), cw = __ldg(P.cs[2] + j);      // row diagonals
        const T du = nb(0, j), dv = nb(1, j), dw = nb(2, j);
        const T ru = visc_row_scaled<T, 3, 0>(P.cs, j, st, cu, du, nb);
        const T rv = visc_row_scaled<T, 3, 1>(P.cs, j, st, cv, dv, nb);
        const T rw = visc_row_scaled<T, 3, 2>(P.cs, j, st, cw, dw, nb);
        if (au) { q[i] = ru; acc += (double)du * (double)ru; if (SR) acc2 += (double)du * (double)du; }
        if (av) { q[NL + i] = rv; acc += (double)dv * (double)rv; if (SR) acc2 += (double)dv * (double)dv; }
        if (aw) { q[2 * NL + i] = rw; acc += (double)dw * (double)rw; if (SR) acc2 += (double)dw * (double)dw; }
        if (DIST && a != 0u) {
            // multi-GPU: my first / last owned planes are the neighbours' halo planes of q — store them straight
            // into the peers' memory over NVLink; the all-reduce that follows publishes them.  (The local halo
            // planes 0 and X-2 carry no computed row: their q is written by the NEIGHBOURS, never by this rank.)
            const long long lo0 = L.sx, hi0 = (long long)(L.X - 3) * L.sx;
            char* const* plo = SR ? hot.w_lo + 3 * par : hot.q_lo;
            char* const* phi = SR ? hot.w_hi + 3 * par : hot.q_hi;
            if (hot.has_lo && i >= lo0 && i < lo0 + L.sx) {
                const long long o = i - lo0;
                if (au) reinterpret_cast<T*>(plo[0])[o] = ru;
                if (av) reinterpret_cast<T*>(plo[1])[o] = rv;
                if (aw) reinterpret_cast<T*>(plo[2])[o] = rw;
                wrote_peer = true;
            }
            if (hot.has_hi && i >= hi0 && i < hi0 + L.sx) {
                const long long o = i - hi0;
                if (au) reinterpret_cast<T*>(phi[0])[o] = ru;
                if (av) reinterpret_cast<T*>(phi[1])[o] = rv;
                if (aw) reinterpret_cast<T*>(phi[2])[o] = rw;
                wrote_peer = true;
            }
        }
    }
    if (SR) *acc2_out = acc2;
    return acc;
}

template <typename T, bool DIST>
__global__ void __launch_bounds__(kK1Threads, K1Occ<T>::value) visc3d_apply_dot_kernel(Visc3Dev<T> P, T s, T s2, const T* __restrict__ d, T* __restrict__ q,
                                                                                       const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                                       CgState* st_, double* partials, PeerInfo* peers, PeerHot hot) {
    if (*(volatile int*)&st_->done) return;
    bool wrote_peer = false;
    const double acc = visc3d_apply_dot_body<T, DIST, false>(P, s, s2, d, q, seg, *nseg_p, hot, wrote_peer);
    const bool block_wrote_peer = DIST ? (__syncthreads_or(wrote_peer ? 1 : 0) != 0) : false;
    grid_sum_finish(acc, partials, &st_->counter[0], [=](double sum) { st_->dq = sum; }, DIST ? peers : nullptr, 0, block_wrote_peer);
}

// ---------------------------------------------------------------------------------------------
// Whole CG iterations in ONE cooperative launch: K1, K2 and K3 become phases of a persistent kernel
// (one 512-thread CTA per SM) separated by grid barriers; the two reductions ride on the barriers.
// No launch gaps and no kernel ramp-up/tail per phase — what bounds an iteration once the active
// set is small (L2-resident scenes, 64^3, thin multi-GPU slabs).  Same arithmetic, same reduction
// tree shape (per-thread -> block -> fixed-order sum of block partials), same device-side
// convergence logic as the three-kernel path; runs up to n_iters iterations and stops early when
// the state says done.
// ---------------------------------------------------------------------------------------------
template <typename T, bool DIST>
__global__ void __launch_bounds__(kPersistThreads, 1) visc3d_cg_persistent_kernel(Visc3Dev<T> P, T s, T s2, T* x, T* r, T* d, T* q,
                                                                                  const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                                  CgState* st, double* partials, GridBar* bar, int n_iters,
                                                                                  PeerInfo* peers, PeerHot hot, unsigned long long* prof) {
    const int nseg = *nseg_p;
    const long long NL = P.L.NL;
    // The CG scalars live in registers, replicated in every thread of the grid: each block derives them from the same
    // per-block partials in the same order, so they stay bit-identical everywhere and nothing global is read in the loop.
    double delta = st->delta, delta_old = st->delta_old, dq = st->dq, alpha_d = st->alpha, beta_d = st->beta;
    const double tol2 = st->tol2;
    long long iter = st->iter;
    const long long max_iter = st->max_iter;
    int done = st->done;
    GridSync gs{bar, 0u};
    // multi-GPU: running sequence numbers of the two cross-GPU reductions (every block keeps its own copy; block 0 writes
    // them back at the end so that the next launch — persistent or three-kernel — continues the count)
    unsigned int seq0 = 0, seq1 = 0;
    if (DIST) { seq0 = peers->seq[0]; seq1 = peers->seq[1]; }
    // optional phase timeline (FLUIDSOLVER_B200_PROFILE): block 0 stamps the global timer after every phase / barrier
    const bool stamp = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    int np = 0;
    auto tick = [&]() {
        if (stamp && np < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); prof[np++] = t; }
    };
    for (int it = 0; it < n_iters && !done; ++it) {
        tick();
        // K1 phase: q = A d, d.q
        bool wrote_peer = false;
        double acc = visc3d_apply_dot_body<T, DIST, true>(P, s, s2, d, q, seg, nseg, hot, wrote_peer);
        const bool block_wrote_peer = DIST ? (__syncthreads_or(wrote_peer ? 1 : 0) != 0) : false;
        tick();
        if (DIST) ++seq0;
        dq = grid_allreduce(acc, partials, gs, DIST ? peers : nullptr, 0, block_wrote_peer, seq0);
        tick();
        // K2 phase: x += alpha d, r -= alpha q, r.r
        alpha_d = delta / dq;
        acc = cg_update_xr_seg_body<T, 3, DIST>(NL, NL, seg, nseg, x, r, d, q, (T)alpha_d, hot);
        tick();
        if (DIST) ++seq1;
        const double rr = grid_allreduce(acc, partials, gs, DIST ? peers : nullptr, 1, false, seq1);
        tick();
        delta_old = delta;
        delta = rr;
        iter += 1;
        if (rr < tol2) done = 1;
        else if (iter >= max_iter || !(rr == rr)) done = 2;      // NaN: the reference would spin to max_iter
        if (done) break;
        // K3 phase: d = r + beta d
        beta_d = delta / delta_old;
        cg_update_d_seg_body<T, 3>(NL, NL, seg, nseg, d, r, (T)beta_d);
        tick();
        gs.sync();
        tick();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->delta = delta; st->delta_old = delta_old; st->dq = dq; st->alpha = alpha_d; st->beta = beta_d;
        st->iter = iter; st->done = done;
        if (DIST) { peers->seq[0] = seq0; peers->seq[1] = seq1; }
    }
}

// ---------------------------------------------------------------------------------------------
// Single-reduction CG (see cg_sr_scalars in fs_common.cuh).  K1s: w = A r fused with BOTH dot products (r.r, w.r);
// its tail computes alpha and beta on the device.  The update p, s, x, r follows in one fused pass (K2s,
// cg_update_sr_seg_kernel).  Two kernels and ONE reduction per iteration instead of three and two; the persistent
// form below has two grid barriers per iteration (one carrying the reduction) instead of three (two carrying one).
// ---------------------------------------------------------------------------------------------
template <typename T, bool DIST, int OCC = K1Occ<T>::value>
__global__ void __launch_bounds__(kK1Threads, OCC) visc3d_apply_dot2_kernel(Visc3Dev<T> P, T s, T s2, const T* __restrict__ r, T* __restrict__ w /*[2][3][NL]*/,
                                                                                        const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                                        CgState* st_, double* partials, PeerInfo* peers, PeerHot hot, int freeze, int blk) {
    if (*(volatile int*)&st_->done) return;
    bool wrote_peer = false;
    double rr = 0.0;
    const int par = DIST ? (int)(st_->iter & 1) : 0;           // (st->iter is stable here: only the update kernel advances it)
    w += (long long)par * 3 * P.L.NL;
    const double wr = visc3d_apply_dot_body<T, DIST, false, true>(P, s, s2, r, w, seg, *nseg_p, hot, wrote_peer, &rr, par, blk);
    const bool block_wrote_peer = DIST ? (__syncthreads_or(wrote_peer ? 1 : 0) != 0) : false;
    grid_sum2_finish(rr, wr, partials, &st_->counter[0], [=](double gamma, double dl) {
        if (freeze) return;                      // profiling hook: repeated launches leave the CG state alone
        cg_sr_after_dots(st_, gamma, dl, par);
    }, (DIST && !freeze) ? peers : nullptr, block_wrote_peer);
}

// ---------------------------------------------------------------------------------------------
// K1t: the stand-alone single-reduction apply for DENSE lattices as a shared-memory tiled kernel that marches along x.
//
// The list kernel above fetches, per 32-point segment, 17 row-lines of r and 13 of coefficients from L2 (a third of its 43
// loads hit in L1): on a dense 256^3 lattice that is 4.4 GB through the L2 -> SM crossbar per launch against 1.75 GB of
// algorithmic (= DRAM) bytes, and the kernel sits at the measured crossbar limit (~6 300 B/clk for the chip), not at the
// HBM limit.  Here a CTA owns `ty` consecutive lattice rows — a CONTIGUOUS range of `tile = ty * Zp` points in every
// x-plane, for every array alike — and walks `xl` planes of them, thread t owning point t of the range in every plane.
// What a plane step reads from other points is staged in shared memory by cp.async (16-byte copies of contiguous ranges,
// one commit group per step, issued kK1tDepth steps ahead: a single step of these short row blocks is shorter than the
// DRAM latency):
//     r      3 components x planes x-1, x, x+1: the range plus one row + one element on either side, which covers every
//            in-plane stencil offset                                                     ring of 3 + depth, tile + 2H each
//     2sVc   planes x-1, x; read at i, i-sx, i-sy, i-1            -> halo before the range    ring of 2 + depth, H + tile
//     sExy   planes x, x+1; read at i, i+sy, i+sx                 -> halo after               ring of 2 + depth, tile + H
//     sExz   planes x, x+1; read at i, i+1, i+sx                  -> one element after        ring of 2 + depth, tile + 4
//     sEyz   plane x;       read at i, i+1, i+sy                  -> halo after               ring of 1 + depth, tile + H
// The three row diagonals and the activity byte of the thread's own point are requested one step ahead into registers.
// A plane step therefore waits for no global load: 27 + 13 shared-memory reads at compile-time offsets, 57 fp64
// operations and up to three stores per point.  Work items (row block, chunk of `xl` planes) are handed out by an atomic
// counter, x-chunk major, so the CTAs running together cover neighbouring row blocks of the same planes and the halo rows
// are shared through L2 (DRAM traffic stays the algorithmic 13 words per point).  Arithmetic per point is that of the list
// kernel (same evaluator, same order); only the reduction tree differs.
// ---------------------------------------------------------------------------------------------
constexpr int kK1tMaxThreads = 544;       // >= the 2 x 260 points of a two-row block of a 256^3 lattice; 120 registers per thread
constexpr int kK1tPlanes = 16;            // x-planes per work item (two warm-up planes of r are loaded on top)
constexpr int kK1tDepth = 2;              // plane steps between the issue of a copy and its first use

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ constexpr int k1t_halo(int Zp) { return Zp + 4; }     // one row + one element, rounded to 16 bytes
// dynamic shared memory of one CTA (elements as listed above)
__host__ __device__ inline size_t k1t_smem_bytes(long long tile, long long H, size_t esz) {
    constexpr int D = kK1tDepth;
    return esz * (size_t)(3 * (3 + D) * (tile + 2 * H) + 2 * (2 + D) * (tile + H) + (2 + D) * (tile + 4) + (1 + D) * (tile + H));
}

template <typename T>
__global__ void __launch_bounds__(kK1tMaxThreads, 1) visc3d_apply_dot2_tile_kernel(Visc3Dev<T> P, const T* __restrict__ r, T* __restrict__ w,
                                                                                  CgState* st_, double* partials, int ty, int xl, int freeze) {
    if (*(volatile int*)&st_->done) return;
    extern __shared__ __align__(16) unsigned char tile_smem[];
    __shared__ int s_item;
    constexpr int D = kK1tDepth, NR = 3 + D, NC = 2 + D, N6 = 1 + D;
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const int Zp = L.Zp;
    constexpr int VEC = 16 / (int)sizeof(T);
    const int H = k1t_halo(Zp);
    const int tile = ty * Zp;                                   // lattice points of one plane step (<= blockDim.x)
    const int Lr = tile + 2 * H, Lc = tile + H, L5 = tile + 4;
    T* const s_r = reinterpret_cast<T*>(tile_smem);             // [NR][3][Lr], point t at +H
    T* const s_c3 = s_r + 3 * NR * Lr;                          // [NC][Lc],    point t at +H
    T* const s_c4 = s_c3 + NC * Lc;                             // [NC][Lc]
    T* const s_c5 = s_c4 + NC * Lc;                             // [NC][L5]
    T* const s_c6 = s_c5 + NC * L5;                             // [N6][Lc]
    const int nthr = blockDim.x, t = threadIdx.x;
    auto stage = [&](T* dst, const T* src, int n) {             // n elements (a multiple of VEC), both sides 16-byte aligned
        for (int e = t * VEC; e < n; e += nthr * VEC) cp_async16(dst + e, src + e);
    };
    const int nyb = (L.Y + ty - 1) / ty;
    const int x_first = 1, x_end = L.nx;                        // planes 1 .. nx-1 can hold computed rows
    const int nxc = (x_end - x_first + xl - 1) / xl;
    const int nitems = nyb * nxc;
    double rr = 0.0, wr = 0.0;
    for (;;) {
        __syncthreads();                                        // everybody is done with the previous item (its buffers, s_item)
        if (t == 0) s_item = (int)atomicAdd(&st_->counter[3], 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= nitems) break;
        const int xc = item / nyb, yb = item - xc * nyb;
        const int xa = x_first + xc * xl;
        const int xb = xa + xl < x_end ? xa + xl : x_end;       // planes [xa, xb)
        const long long row0 = (long long)yb * tile;            // offset of the row block inside a plane
        const bool mine = t < tile && row0 + t < L.sx;          // (the last row block of a plane may overhang the lattice)
        auto issue_r = [&](int p) {
            const long long base = (long long)p * L.sx + row0 - H;
#pragma unroll
            for (int c = 0; c < 3; ++c) stage(s_r + (size_t)((p % NR) * 3 + c) * Lr, r + c * NL + base, Lr);
        };
        auto issue_c3 = [&](int p) { stage(s_c3 + (size_t)(p % NC) * Lc, P.cs[3] + (long long)p * L.sx + row0 - H, Lc); };
        auto issue_c45 = [&](int p) {
            const long long base = (long long)p * L.sx + row0;
            stage(s_c4 + (size_t)(p % NC) * Lc, P.cs[4] + base, Lc);
            stage(s_c5 + (size_t)(p % NC) * L5, P.cs[5] + base, L5);
        };
        auto issue_c6 = [&](int p) { stage(s_c6 + (size_t)(p % N6) * Lc, P.cs[6] + (long long)p * L.sx + row0, Lc); };
        // copy group of step x: what step x + D needs on top of step x + D - 1  (planes inside this item only)
        auto issue_step = [&](int x) {
            if (x + 1 + D <= xb) { issue_r(x + 1 + D); issue_c45(x + 1 + D); }
            if (x + D < xb) { issue_c3(x + D); issue_c6(x + D); }
            cp_async_commit();                                  // (possibly empty: keeps one group per step)
        };
        issue_r(xa - 1); issue_r(xa); issue_c45(xa); issue_c3(xa - 1);
        for (int k = -D; k < 0; ++k) issue_step(xa + k);        // groups of the D steps before the first one
        // own-point values straight into registers, one step ahead
        long long i = (long long)xa * L.sx + row0 + t;
        unsigned int a_cur = mine ? ((unsigned int)__ldg(P.act + i) & kActCompute) : 0u;
        T cd_cur[3] = {T(0), T(0), T(0)};
        if (a_cur) { cd_cur[0] = __ldg(P.cs[0] + i); cd_cur[1] = __ldg(P.cs[1] + i); cd_cur[2] = __ldg(P.cs[2] + i); }
        for (int x = xa; x < xb; ++x, i += L.sx) {
            issue_step(x);
            unsigned int a_nxt = 0u;
            T cd_nxt[3] = {T(0), T(0), T(0)};
            if (mine && x + 1 < xb) {
                a_nxt = (unsigned int)__ldg(P.act + i + L.sx) & kActCompute;
                cd_nxt[0] = __ldg(P.cs[0] + i + L.sx); cd_nxt[1] = __ldg(P.cs[1] + i + L.sx); cd_nxt[2] = __ldg(P.cs[2] + i + L.sx);
            }
            cp_async_wait<D>();                                 // all but the newest D groups have landed: everything step x reads
            __syncthreads();
            if (a_cur) {
                const T* rm = s_r + (size_t)(((x - 1) % NR) * 3) * Lr + H;
                const T* r0 = s_r + (size_t)((x % NR) * 3) * Lr + H;
                const T* rp = s_r + (size_t)(((x + 1) % NR) * 3) * Lr + H;
                const T* c3m = s_c3 + (size_t)((x - 1) % NC) * Lc + H;
                const T* c30 = s_c3 + (size_t)(x % NC) * Lc + H;
                const T* c40 = s_c4 + (size_t)(x % NC) * Lc;
                const T* c4p = s_c4 + (size_t)((x + 1) % NC) * Lc;
                const T* c50 = s_c5 + (size_t)(x % NC) * L5;
                const T* c5p = s_c5 + (size_t)((x + 1) % NC) * L5;
                const T* c60 = s_c6 + (size_t)(x % N6) * Lc;
                auto cf = [&](int plane, int axis, int sign) -> T {      // all three arguments are compile-time after unrolling
                    if (plane == 3) return sign == 0 ? c30[t] : (axis == 0 ? c3m[t] : (axis == 1 ? c30[t - Zp] : c30[t - 1]));
                    if (plane == 4) return sign == 0 ? c40[t] : (axis == 1 ? c40[t + Zp] : c4p[t]);
                    if (plane == 5) return sign == 0 ? c50[t] : (axis == 2 ? c50[t + 1] : c5p[t]);
                    return sign == 0 ? c60[t] : (axis == 2 ? c60[t + 1] : c60[t + Zp]);
                };
                auto nbu = [&](int comp, int p, int m) -> T {   // component comp at i + e_p - e_m
                    const int dx = (p == 0 ? 1 : 0) - (m == 0 ? 1 : 0);
                    const int off = ((p == 1 ? 1 : 0) - (m == 1 ? 1 : 0)) * Zp + ((p == 2 ? 1 : 0) - (m == 2 ? 1 : 0));
                    const T* b = dx > 0 ? rp : (dx < 0 ? rm : r0);
                    return b[comp * Lr + t + off];
                };
                const T du = r0[t], dv = r0[Lr + t], dw = r0[2 * Lr + t];
                const T ru = visc_row_scaled_u<T, 3, 0>(cd_cur[0], du, cf, nbu);
                const T rv = visc_row_scaled_u<T, 3, 1>(cd_cur[1], dv, cf, nbu);
                const T rw = visc_row_scaled_u<T, 3, 2>(cd_cur[2], dw, cf, nbu);
                if (a_cur & 1u) { w[i] = ru; wr += (double)du * (double)ru; rr += (double)du * (double)du; }
                if (a_cur & 2u) { w[NL + i] = rv; wr += (double)dv * (double)rv; rr += (double)dv * (double)dv; }
                if (a_cur & 4u) { w[2 * NL + i] = rw; wr += (double)dw * (double)rw; rr += (double)dw * (double)dw; }
            }
            a_cur = a_nxt;
            cd_cur[0] = cd_nxt[0]; cd_cur[1] = cd_nxt[1]; cd_cur[2] = cd_nxt[2];
            __syncthreads();                                    // the next step's copies overwrite what this step read last
        }
    }
    grid_sum2_finish(rr, wr, partials, &st_->counter[0], [=](double gamma, double dl) {
        st_->counter[3] = 0;                                    // work counter of the next launch (every block has left its loop)
        if (freeze) return;
        cg_sr_after_dots(st_, gamma, dl, 0);
    });
}

template <typename T, bool DIST>
__global__ void __launch_bounds__(kPersistThreads, 1) visc3d_cg_sr_persistent_kernel(Visc3Dev<T> P, T s, T s2, T* x, T* r, T* p, T* sv, T* w,
                                                                                     const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                                     CgState* st, double* partials, GridBar* bar, int n_slots,
                                                                                     PeerInfo* peers, PeerHot hot, unsigned long long* prof) {
    const int nseg = *nseg_p;
    const long long NL = P.L.NL;
    double delta = st->delta, gamma_old = st->delta_old, dl = st->dq, alpha_d = st->alpha, beta_d = st->beta;
    bool first = st->sr_first != 0;
    const double tol2 = st->tol2;
    long long iter = st->iter;
    const long long max_iter = st->max_iter;
    int done = st->done;
    GridSync gs{bar, 0u};
    unsigned int seq0 = 0, seq1 = 0;
    if (DIST) { seq0 = peers->seq[0]; seq1 = peers->seq[1]; }
    const bool stamp = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    int np = 0;
    auto tick = [&]() {
        if (stamp && np < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); prof[np++] = t; }
    };
    for (int it = 0; it < n_slots && !done; ++it) {
        tick();
        // phase A: w = A r, (r.r, w.r)
        bool wrote_peer = false;
        double rr = 0.0;
        const int par = DIST ? (int)(iter & 1) : 0;              // w is double-buffered by iteration parity (see PeerHot)
        T* const wk = w + (long long)par * 3 * NL;
        double wr = visc3d_apply_dot_body<T, DIST, true, true>(P, s, s2, r, wk, seg, nseg, hot, wrote_peer, &rr, par);
        const bool block_wrote_peer = DIST ? (__syncthreads_or(wrote_peer ? 1 : 0) != 0) : false;
        tick();
        if (DIST) { ++seq0; ++seq1; }
        grid_allreduce2(rr, wr, partials, gs, DIST ? peers : nullptr, block_wrote_peer, seq0, seq1);
        tick();
        delta = rr;                               // r.r of the residual after `iter` updates
        if (rr < tol2) done = 1;
        else if (iter >= max_iter || !(rr == rr)) done = 2;
        if (done) break;
        dl = wr;
        {
            double a, b;
            cg_sr_scalars(rr, wr, gamma_old, alpha_d, first, a, b);
            alpha_d = a; beta_d = b;
        }
        gamma_old = rr;
        first = false;
        // phase B: p = r + beta p, s = w + beta s, x += alpha p, r -= alpha s   (halo rows included: they mirror the owner's)
        cg_update_sr_seg_body<T, 3>(NL, NL, seg, nseg, x, r, p, sv, wk, (T)alpha_d, (T)beta_d);
        iter += 1;
        tick();
        gs.sync();
        tick();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->delta = delta; st->delta_old = gamma_old; st->dq = dl; st->alpha = alpha_d; st->beta = beta_d;
        st->iter = iter; st->done = done; st->sr_first = first ? 1 : 0;
        if (DIST) { peers->seq[0] = seq0; peers->seq[1] = seq1; }
    }
}

// ---------------------------------------------------------------------------------------------
// Single-reduction CG with the POINT-PRIVATE data resident in shared memory.  In this recurrence only r is ever read by
// another thread (the stencil of the apply); p, s, w, x and the 16 pre-scaled coefficient values of a point are touched by
// the lane that owns the point and by nobody else.  Every CTA owns a contiguous run of the (lattice-ordered) active list —
// the y / z neighbours of its segments are mostly its own segments and hit in L1 during phase A — and position li of the
// run belongs to warp li % 16, so all of that stays in shared memory for the whole launch, together with the residual of
// the own points and the segment ids: 31 values per point (16 coefficients, p, s, w, x, r) + the activity byte + the id =
// 7 972 B per fp64 segment.  Slots are addressed by the position of a segment in the CTA's run, so exactly `chunk` slots are
// needed: 29 for the 4 253 segments of the 256^3 benchmark scene = 231 KB of the 227 KiB an SM offers.  Phase A reads the r
// neighbourhood from L2 / L1; phase B touches global memory only to STORE the new r for the neighbours' stencils — no
// dependent L2 read in it (a first form of this kernel kept r and the segment ids in global memory and spent two sequential
// round trips there per trip: 8.1 instead of 7.6 us per iteration).  Positions beyond `res_slots` (longer lists) run
// through global memory exactly like visc3d_cg_sr_persistent_kernel.  p, s, x are loaded at kernel entry and written back
// at exit (once per launch of up to 256 iterations), so the global state between launches is unchanged.
// ---------------------------------------------------------------------------------------------
constexpr int kRes2Vals = 31;     // per point: 16 coefficients, p[3], s[3], w[3], x[3], r[3]
constexpr int kResidentFormDefault = 2;   // 0 = persistent CG through global memory, non-zero = shared-memory resident kernel ("resident_form" / FLUIDSOLVER_B200_RESIDENT)
constexpr int kRes2Threads = 512;
template <typename T> __host__ __device__ constexpr size_t res2_slot_bytes() { return (size_t)kRes2Vals * 32 * sizeof(T) + 32 + sizeof(int); }

template <typename T>
__global__ void __launch_bounds__(kRes2Threads, 1) visc3d_cg_sr_resident2_kernel(Visc3Dev<T> P, T* x, T* r, T* p, T* sv, T* w,
                                                                                        const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                                        CgState* st, double* partials, GridBar* bar, int n_slots, int res_slots,
                                                                                        unsigned long long* prof) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    constexpr int kWarps = kRes2Threads / 32;
    T* const svals = reinterpret_cast<T*>(res_smem);                                            // [res_slots][kRes2Vals][32]
    uint8_t* const sact = res_smem + (size_t)res_slots * kRes2Vals * 32 * sizeof(T);            // [res_slots][32]
    int* const sseg = reinterpret_cast<int*>(sact + (size_t)res_slots * 32);                    // [res_slots]
    const Lat3& L = P.L;
    const long long NL = L.NL;
    const long long stv[3] = {L.sx, L.sy, 1};
    const int nseg = *nseg_p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // contiguous run of the (lattice-ordered) active list per CTA; position li of the run belongs to warp li % kWarps
    const long long chunk = ((long long)nseg + gridDim.x - 1) / gridDim.x;
    const long long c_lo = (long long)blockIdx.x * chunk;
    const int n_own = (int)((c_lo + chunk < nseg ? c_lo + chunk : (long long)nseg) - c_lo);     // (<= 0 for CTAs past the end of the list)

    // ---- prologue: point-private data of the resident positions -> shared memory
    for (int li = warp; li < n_own && li < res_slots; li += kWarps) {
        const int sg = __ldg(seg + c_lo + li);
        const long long i = (long long)sg * kSegPts + lane;
        const bool in = i < NL;
        const long long jj = in ? i : (NL - 1);
        T* S = svals + ((size_t)li * kRes2Vals) * 32 + lane;
        S[0 * 32] = __ldg(P.cs[0] + jj); S[1 * 32] = __ldg(P.cs[1] + jj); S[2 * 32] = __ldg(P.cs[2] + jj);
        S[3 * 32] = __ldg(P.cs[3] + jj);
        S[4 * 32] = __ldg(P.cs[3] + jj - stv[0]); S[5 * 32] = __ldg(P.cs[3] + jj - stv[1]); S[6 * 32] = __ldg(P.cs[3] + jj - stv[2]);
        S[7 * 32] = __ldg(P.cs[4] + jj); S[8 * 32] = __ldg(P.cs[4] + jj + stv[1]); S[9 * 32] = __ldg(P.cs[4] + jj + stv[0]);
        S[10 * 32] = __ldg(P.cs[5] + jj); S[11 * 32] = __ldg(P.cs[5] + jj + stv[2]); S[12 * 32] = __ldg(P.cs[5] + jj + stv[0]);
        S[13 * 32] = __ldg(P.cs[6] + jj); S[14 * 32] = __ldg(P.cs[6] + jj + stv[2]); S[15 * 32] = __ldg(P.cs[6] + jj + stv[1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            S[(16 + c) * 32] = in ? p[c * NL + i] : T(0);
            S[(19 + c) * 32] = in ? sv[c * NL + i] : T(0);
            S[(22 + c) * 32] = T(0);
            S[(25 + c) * 32] = in ? x[c * NL + i] : T(0);
            S[(28 + c) * 32] = in ? r[c * NL + i] : T(0);
        }
        sact[(size_t)li * 32 + lane] = in ? (uint8_t)((unsigned int)__ldg(P.act + i) & kActCompute) : (uint8_t)0;
        if (lane == 0) sseg[li] = sg;
    }
    __syncwarp();

    double delta = st->delta, gamma_old = st->delta_old, dl = st->dq, alpha_d = st->alpha, beta_d = st->beta;
    bool first = st->sr_first != 0;
    const double tol2 = st->tol2;
    long long iter = st->iter;
    const long long max_iter = st->max_iter;
    int done = st->done;
    GridSync gs{bar, 0u};
    const bool stamp = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    int np = 0;
    auto tick = [&]() {
        if (stamp && np < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); prof[np++] = t; }
    };
    auto nb = [&](int comp, long long j) -> T { return r[comp * NL + j]; };      // coherent: other CTAs rewrote r before the last barrier
    for (int it = 0; it < n_slots && !done; ++it) {
        tick();
        // ---- phase A: w = A r, (r.r, w.r)
        double rr = 0.0, wr = 0.0;
        for (int li = warp; li < n_own; li += kWarps) {
            const bool res = li < res_slots;
            const int sg = res ? sseg[li] : __ldg(seg + c_lo + li);
            const long long i = (long long)sg * kSegPts + lane;
            const bool in = i < NL;
            const long long jj = in ? i : (NL - 1);
            const T du = nb(0, jj), dv = nb(1, jj), dw = nb(2, jj);
            T ru, rv, rw;
            unsigned int a;
            if (res) {
                T* S = svals + ((size_t)li * kRes2Vals) * 32 + lane;
                a = sact[(size_t)li * 32 + lane];
                auto cf = [&](int plane, int axis, int sign) -> T { return S[visc3_cslot(plane, axis, sign) * 32]; };
                ru = visc_row_scaled_cf<T, 3, 0>(jj, stv, S[0 * 32], du, cf, nb);
                rv = visc_row_scaled_cf<T, 3, 1>(jj, stv, S[1 * 32], dv, cf, nb);
                rw = visc_row_scaled_cf<T, 3, 2>(jj, stv, S[2 * 32], dw, cf, nb);
                S[22 * 32] = (a & 1u) ? ru : T(0);          // w stays in shared memory (zero on rows that are not computed)
                S[23 * 32] = (a & 2u) ? rv : T(0);
                S[24 * 32] = (a & 4u) ? rw : T(0);
            } else {
                a = in ? ((unsigned int)__ldg(P.act + i) & kActCompute) : 0u;
                ru = visc_row_scaled<T, 3, 0>(P.cs, jj, stv, __ldg(P.cs[0] + jj), du, nb);
                rv = visc_row_scaled<T, 3, 1>(P.cs, jj, stv, __ldg(P.cs[1] + jj), dv, nb);
                rw = visc_row_scaled<T, 3, 2>(P.cs, jj, stv, __ldg(P.cs[2] + jj), dw, nb);
                if (a & 1u) w[i] = ru;
                if (a & 2u) w[NL + i] = rv;
                if (a & 4u) w[2 * NL + i] = rw;
            }
            if (a & 1u) { wr += (double)du * (double)ru; rr += (double)du * (double)du; }
            if (a & 2u) { wr += (double)dv * (double)rv; rr += (double)dv * (double)dv; }
            if (a & 4u) { wr += (double)dw * (double)rw; rr += (double)dw * (double)dw; }
        }
        tick();
        grid_allreduce2(rr, wr, partials, gs);
        tick();
        delta = rr;
        if (rr < tol2) done = 1;
        else if (iter >= max_iter || !(rr == rr)) done = 2;
        if (done) break;
        dl = wr;
        {
            double a, b;
            cg_sr_scalars(rr, wr, gamma_old, alpha_d, first, a, b);
            alpha_d = a; beta_d = b;
        }
        gamma_old = rr;
        first = false;
        // ---- phase B: p = r + beta p, s = w + beta s, x += alpha p, r -= alpha s
        {
            const T alpha = (T)alpha_d, beta = (T)beta_d;
            for (int li = warp; li < n_own; li += kWarps) {
                if (li < res_slots) {                 // everything from shared memory; the new r also goes out for the neighbours
                    const long long i = (long long)sseg[li] * kSegPts + lane;
                    const bool in = i < NL;
                    T* S = svals + ((size_t)li * kRes2Vals) * 32 + lane;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const T rc = S[(28 + c) * 32];
                        const T pn = rc + beta * S[(16 + c) * 32];
                        const T sn = S[(22 + c) * 32] + beta * S[(19 + c) * 32];
                        const T rn = rc - alpha * sn;
                        S[(16 + c) * 32] = pn;
                        S[(19 + c) * 32] = sn;
                        S[(25 + c) * 32] += alpha * pn;
                        S[(28 + c) * 32] = rn;
                        if (in) r[c * NL + i] = rn;
                    }
                } else {                              // beyond the resident positions: through global memory
                    const long long i = (long long)__ldg(seg + c_lo + li) * kSegPts + lane;
                    if (i >= NL) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const long long e = c * NL + i;
                        const T rc = r[e];
                        const T pn = rc + beta * p[e];
                        const T sn = w[e] + beta * sv[e];
                        p[e] = pn;
                        sv[e] = sn;
                        x[e] += alpha * pn;
                        r[e] = rc - alpha * sn;
                    }
                }
            }
        }
        iter += 1;
        tick();
        gs.sync();
        tick();
    }
    // ---- epilogue: the resident p, s, x go back to global memory (w is scratch, r is current there already)
    for (int li = warp; li < n_own && li < res_slots; li += kWarps) {
        const long long i = (long long)sseg[li] * kSegPts + lane;
        if (i >= NL) continue;
        const T* S = svals + ((size_t)li * kRes2Vals) * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p[c * NL + i] = S[(16 + c) * 32];
            sv[c * NL + i] = S[(19 + c) * 32];
            x[c * NL + i] = S[(25 + c) * 32];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->delta = delta; st->delta_old = gamma_old; st->dq = dl; st->alpha = alpha_d; st->beta = beta_d;
        st->iter = iter; st->done = done; st->sr_first = first ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Gathered multi-GPU solve (L2-sized active sets): every rank packs / loads / extrapolates only its own x-window of the
// GLOBAL lattice, then the ranks exchange just the lattice segments the CG touches — the active segments of the planes a
// rank owns plus every segment their stencils read — and each rank runs the whole (small) CG locally, with no
// per-iteration traffic between GPUs.  One record per published segment:
//     [int32 segment id, 12 bytes pad][mask_u[32] mask_v[32] mask_w[32] act[32]][7 coefficient + 3 x values][32 points]
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t gather_record_bytes(size_t esz) { return 16 + 128 + 320 * esz; }

// flags[t] = 1 for every segment t that holds a computed row of the owned planes, and for every segment such a row's
// stencil can read (row offsets 0, +-sy, +-sx, +sx-sy, -sx+sy with z-shifts -1..+1)
__global__ void __launch_bounds__(256) visc3d_mark_export_kernel(Lat3 L, const uint8_t* __restrict__ act, long long seg_lo, long long seg_hi,
                                                                 long long nseg_total, uint8_t* __restrict__ flags) {
    const long long sg = seg_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= seg_hi || !seg_flag(act, sg, nseg_total)) return;
    const long long off[7] = {0, L.sy, -L.sy, L.sx, -L.sx, L.sx - L.sy, L.sy - L.sx};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const long long lo = sg * kSegPts + off[k] - 1, hi = sg * kSegPts + off[k] + kSegPts;   // first / last point touched
        long long t0 = lo < 0 ? 0 : lo / kSegPts, t1 = hi / kSegPts;
        if (t1 >= nseg_total) t1 = nseg_total - 1;
        for (long long t = t0; t <= t1; ++t) flags[t] = 1;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_export_kernel(long long NL, const T* __restrict__ coef, const T* __restrict__ xv, const uint8_t* __restrict__ mask,
                                                                 const uint8_t* __restrict__ act, const int* __restrict__ seg, const int* __restrict__ nseg_p,
                                                                 long long cap, char* __restrict__ records) {
    const long long nseg = *nseg_p < cap ? *nseg_p : cap;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const size_t rb = gather_record_bytes(sizeof(T));
    for (long long k = w0; k < nseg; k += nw) {
        const int sg = __ldg(seg + k);
        char* rec = records + (size_t)k * rb;
        const long long i = (long long)sg * kSegPts + lane;
        const bool in = i < NL;
        if (lane == 0) *reinterpret_cast<int*>(rec) = sg;
        uint8_t* bytes = reinterpret_cast<uint8_t*>(rec + 16);
#pragma unroll
        for (int c = 0; c < 3; ++c) bytes[c * 32 + lane] = in ? mask[c * NL + i] : (uint8_t)0;
        bytes[96 + lane] = in ? act[i] : (uint8_t)0;
        T* vals = reinterpret_cast<T*>(rec + 16 + 128);
#pragma unroll
        for (int p = 0; p < 7; ++p) vals[p * 32 + lane] = in ? coef[p * NL + i] : T(0);
#pragma unroll
        for (int c = 0; c < 3; ++c) vals[(7 + c) * 32 + lane] = in ? xv[c * NL + i] : T(0);
    }
}

// records of rank r: [r*stride, r*stride + counts[r]); the local rank's own block is skipped (its lattice already holds the data)
template <typename T>
__global__ void __launch_bounds__(kThreads) visc3d_import_kernel(long long NL, T* __restrict__ coef, T* __restrict__ xv, uint8_t* __restrict__ mask, uint8_t* __restrict__ act,
                                                                 const char* __restrict__ records, const long long* __restrict__ counts, int nranks,
                                                                 long long stride, int skip_rank) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const size_t rb = gather_record_bytes(sizeof(T));
    for (long long k = w0; k < (long long)nranks * stride; k += nw) {
        const int r = (int)(k / stride);
        if (r == skip_rank || k - r * stride >= counts[r]) continue;
        const char* rec = records + (size_t)k * rb;
        const long long i = (long long)(*reinterpret_cast<const int*>(rec)) * kSegPts + lane;
        if (i >= NL) continue;
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(rec + 16);
#pragma unroll
        for (int c = 0; c < 3; ++c) mask[c * NL + i] = bytes[c * 32 + lane];
        act[i] = bytes[96 + lane];
        const T* vals = reinterpret_cast<const T*>(rec + 16 + 128);
#pragma unroll
        for (int p = 0; p < 7; ++p) coef[p * NL + i] = vals[p * 32 + lane];
#pragma unroll
        for (int c = 0; c < 3; ++c) xv[c * NL + i] = vals[(7 + c) * 32 + lane];
    }
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
    char* coefs;     // [7][NL] T pre-scaled coefficients of the CG-loop apply (valid on the active segments and the slots their rows read)
    char* vecs;      // [5][3][NL] T
    uint8_t* mask;   // [3][NL]
    uint8_t* valid;  // [3][NL] extrapolation validity generations
    GridBar* bar;    // grid barrier of the persistent CG kernel
    ExtrapWork work; // extrapolation work lists
    uint8_t* rowflag; // [X*Y] lattice row holds a fluid face (written by pack, read by extrapolation sweep 1)
    uint8_t* rownz;   // [X*Y] the row's coefficient planes hold a non-zero (pack skips rewriting rows that stay all-zero)
    char* d2;        // [2][3][NL] w = A r of the single-reduction CG (two iteration parities); zero outside the active segments like r,d,q,b
    int cg_mode;     // FS_CG_*
    bool coop_failed; // a cooperative launch was refused on this context: use the stand-alone kernels from now on
    bool resident_failed; // the shared-memory resident persistent kernel cannot run here (or is switched off)
    bool sparse_clean;   // r,d,q,b are zero outside the segments of the current active list (sparse begin / clear may be used)
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
    SegList seg;     // sorted list of active 32-point lattice segments (rebuilt by every pack)
    int active_mode; // FS_ACTIVE_*
    long long active_rows;   // computed rows of the last pack (host copy, filled lazily)
    // gathered multi-GPU solve / windowed inputs
    bool windowed;       // pack / load / extrapolation work on the x-window L.wlo..L.whi of the lattice only
    uint8_t* xflags;     // [segments] publish flags of the current solve
    SegList xseg;        // segments this rank publishes (built from xflags)
    int k1t_smem;        // dynamic shared memory the tiled K1s may use on this device (0: not available)
};

// threads of the one-block-per-lattice-row kernels: a whole number of warps covering the row once, at most 512
static int row_block(const Lat3& L) {
    int b = (L.Zp + 31) / 32 * 32;
    return b < 512 ? b : 512;
}

static Lat3 make_lat3(int nx, int ny, int nz) {
    Lat3 L;
    L.nx = nx; L.ny = ny; L.nz = nz;
    L.X = nx + 1; L.Y = ny + 1; L.Zp = (nz + 1 + 3) / 4 * 4;
    L.sy = L.Zp; L.sx = (long long)L.Y * L.Zp; L.NL = L.sx * L.X;
    L.u_xhi = nx - 1;
    L.has_lo = 0; L.has_hi = 0;
    L.wlo = 0; L.whi = nx; L.wcells = nx;
    return L;
}

struct Visc3Layout { size_t coefs, xflags, xlist, xscratch, rownz; size_t coef, vecs, mask, valid, act, partials, st, seglist, segscratch, bar, wlist, wcount, d2, rowflag, total; int grid_pts; unsigned int wcap; };

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
    o.valid = p; p = align_up(p + 3 * L.NL, 256);
    o.act = p; p = align_up(p + L.NL + 64, 256);
    size_t np = (size_t)(o.grid_pts > kVecGrid ? o.grid_pts : kVecGrid);
    o.partials = p; p = align_up(p + np * sizeof(double), 256);
    o.st = p; p = align_up(p + sizeof(CgState), 256);
    // appended after everything the peers address (fs_visc3d_set_peers derives the neighbours' q planes from `vecs`)
    o.seglist = p; p = align_up(p + SegList::list_bytes(L.NL), 256);
    o.segscratch = p; p = align_up(p + SegList::scratch_bytes(L.NL), 256);
    o.bar = p; p = align_up(p + sizeof(GridBar), 256);
    // extrapolation work lists: 2 x NL/2 face ids (a sweep that fills more falls back to full passes); needs 3*NL < 2^32
    o.wcap = (3 * L.NL < 0xffffffffLL) ? (unsigned int)(L.NL / 2 + 1024) : 0u;
    o.wlist = p; p = align_up(p + (size_t)2 * o.wcap * sizeof(unsigned int), 256);
    o.wcount = p; p = align_up(p + 4 * sizeof(unsigned int), 256);
    // w = A r buffer of the single-reduction CG, with the same guard bands as `vecs`
    p += guard;
    o.d2 = p; p = align_up(p + 6 * L.NL * esz, 256) + guard;       // two parities (multi-GPU double buffering)
    o.rowflag = p; p = align_up(p + (size_t)L.X * L.Y, 256);
    // pre-scaled coefficient planes of the CG-loop apply (same guard bands as `coef`: the branch-free K1 reads neighbours of discarded lanes)
    p += guard;
    o.coefs = p; p = align_up(p + 7 * L.NL * esz, 256) + guard;
    // gathered multi-GPU solve: publish flags (one byte per segment) and the list built from them
    o.xflags = p; p = align_up(p + (size_t)((L.NL + kSegPts - 1) / kSegPts) + 64, 256);
    o.xlist = p; p = align_up(p + SegList::list_bytes(L.NL), 256);
    o.xscratch = p; p = align_up(p + SegList::scratch_bytes(L.NL), 256);
    o.rownz = p; p = align_up(p + (size_t)L.X * L.Y, 256);
    o.total = p;
    return o;
}

template <typename T> static Visc3Dev<T> dev_view(const fs_visc3d* h) {
    Visc3Dev<T> P;
    P.L = h->L;
    for (int k = 0; k < 7; ++k) P.coef[k] = reinterpret_cast<const T*>(h->coef) + k * h->L.NL;
    for (int k = 0; k < 7; ++k) P.cs[k] = reinterpret_cast<const T*>(h->coefs) + k * h->L.NL;
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

static bool visc3d_use_persistent(const fs_visc3d* h);
static bool visc3d_use_sr(const fs_visc3d* h);

int fs_visc3d_set_slab(fs_visc3d* h, fs_comm* comm, int has_lo, int has_hi) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if ((has_lo || has_hi) && !comm) return fail(FS_ERR_ARG, "fs_visc3d_set_slab: neighbours need a communicator");
    if (h->L.nx < 2 + (has_lo ? 1 : 0) + (has_hi ? 1 : 0)) return fail(FS_ERR_ARG, "fs_visc3d_set_slab: slab too thin");
    h->comm = comm;
    h->has_lo = has_lo ? 1 : 0;
    h->has_hi = has_hi ? 1 : 0;
    h->L.u_xhi = h->L.nx - 1 - h->has_hi;
    h->L.has_lo = h->has_lo; h->L.has_hi = h->has_hi;
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
        for (int c = 0; c < 3; ++c) {
            pi.q_lo[c] = (char*)lo_ws + ln.vecs + (((size_t)FS_VEC_Q * 3 + c) * Ln.NL + (size_t)(Ln.X - 2) * Ln.sx) * esz;
            for (int par = 0; par < 2; ++par)
                h->hot.w_lo[par * 3 + c] = (char*)lo_ws + ln.d2 + ((size_t)(par * 3 + c) * Ln.NL + (size_t)(Ln.X - 2) * Ln.sx) * esz;
        }
    }
    if (h->has_hi) {
        Lat3 Ln = make_lat3(hi_nx, h->L.ny, h->L.nz);
        Visc3Layout ln = visc3_layout(Ln, esz);
        for (int c = 0; c < 3; ++c) {
            pi.q_hi[c] = (char*)hi_ws + ln.vecs + (((size_t)FS_VEC_Q * 3 + c) * Ln.NL) * esz;
            for (int par = 0; par < 2; ++par)
                h->hot.w_hi[par * 3 + c] = (char*)hi_ws + ln.d2 + ((size_t)(par * 3 + c) * Ln.NL) * esz;
        }
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

int fs_visc3d_debug_read(fs_visc3d* h, int what, void* out, size_t bytes) {
    if (!h || !out) return fail(FS_ERR_ARG, "null argument");
    const void* src = what == 0 ? (const void*)h->valid : (const void*)h->act;
    FS_CUDA(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return FS_OK;
}

int fs_visc3d_peer_error(fs_visc3d* h) {
    if (!h || !h->peers) return 0;
    PeerInfo pi;
    if (cudaMemcpy(&pi, h->peers, sizeof(pi), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return pi.error;
}

int fs_visc3d_set_cg_mode(fs_visc3d* h, int mode) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (mode != FS_CG_AUTO && mode != FS_CG_KERNELS && mode != FS_CG_PERSISTENT && mode != FS_CG_KERNELS_SR && mode != FS_CG_PERSISTENT_SR)
        return fail(FS_ERR_ARG, "fs_visc3d_set_cg_mode: bad mode");
    h->cg_mode = mode;
    h->graph.valid = false;
    return FS_OK;
}

int fs_visc3d_cg_mode_in_use(fs_visc3d* h) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_cg_mode_in_use before fs_visc3d_pack");
    FS_TRY(h->seg.finish());
    if (visc3d_use_sr(h)) return visc3d_use_persistent(h) ? FS_CG_PERSISTENT_SR : FS_CG_KERNELS_SR;
    return visc3d_use_persistent(h) ? FS_CG_PERSISTENT : FS_CG_KERNELS;
}

int fs_visc3d_set_active_mode(fs_visc3d* h, int mode) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (mode != FS_ACTIVE_FLUID && mode != FS_ACTIVE_NONZERO) return fail(FS_ERR_ARG, "fs_visc3d_set_active_mode: bad mode");
    h->active_mode = mode;
    h->packed = false;
    h->graph.valid = false;
    return FS_OK;
}

__global__ void visc3d_count_rows_kernel(const uint8_t* act, long long n, unsigned long long* out) {
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c += __popc((unsigned int)act[i] & fs::kActCompute);
    c = __reduce_add_sync(0xffffffffu, (unsigned int)c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int fs_visc3d_active_info(fs_visc3d* h, int64_t* segments, int64_t* segments_total, int64_t* rows, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_active_info before fs_visc3d_pack");
    FS_TRY(h->seg.finish());
    cudaStream_t s = (cudaStream_t)stream;
    if (rows) {
        if (h->active_rows < 0) {       // counted on demand; `partials` is free between solves
            unsigned long long* tmp = reinterpret_cast<unsigned long long*>(h->partials);
            FS_CUDA(cudaMemsetAsync(tmp, 0, sizeof(*tmp), s));
            visc3d_count_rows_kernel<<<kSMs * 4, 256, 0, s>>>(h->act, h->L.NL, tmp);
            FS_LAUNCH_CHECK();
            unsigned long long host = 0;
            FS_CUDA(cudaMemcpyAsync(&host, tmp, sizeof(host), cudaMemcpyDeviceToHost, s));
            FS_CUDA(cudaStreamSynchronize(s));
            h->active_rows = (long long)host;
        }
        *rows = h->active_rows;
    }
    if (segments) *segments = h->seg.nseg;
    if (segments_total) *segments_total = h->seg.nseg_total;
    return FS_OK;
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
    h->coef = h->ws + lay.coef; h->vecs = h->ws + lay.vecs; h->coefs = h->ws + lay.coefs;
    h->mask = (uint8_t*)(h->ws + lay.mask); h->valid = (uint8_t*)(h->ws + lay.valid); h->act = (uint8_t*)(h->ws + lay.act);
    h->partials = (double*)(h->ws + lay.partials); h->st = (CgState*)(h->ws + lay.st);
    h->bar = (GridBar*)(h->ws + lay.bar);
    h->d2 = h->ws + lay.d2;
    h->rowflag = (uint8_t*)(h->ws + lay.rowflag);
    h->rownz = (uint8_t*)(h->ws + lay.rownz);
    h->work.cap = lay.wcap;
    if (const char* e = getenv("FLUIDSOLVER_B200_EXTRAP_CAP")) {   // test hook: tiny lists force the overflow fall-back
        const long long c = atoll(e);
        if (c >= 0 && c < (long long)lay.wcap) h->work.cap = (unsigned int)c;
    }
    h->work.list[0] = (unsigned int*)(h->ws + lay.wlist);
    h->work.list[1] = h->work.list[0] + lay.wcap;
    h->work.count = (unsigned int*)(h->ws + lay.wcount);
    h->cg_mode = FS_CG_AUTO;
    h->coop_failed = false;
    h->resident_failed = false;
    h->sparse_clean = true;          // the workspace is zeroed below and the list is empty
    h->grid_pts = lay.grid_pts;
    h->packed = false;
    h->comm = nullptr; h->has_lo = 0; h->has_hi = 0; h->peers = nullptr;
    h->active_mode = FS_ACTIVE_NONZERO; h->active_rows = -1;
    memset(&h->hot, 0, sizeof(h->hot));
    int s = h->cg.init();
    if (s < 0) { delete h; return s; }
    s = h->seg.init(h->L.NL, h->ws + lay.seglist, h->ws + lay.segscratch);
    if (s < 0) { h->cg.destroy(); delete h; return s; }
    s = h->xseg.init(h->L.NL, h->ws + lay.xlist, h->ws + lay.xscratch);
    if (s < 0) { h->cg.destroy(); h->seg.destroy(); delete h; return s; }
    h->xflags = (uint8_t*)(h->ws + lay.xflags);
    h->k1t_smem = 0;
    {   // opt the tiled K1s into the large shared-memory carve-out now (not inside a stream capture later)
        int dev = 0, smem_max = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && smem_max > 2048) {
            cudaError_t ea = cudaSuccess;
            FS_DISPATCH(h, ea = cudaFuncSetAttribute((const void*)visc3d_apply_dot2_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
            if (ea == cudaSuccess) h->k1t_smem = smem_max - 1024;
        }
        cudaGetLastError();
    }
    h->windowed = false;
    h->cg.st_dev = h->st; h->cg.partials_dev = h->partials;
    cudaError_t e = cudaMemset(ws, 0, lay.total);     // cp.zeros semantics for every solver vector
    if (e != cudaSuccess) { h->cg.destroy(); h->seg.destroy(); h->xseg.destroy(); delete h; return fail(FS_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    *out = h;
    return FS_OK;
}

void fs_visc3d_destroy(fs_visc3d* h) {
    if (!h) return;
    if (h->peers) cudaFree(h->peers);
    h->graph.destroy();
    h->cg.destroy();
    h->seg.destroy();
    h->xseg.destroy();
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
    FS_TRY(h->seg.finish());              // (a previous pack whose list length was never consumed)
    // keep "r, d, q, b are zero outside the active segments": wipe the previous solve's active segments (or everything,
    // if a dense API call wrote into those vectors since)
    if (h->sparse_clean) {
        if (h->seg.nseg > 0) {
            const int grid = seg_grid(h->seg.nseg, kThreads / 32, kSMs * 8);
            FS_DISPATCH(h, visc3d_clear_kernel<T><<<grid, kThreads, 0, s>>>(h->L.NL, reinterpret_cast<T*>(h->vecs), reinterpret_cast<T*>(h->d2), h->seg.list, h->seg.nseg_dev,
                                                                            h->windowed ? h->act : nullptr));
            FS_LAUNCH_CHECK();
        }
    } else {
        FS_CUDA(cudaMemsetAsync(h->vecs + (size_t)FS_VEC_R * 3 * h->L.NL * h->esz, 0, (size_t)12 * h->L.NL * h->esz, s));
        FS_CUDA(cudaMemsetAsync(h->d2, 0, (size_t)6 * h->L.NL * h->esz, s));
        h->sparse_clean = true;
    }
    FS_DISPATCH(h, visc3d_pack_kernel<T><<<(h->L.whi - h->L.wlo + 1) * h->L.Y, row_block(h->L), 0, s>>>(h->L, sphi, lvol, vol_norm, reinterpret_cast<T*>(h->coef), h->mask, h->act, h->rowflag, h->rownz,
                                                                                      h->active_mode == FS_ACTIVE_NONZERO ? 1 : 0));
    FS_LAUNCH_CHECK();
    FS_TRY(h->seg.enqueue(h->act, s));    // the list length (read back asynchronously) sizes the CG launches: see visc3d_list_ready
    h->active_rows = -1;
    h->packed = true;
    return FS_OK;
}

int fs_visc3d_load(fs_visc3d* h, int vec, const void* vx, const void* vy, const void* vz, int src_dtype, void* stream) {
    if (!h || !vx || !vy || !vz) return fail(FS_ERR_ARG, "fs_visc3d_load: null argument");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc3d_load: bad vector id");
    if (vec != FS_VEC_X) h->sparse_clean = false;
    cudaStream_t s = (cudaStream_t)stream;
    if (src_dtype == FS_F32) {
        FS_DISPATCH(h, visc3d_load_kernel<T, float><<<(h->L.whi - h->L.wlo + 1) * h->L.Y, row_block(h->L), 0, s>>>(h->L, (const float*)vx, (const float*)vy, (const float*)vz, vec_ptr<T>(h, vec)));
    } else if (src_dtype == FS_F64) {
        FS_DISPATCH(h, visc3d_load_kernel<T, double><<<(h->L.whi - h->L.wlo + 1) * h->L.Y, row_block(h->L), 0, s>>>(h->L, (const double*)vx, (const double*)vy, (const double*)vz, vec_ptr<T>(h, vec)));
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

static int visc3d_extrapolate_impl(fs_visc3d* h, int vec, int sweeps, void* stream, const SegList* region /*sweep 1 only looks at these segments, or null*/) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_extrapolate: call fs_visc3d_pack first");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc3d_extrapolate: bad vector id");
    if (sweeps > 250) return fail(FS_ERR_ARG, "fs_visc3d_extrapolate: at most 250 sweeps");
    if (sweeps <= 0) return FS_OK;
    if (vec != FS_VEC_X) h->sparse_clean = false;
    cudaStream_t s = (cudaStream_t)stream;
    const long long NL3 = 3 * h->L.NL;
    // x-planes swept: the whole lattice, or — windowed inputs — the window minus its edge planes (a sweep reads x+-1)
    const Lat3& L = h->L;
    const int p_lo = L.wlo > 0 ? L.wlo + 1 : 0, p_hi = L.whi < L.nx ? L.whi - 1 : L.nx;
    if (p_hi < p_lo) return FS_OK;
    for (int c = 0; c < 3; ++c)                                                           // generation 1 = (sphi >= 0)  (:479-481)
        FS_CUDA(cudaMemcpyAsync(h->valid + c * L.NL + (size_t)L.wlo * L.sx, h->mask + c * L.NL + (size_t)L.wlo * L.sx,
                                (size_t)(L.whi - L.wlo + 1) * L.sx, cudaMemcpyDeviceToDevice, s));
    (void)NL3;
    // single GPU: sweep 1 is a full pass that records what it filled, later sweeps only look around those faces.
    // Multi-GPU slabs always run full passes (the neighbours' fills arrive through the halo exchange, not the list).
    ExtrapWork W = h->work;
    if (h->comm) W.cap = 0u;
    if (W.cap) FS_CUDA(cudaMemsetAsync(W.count, 0, 4 * sizeof(unsigned int), s));
    const long long g_begin = (long long)p_lo * L.sx / 4, g_end = (long long)(p_hi + 1) * L.sx / 4;
    const int full_grid = (int)((g_end - g_begin + kThreads - 1) / kThreads);
    for (int k = 1; k <= sweeps; ++k) {
        const int push_to = (W.cap && k < sweeps) ? ((k - 1) & 1) : -1;
        if (k == 1 && region && W.cap) {
            FS_DISPATCH(h, visc3d_extrapolate_region_kernel<T><<<kSMs * 16, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->valid, k, W, push_to, h->rowflag,
                                                                                             region->list, region->nseg_dev, g_begin, g_end));
            FS_LAUNCH_CHECK();
        } else if (k == 1 || !W.cap) {
            FS_DISPATCH(h, visc3d_extrapolate_kernel<T><<<full_grid, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->valid, k, W, push_to, 0,
                                                                                       k == 1 ? h->rowflag : nullptr, g_begin, g_end));
            FS_LAUNCH_CHECK();
        } else {
            const int from = (k - 2) & 1;
            if (push_to >= 0) FS_CUDA(cudaMemsetAsync(W.count + push_to, 0, sizeof(unsigned int), s));
            FS_DISPATCH(h, visc3d_extrapolate_list_kernel<T><<<kSMs * 16, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->valid, k, W, from, push_to));
            FS_LAUNCH_CHECK();
            FS_DISPATCH(h, visc3d_extrapolate_kernel<T><<<kSMs * 8, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, vec), h->valid, k, W, -1, 1, nullptr, g_begin, g_end));   // only if a list overflowed
            FS_LAUNCH_CHECK();
        }
        if (h->comm) {   // the sweep is Jacobi over the GLOBAL grid: refresh the halo planes of the new values and generations
            FS_TRY(visc3d_halo_vec(h, vec, s));
            FS_TRY(visc3d_halo(h, (char*)h->valid, 1, COMM_U8, s));
        }
    }
    return FS_OK;
}

int fs_visc3d_extrapolate(fs_visc3d* h, int vec, int sweeps, void* stream) {
    return visc3d_extrapolate_impl(h, vec, sweeps, stream, nullptr);
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
    h->sparse_clean = false;
    return visc3d_general(h, scale, mu, src_vec, dst_vec, ROW_RHS, (cudaStream_t)stream);
}

int fs_visc3d_apply(fs_visc3d* h, double scale, double mu, int src_vec, int dst_vec, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    h->sparse_clean = false;
    return visc3d_general(h, scale, mu, src_vec, dst_vec, ROW_APPLY, (cudaStream_t)stream);
}

static int visc3d_k1(fs_visc3d* h, double sm, cudaStream_t s) {
    const int cap = kSMs * (h->dtype == FS_F32 ? K1Occ<float>::value : K1Occ<double>::value);
    const int grid = seg_grid(h->seg.nseg, kK1SegsPerBlock, cap);
    if (h->peers) {
        FS_DISPATCH(h, visc3d_apply_dot_kernel<T, true><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->seg.list, h->seg.nseg_dev, h->st, h->partials, h->peers, h->hot));
    } else {
        FS_DISPATCH(h, visc3d_apply_dot_kernel<T, false><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->seg.list, h->seg.nseg_dev, h->st, h->partials, nullptr, h->hot));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}
static int visc3d_k2(fs_visc3d* h, cudaStream_t s, int freeze = 0) {
    FS_DISPATCH(h, FS_TRY((cg_launch_update_xr_seg<T, 3>(h->L.NL, h->L.NL, h->seg, vec_ptr<T>(h, FS_VEC_X), vec_ptr<T>(h, FS_VEC_R), vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), h->st, h->partials, s, freeze, h->peers, h->peers ? &h->hot : nullptr))));
    return FS_OK;
}
static int visc3d_k3(fs_visc3d* h, cudaStream_t s) {
    FS_DISPATCH(h, FS_TRY((cg_launch_update_d_seg<T, 3>(h->L.NL, h->L.NL, h->seg, vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_R), h->st, s))));
    return FS_OK;
}
// Shared-memory tiled form of the stand-alone K1s (visc3d_apply_dot2_tile_kernel) for dense lattices on a single GPU.
// "k1_tile" / FLUIDSOLVER_B200_K1TILE: 0 = off, 1 = on when at least 60 % of the lattice segments are active, 2 = on every
// lattice (tests).  Returns 1 when it does not apply (nothing enqueued).
static int visc3d_k1t(fs_visc3d* h, double sm, cudaStream_t s, int freeze) {
    const int opt = tuning(OPT_K1TILE);
    const int v = opt < 0 ? kK1TileDefault : opt;
    const int mode = v % 10, ty_req = (v / 10) % 100, xl_req = v / 1000;     // (experiments: 1000*planes + 10*rows + mode)
    if (mode == 0 || h->peers || h->comm || h->k1t_smem <= 0) return 1;
    if (mode != 2 && (double)h->seg.nseg < 0.6 * (double)h->seg.nseg_total) return 1;
    if (h->L.nx < 2) return 1;
    const int Zp = h->L.Zp;
    const long long halo = k1t_halo(Zp);
    // rows per block: one lattice point per thread, as many rows as fit the shared memory of an SM (the halo rows are loaded
    // per block, so more rows = less traffic)
    long long ty = 0;
    for (long long c = 1; c * Zp <= kK1tMaxThreads && c <= h->L.Y; ++c)
        if (k1t_smem_bytes(c * Zp, halo, h->esz) <= (size_t)h->k1t_smem) ty = c;
    if (ty_req > 0 && ty_req < ty) ty = ty_req;
    if (ty < 1) return 1;                          // rows too long: the list kernel handles it
    const long long tile = ty * Zp;
    const size_t smem = k1t_smem_bytes(tile, halo, h->esz);
    const int xl = xl_req > 0 ? xl_req : kK1tPlanes;
    const int nyb = (int)((h->L.Y + ty - 1) / ty), nxc = (h->L.nx - 1 + xl - 1) / xl;
    const long long nitems = (long long)nyb * nxc;
    const int grid = (int)(nitems < kSMs ? nitems : kSMs);
    const int threads = (int)((tile + 31) / 32 * 32);
    const int tyi = (int)ty;
    FS_DISPATCH(h, visc3d_apply_dot2_tile_kernel<T><<<grid, threads, smem, s>>>(dev_view<T>(h), vec_ptr<T>(h, FS_VEC_R), reinterpret_cast<T*>(h->d2), h->st, h->partials, tyi, xl, freeze));
    FS_LAUNCH_CHECK();
    (void)sm;
    return FS_OK;
}

// single-reduction CG: K1s (w = A r, r.r, w.r, alpha/beta on the device) and K2s (p, s, x, r in one pass)
static int visc3d_k1s(fs_visc3d* h, double sm, cudaStream_t s, int freeze = 0) {
    {
        const int st = visc3d_k1t(h, sm, s, freeze);           // dense lattice: shared-memory tiled form
        if (st != 1) return st;
    }
    // Warp -> segment mapping of the stand-alone apply: "k1_block" / FLUIDSOLVER_B200_K1BLOCK = consecutive trips per CTA
    // block (0 = interleaved).  Blocking only pays on HBM-sized lists (it trades L2 requests for L1 hits); lists that fit in
    // L2 run through the persistent kernels anyway.
    const bool big = (double)h->seg.nseg * kSegPts * (13.0 * h->esz) > 64e6;
    const int blk_opt = tuning(OPT_K1BLOCK);
    int blk = blk_opt < 0 ? kK1BlockDefault : blk_opt;
    if (blk >= 1000) blk -= 1000;                // 1000 + n: also on small lists (tests)
    else if (!big) blk = 0;
    const int cap = kSMs * (h->dtype == FS_F32 ? K1Occ<float>::value : K1Occ<double>::value);
    const int grid = seg_grid(h->seg.nseg, kK1SegsPerBlock, cap);
    if (h->peers) {
        FS_DISPATCH(h, visc3d_apply_dot2_kernel<T, true><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_R), reinterpret_cast<T*>(h->d2), h->seg.list, h->seg.nseg_dev, h->st, h->partials, h->peers, h->hot, freeze, blk));
    } else {
        FS_DISPATCH(h, visc3d_apply_dot2_kernel<T, false><<<grid, kK1Threads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_R), reinterpret_cast<T*>(h->d2), h->seg.list, h->seg.nseg_dev, h->st, h->partials, nullptr, h->hot, freeze, blk));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}
static int visc3d_k2s(fs_visc3d* h, cudaStream_t s, int freeze = 0) {
    const int grid = seg_grid(h->seg.nseg, kSegsPerVecBlock, kVecGrid);
    FS_DISPATCH(h, cg_update_sr_seg_kernel<T, 3><<<grid, kVecThreads, 0, s>>>(h->L.NL, h->L.NL, h->seg.list, h->seg.nseg_dev, vec_ptr<T>(h, FS_VEC_X), vec_ptr<T>(h, FS_VEC_R),
                                                                              vec_ptr<T>(h, FS_VEC_D), vec_ptr<T>(h, FS_VEC_Q), reinterpret_cast<const T*>(h->d2), h->st, freeze));
    FS_LAUNCH_CHECK();
    return FS_OK;
}

// Which CG recurrence runs: the single-reduction form (default) or the reference's two-reduction loop
// (FS_CG_KERNELS / FS_CG_PERSISTENT, and always with the NCCL transport, whose collectives are host-launched).
static bool visc3d_use_sr(const fs_visc3d* h) {
    if (h->comm && !h->peers) return false;
    if (h->cg_mode == FS_CG_KERNELS || h->cg_mode == FS_CG_PERSISTENT) return false;
    if (h->cg_mode == FS_CG_KERNELS_SR || h->cg_mode == FS_CG_PERSISTENT_SR) return true;
    // fp32 STORAGE (opt-in) keeps the two-reduction recurrence by default: s = w + beta*s accumulates the fp32 rounding of
    // A p over the iterations, and on the tiny stiff fixtures the velocities then miss the 1e-4 bar (1.02e-4 measured)
    if (h->dtype == FS_F32) return false;
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("FLUIDSOLVER_B200_SR");
        mode = !e ? -1 : (e[0] == '0' ? 0 : 1);
    }
    return mode != 0;
}

static int visc3d_iteration(fs_visc3d* h, double sm, cudaStream_t s) {
    if (visc3d_use_sr(h)) {
        // one reduction per iteration; multi-GPU (fused transport): K1s pushes its boundary w rows into the neighbours and
        // all-reduces both scalars in its tail, K2s keeps the mirrored halo rows of p, s, x, r current.  No other traffic.
        FS_TRY(visc3d_k1s(h, sm, s));
        FS_TRY(visc3d_k2s(h, s));
        return FS_OK;
    }
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
// Persistent whole-iteration kernel: used when the per-iteration work is small enough that launch gaps and kernel
// ramp-up/tails matter (FLUIDSOLVER_B200_PERSISTENT=0/1 forces it off/on; default: CG working set <= 256 MB).
// Not with the NCCL transport (host-launched collectives between the phases), and not on a context that cannot hold
// the cooperative grid (MPS / green contexts: the launch failed once -> the stand-alone kernels take over).
static bool visc3d_use_persistent(const fs_visc3d* h) {
    if (h->comm && !h->peers) return false;
    if (h->coop_failed) return false;
    if (h->cg_mode == FS_CG_KERNELS || h->cg_mode == FS_CG_KERNELS_SR) return false;
    if (h->cg_mode == FS_CG_PERSISTENT || h->cg_mode == FS_CG_PERSISTENT_SR) return true;
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("FLUIDSOLVER_B200_PERSISTENT");
        mode = !e ? -1 : (e[0] == '0' ? 0 : 1);
    }
    if (mode >= 0) return mode == 1;
    const double ws = (double)h->seg.nseg * kSegPts * (22.0 * h->esz + 1.0);
    return ws <= 256e6;
}

// `n` = iteration slots (classic: iterations; single-reduction: an extra closing slot evaluates r.r of the last iterate)
// Shared-memory resident form of the single-reduction persistent kernel (single GPU / gathered solve); "resident_form" /
// FLUIDSOLVER_B200_RESIDENT = 0 selects the global-memory kernel.  Returns 1 if it cannot run here (nothing enqueued), so
// that the caller falls back.
static int resident_form() {
    const int v = tuning(OPT_RESIDENT_FORM);
    return v < 0 ? kResidentFormDefault : v;
}

static int visc3d_persistent_resident2(fs_visc3d* h, long long n, cudaStream_t s) {
    int dev = 0, smem_max = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 1;
    const void* fn = nullptr;
    size_t slot = 0;
    FS_DISPATCH(h, { fn = (const void*)visc3d_cg_sr_resident2_kernel<T>; slot = res2_slot_bytes<T>(); });
    const int kWarps = kRes2Threads / 32;
    const int max_slots = (int)(((size_t)smem_max - 1024) / slot);                  // 1 KB left for the static reduction scratch
    if (max_slots < 1) return 1;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(max_slots * slot)) != cudaSuccess) { cudaGetLastError(); return 1; }
    const int cap = coop_max_blocks(fn, kRes2Threads, max_slots * slot);
    if (cap < 1) return 1;
    int grid = seg_grid(h->seg.nseg, kWarps, cap < kSMs ? cap : kSMs);
    if (const char* e = getenv("FLUIDSOLVER_B200_PERSIST_GRID")) { const int g = atoi(e); if (g >= 1 && g <= cap) grid = g; }
    const long long chunk = ((long long)h->seg.nseg + grid - 1) / grid;             // run length per CTA
    int res_slots = (int)(chunk < max_slots ? chunk : max_slots);
    if (res_slots < 1) res_slots = 1;
    const size_t smem = (size_t)res_slots * slot;
    unsigned long long* prof = getenv("FLUIDSOLVER_B200_PROFILE") ? reinterpret_cast<unsigned long long*>(h->valid) : nullptr;
    while (n > 0) {
        int ni = (int)(n < (1 << 20) ? n : (1 << 20));
        cudaError_t e = cudaMemsetAsync(h->bar, 0, sizeof(GridBar), s);
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        FS_DISPATCH(h, {
            Visc3Dev<T> P = dev_view<T>(h);
            T* x = vec_ptr<T>(h, FS_VEC_X); T* r = vec_ptr<T>(h, FS_VEC_R); T* d = vec_ptr<T>(h, FS_VEC_D); T* q = vec_ptr<T>(h, FS_VEC_Q);
            T* w = reinterpret_cast<T*>(h->d2);
            const int* seg = h->seg.list; const int* nsegp = h->seg.nseg_dev;
            CgState* st = h->st; double* partials = h->partials; GridBar* bar = h->bar;
            void* args[] = {&P, &x, &r, &d, &q, &w, &seg, &nsegp, &st, &partials, &bar, &ni, &res_slots, &prof};
            e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kRes2Threads), args, smem, s);
        });
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported || e == cudaErrorInvalidValue) {
            cudaGetLastError();
            return 1;
        }
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
        FS_LAUNCH_CHECK();
        n -= ni;
    }
    return FS_OK;
}

static int visc3d_persistent(fs_visc3d* h, double sm, long long n, cudaStream_t s) {
    const bool sr = visc3d_use_sr(h);
    if (sr && !h->peers && !h->resident_failed) {
        const int form = resident_form();
        const int st = form != 0 ? visc3d_persistent_resident2(h, n, s) : 1;
        if (st != 1) return st;
        if (form != 0) h->resident_failed = true;   // not possible on this context: the global-memory form below takes over
    }
    const void* fn = nullptr;
    FS_DISPATCH(h, fn = sr ? (h->peers ? (const void*)visc3d_cg_sr_persistent_kernel<T, true> : (const void*)visc3d_cg_sr_persistent_kernel<T, false>)
                           : (h->peers ? (const void*)visc3d_cg_persistent_kernel<T, true> : (const void*)visc3d_cg_persistent_kernel<T, false>));
    const int cap = coop_max_blocks(fn, kPersistThreads);        // SMs of this context x resident CTAs per SM
    if (cap < 1) { h->coop_failed = true; return 1; }
    int grid = seg_grid(h->seg.nseg, kPersistThreads / 32, cap < kSMs ? cap : kSMs);
    if (const char* e = getenv("FLUIDSOLVER_B200_PERSIST_GRID")) { const int g = atoi(e); if (g >= 1 && g <= cap) grid = g; }
    unsigned long long* prof = getenv("FLUIDSOLVER_B200_PROFILE") ? reinterpret_cast<unsigned long long*>(h->valid) : nullptr;   // scratch between solves
    while (n > 0) {
        int ni = (int)(n < (1 << 20) ? n : (1 << 20));
        cudaError_t e = cudaMemsetAsync(h->bar, 0, sizeof(GridBar), s);     // arrival counter / flag count from zero in every launch
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        FS_DISPATCH(h, {
            Visc3Dev<T> P = dev_view<T>(h);
            T sv = (T)sm, s2v = (T)(2 * sm);
            T* x = vec_ptr<T>(h, FS_VEC_X); T* r = vec_ptr<T>(h, FS_VEC_R); T* d = vec_ptr<T>(h, FS_VEC_D); T* q = vec_ptr<T>(h, FS_VEC_Q);
            T* w = reinterpret_cast<T*>(h->d2);
            const int* seg = h->seg.list; const int* nsegp = h->seg.nseg_dev;
            CgState* st = h->st; double* partials = h->partials; GridBar* bar = h->bar;
            PeerInfo* peers = h->peers; PeerHot hot = h->hot;
            void* args_sr[] = {&P, &sv, &s2v, &x, &r, &d, &q, &w, &seg, &nsegp, &st, &partials, &bar, &ni, &peers, &hot, &prof};
            void* args_cl[] = {&P, &sv, &s2v, &x, &r, &d, &q, &seg, &nsegp, &st, &partials, &bar, &ni, &peers, &hot, &prof};
            e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPersistThreads), sr ? args_sr : args_cl, 0, s);
        });
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
            cudaGetLastError();                      // this context cannot run the cooperative grid: use the stand-alone kernels
            h->coop_failed = true;
            return 1;
        }
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
        FS_LAUNCH_CHECK();
        n -= ni;
    }
    return FS_OK;
}

// iterations per enqueue of cg_drive: stand-alone kernels go out in small batches (the host checks the convergence flag in
// between); a persistent kernel stops by itself when the state says done, so its launches are long — and longer still for
// the resident kernels, whose every launch first loads and finally stores the shared-memory resident state
static int visc3d_batch(const fs_visc3d* h) {
    if (!visc3d_use_persistent(h)) return kCgBatch;
    const int form = tuning(OPT_RESIDENT_FORM);
    const bool resident = visc3d_use_sr(h) && !h->peers && !h->resident_failed && (form < 0 ? kResidentFormDefault : form) != 0;
    return resident ? kCgBatchResident : kCgBatchPersistent;
}

static int visc3d_iterations(fs_visc3d* h, double sm, long long n, cudaStream_t s) {
    if (visc3d_use_persistent(h)) {
        const int st = visc3d_persistent(h, sm, n, s);
        if (st != 1) return st;                      // 1 = cooperative launch impossible here, nothing was enqueued
    }
    const bool graph_ok = !(h->comm && !h->peers);
    return cg_enqueue_iterations(h->graph, graph_ok, sm, n, [&](cudaStream_t ss) { return visc3d_iteration(h, sm, ss); }, s,
                                 seg_level(h->seg.nseg) * 4096 + (tuning_epoch() & 4095));
}

// pre-scaled coefficients of the CG-loop apply for this solve's operator (scale*mu) on the current active list
static int visc3d_scale(fs_visc3d* h, double sm, cudaStream_t s) {
    const int grid = seg_grid(h->seg.nseg, kThreads / 32, kSMs * 4);
    FS_DISPATCH(h, visc3d_scale_kernel<T><<<grid, kThreads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), reinterpret_cast<T*>(h->coefs), h->seg.list, h->seg.nseg_dev));
    FS_LAUNCH_CHECK();
    return FS_OK;
}

// Start of a solve on the active set (fs_visc3d_solve): one kernel builds b, q = A x, d = r = b - q and delta0.
static int visc3d_cg_begin_sparse(fs_visc3d* h, double scale, double mu, double tol, int64_t max_iter, cudaStream_t s) {
    FS_TRY(h->seg.finish());                      // list length: enqueued by pack, needed from here on for the launch sizes
    const double sm = scale * mu;
    FS_TRY(visc3d_scale(h, sm, s));
    cg_state_init_kernel<<<1, 1, 0, s>>>(h->st, tol * tol, (long long)max_iter, (h->comm && !h->peers) ? 1 : 0);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaMemsetAsync(h->bar, 0, sizeof(GridBar), s));
    if (h->peers) {   // belt and braces: the halo planes of q (and w) must read as zero until the neighbours store into them
        for (int c = 0; c < 3; ++c) {
            char* comps[3] = {h->vecs + ((size_t)FS_VEC_Q * 3 + c) * h->L.NL * h->esz, h->d2 + (size_t)c * h->L.NL * h->esz,
                              h->d2 + (size_t)(3 + c) * h->L.NL * h->esz};
            for (char* comp : comps) {
                if (h->has_lo) FS_CUDA(cudaMemsetAsync(comp, 0, (size_t)h->L.sx * h->esz, s));
                if (h->has_hi) FS_CUDA(cudaMemsetAsync(comp + (size_t)(h->L.X - 2) * h->L.sx * h->esz, 0, (size_t)h->L.sx * h->esz, s));
            }
        }
    }
    const int grid = seg_grid(h->seg.nseg, kThreads / 32, kSMs * 4);
    FS_DISPATCH(h, visc3d_begin_kernel<T><<<grid, kThreads, 0, s>>>(dev_view<T>(h), (T)sm, (T)(2 * sm), vec_ptr<T>(h, FS_VEC_X), vec_ptr<T>(h, FS_VEC_B),
                                                                  vec_ptr<T>(h, FS_VEC_Q), vec_ptr<T>(h, FS_VEC_R), vec_ptr<T>(h, FS_VEC_D),
                                                                  h->seg.list, h->seg.nseg_dev, h->st, h->partials, h->peers));
    FS_LAUNCH_CHECK();
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

static int visc3d_cg_begin(fs_visc3d* h, double scale, double mu, double tol, int64_t max_iter, cudaStream_t s) {
    FS_TRY(h->seg.finish());
    const long long n = 3 * h->L.NL;
    FS_TRY(visc3d_scale(h, scale * mu, s));
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
    h->sparse_clean = false;                      // dense begin: q, r, d are rewritten everywhere
    FS_CUDA(cudaMemsetAsync(h->bar, 0, sizeof(GridBar), s));
    FS_TRY(visc3d_cg_begin(h, scale, mu, tol, max_iter, s));
    const double sm = scale * mu;
    // single-reduction CG: one extra slot whose apply evaluates r.r of the last iterate (and sets the final status)
    return cg_drive(h->cg, [&](cudaStream_t ss, long long nb) { return visc3d_iterations(h, sm, nb, ss); }, (long long)max_iter + (visc3d_use_sr(h) ? 1 : 0), stats, s,
                    visc3d_batch(h));
}

int fs_visc3d_cg_enqueue(fs_visc3d* h, double scale, double mu, int64_t n, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_cg_enqueue before fs_visc3d_pack");
    FS_TRY(h->seg.finish());
    const double sm = scale * mu;
    cg_state_unlimit_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->st);
    FS_LAUNCH_CHECK();
    return visc3d_iterations(h, sm, n, (cudaStream_t)stream);
}

int fs_visc3d_kernel_enqueue(fs_visc3d* h, int which, double scale, double mu, int64_t n, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_kernel_enqueue before fs_visc3d_pack");
    FS_TRY(h->seg.finish());
    if (which < 1 || which > 3) return fail(FS_ERR_ARG, "fs_visc3d_kernel_enqueue: which must be 1, 2 or 3");
    cudaStream_t s = (cudaStream_t)stream;
    const double sm = scale * mu;
    cg_state_unlimit_kernel<<<1, 1, 0, s>>>(h->st);
    FS_LAUNCH_CHECK();
    const bool sr = visc3d_use_sr(h);
    if (sr && which == 3) return fail(FS_ERR_ARG, "fs_visc3d_kernel_enqueue: the single-reduction CG has two kernels (1, 2)");
    for (int64_t k = 0; k < n; ++k) {
        if (which == 1) FS_TRY(sr ? visc3d_k1s(h, sm, s, 1) : visc3d_k1(h, sm, s));
        else if (which == 2) FS_TRY(sr ? visc3d_k2s(h, s, 1) : visc3d_k2(h, s, 1));
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

int fs_visc3d_set_window(fs_visc3d* h, int cell_lo, int cell_hi) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (h->comm) return fail(FS_ERR_STATE, "fs_visc3d_set_window: not on a slab handle (fs_visc3d_set_slab)");
    if (cell_lo < 0 || cell_hi > h->L.nx || cell_hi - cell_lo < 1) return fail(FS_ERR_ARG, "fs_visc3d_set_window: bad cell range");
    h->L.wlo = cell_lo; h->L.whi = cell_hi; h->L.wcells = cell_hi;
    h->windowed = !(cell_lo == 0 && cell_hi == h->L.nx);
    h->packed = false;
    return FS_OK;
}

// :569-573 — the caller's velocities into the lattice vector x and the three extrapolation sweeps: around the active set
// only (see visc3d_mark_region_kernel; "sparse_setup" / FLUIDSOLVER_B200_SPARSE_SETUP = 0 switches back), else on the whole
// lattice / x-window.  Needs the pack of this solve (its activity map and segment list) enqueued on the same stream.  The
// region list lives in `xseg` until the gathered solve reuses that list for its publish flags (stream order keeps the two apart).
static int visc3d_load_extrapolate(fs_visc3d* h, const void* vx, const void* vy, const void* vz, int vel_dtype, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int sp = tuning(OPT_SPARSE_SETUP);
    if ((sp < 0 ? kSparseSetupDefault : sp) != 0 && !h->comm && h->work.cap != 0u) {
        FS_CUDA(cudaMemsetAsync(h->xflags, 0, (size_t)h->xseg.nseg_total, s));
        visc3d_mark_region_kernel<<<kSMs * 8, 256, 0, s>>>(h->L, h->seg.list, h->seg.nseg_dev, h->xseg.nseg_total, h->xflags);
        FS_LAUNCH_CHECK();
        FS_TRY(h->xseg.enqueue(h->xflags, s, 1));
        if (vel_dtype == FS_F32) {
            FS_DISPATCH(h, visc3d_load_region_kernel<T, float><<<kSMs * 8, kThreads, 0, s>>>(h->L, (const float*)vx, (const float*)vy, (const float*)vz, vec_ptr<T>(h, FS_VEC_X), h->xseg.list, h->xseg.nseg_dev));
        } else {
            FS_DISPATCH(h, visc3d_load_region_kernel<T, double><<<kSMs * 8, kThreads, 0, s>>>(h->L, (const double*)vx, (const double*)vy, (const double*)vz, vec_ptr<T>(h, FS_VEC_X), h->xseg.list, h->xseg.nseg_dev));
        }
        FS_LAUNCH_CHECK();
        return visc3d_extrapolate_impl(h, FS_VEC_X, 3, stream, &h->xseg);
    }
    FS_TRY(fs_visc3d_load(h, FS_VEC_X, vx, vy, vz, vel_dtype, stream));
    return fs_visc3d_extrapolate(h, FS_VEC_X, 3, stream);
}

size_t fs_visc3d_gather_record_bytes(const fs_visc3d* h) { return h ? gather_record_bytes(h->esz) : 0; }
int fs_visc3d_gather_reexport(fs_visc3d* h, void* records, int64_t cap, void* stream);

int fs_visc3d_gather_export(fs_visc3d* h, const void* vx, const void* vy, const void* vz, int vel_dtype, const double* sphi, const double* lvol,
                            double vol_norm, int own_lo, int own_hi, void* records, int64_t cap, int64_t* count, void* stream) {
    if (!h || !vx || !vy || !vz || !sphi || !lvol || !records || !count) return fail(FS_ERR_ARG, "fs_visc3d_gather_export: null argument");
    if (h->comm) return fail(FS_ERR_STATE, "fs_visc3d_gather_export: not on a slab handle");
    const Lat3& L = h->L;
    if (own_lo < L.wlo || own_hi > L.whi + 1 || own_hi <= own_lo) return fail(FS_ERR_ARG, "fs_visc3d_gather_export: owned planes outside the window");
    // clean results need (sweeps + 1) = 4 window planes beyond the owned ones wherever the window does not end at the grid boundary
    if ((L.wlo > 0 && own_lo - L.wlo < 4) || (L.whi < L.nx && L.whi - own_hi < 4))
        return fail(FS_ERR_ARG, "fs_visc3d_gather_export: the window must extend 4 cells beyond the owned planes");
    cudaStream_t s = (cudaStream_t)stream;
    FS_TRY(fs_visc3d_pack(h, sphi, lvol, vol_norm, stream));                    // window planes only (+ wipes the previous solve's segments)
    FS_TRY(visc3d_load_extrapolate(h, vx, vy, vz, vel_dtype, stream));          // (around the window's active rows only, unless switched off)
    FS_TRY(h->seg.finish());                                                    // (the list pack enqueued is not used: the global one follows the import)
    const long long nseg_total = h->xseg.nseg_total;
    FS_CUDA(cudaMemsetAsync(h->xflags, 0, (size_t)nseg_total, s));
    const long long seg_lo = (long long)own_lo * L.sx / kSegPts;
    long long seg_hi = ((long long)own_hi * L.sx + kSegPts - 1) / kSegPts;
    if (seg_hi > nseg_total) seg_hi = nseg_total;
    visc3d_mark_export_kernel<<<(unsigned)((seg_hi - seg_lo + 255) / 256), 256, 0, s>>>(L, h->act, seg_lo, seg_hi, nseg_total, h->xflags);
    FS_LAUNCH_CHECK();
    FS_TRY(h->xseg.enqueue(h->xflags, s, 1));
    FS_TRY(h->xseg.finish());
    *count = h->xseg.nseg;
    if (h->xseg.nseg > cap) return FS_OK;       // nothing written: the caller grows its buffer and calls fs_visc3d_gather_reexport
    return fs_visc3d_gather_reexport(h, records, cap, stream);
}

int fs_visc3d_gather_reexport(fs_visc3d* h, void* records, int64_t cap, void* stream) {
    if (!h || !records) return fail(FS_ERR_ARG, "fs_visc3d_gather_reexport: null argument");
    if (h->xseg.nseg > cap) return fail(FS_ERR_ARG, "fs_visc3d_gather_reexport: record buffer too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->xseg.nseg > 0) {
        const int grid = seg_grid(h->xseg.nseg, kThreads / 32, kSMs * 8);
        FS_DISPATCH(h, visc3d_export_kernel<T><<<grid, kThreads, 0, s>>>(h->L.NL, reinterpret_cast<const T*>(h->coef), vec_ptr<T>(h, FS_VEC_X), h->mask, h->act,
                                                                         h->xseg.list, h->xseg.nseg_dev, (long long)cap, (char*)records));
        FS_LAUNCH_CHECK();
    }
    return FS_OK;
}

int fs_visc3d_gather_import(fs_visc3d* h, const void* records, const int64_t* counts_dev, int nranks, int64_t stride, int skip_rank, void* stream) {
    if (!h || !records || !counts_dev) return fail(FS_ERR_ARG, "fs_visc3d_gather_import: null argument");
    if (nranks < 1 || stride < 0) return fail(FS_ERR_ARG, "fs_visc3d_gather_import: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    if (stride > 0) {
        const long long total = (long long)nranks * stride;
        long long blocks = (total + kThreads / 32 - 1) / (kThreads / 32);
        if (blocks > kSMs * 8) blocks = kSMs * 8;
        FS_DISPATCH(h, visc3d_import_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(h->L.NL, reinterpret_cast<T*>(h->coef), vec_ptr<T>(h, FS_VEC_X), h->mask, h->act,
                                                                                    (const char*)records, (const long long*)counts_dev, nranks, (long long)stride, skip_rank));
        FS_LAUNCH_CHECK();
    }
    FS_TRY(h->seg.enqueue(h->act, s));          // the GLOBAL active list: identical on every rank
    h->active_rows = -1;
    h->packed = true;
    return FS_OK;
}

int fs_visc3d_solve_packed(fs_visc3d* h, double dt, double mu, double rho, double cell_vol, void* vx, void* vy, void* vz, int vel_dtype,
                           int own_lo, int own_hi, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h || !vx || !vy || !vz) return fail(FS_ERR_ARG, "fs_visc3d_solve_packed: null argument");
    if (!h->packed) return fail(FS_ERR_STATE, "fs_visc3d_solve_packed before pack / gather_import");
    if (max_iter < 0) return fail(FS_ERR_ARG, "fs_visc3d_solve_packed: max_iter < 0");
    if (vel_dtype != FS_F32 && vel_dtype != FS_F64) return fail(FS_ERR_ARG, "fs_visc3d_solve_packed: bad velocity dtype");
    cudaStream_t s = (cudaStream_t)stream;
    const double scale = dt / cell_vol / rho;
    const double sm = scale * mu;
    FS_TRY(visc3d_cg_begin_sparse(h, scale, mu, tol, max_iter, s));
    int status = cg_drive(h->cg, [&](cudaStream_t ss, long long nb) { return visc3d_iterations(h, sm, nb, ss); }, (long long)max_iter + (visc3d_use_sr(h) ? 1 : 0), stats, s,
                          visc3d_batch(h));
    if (status != FS_OK) return status;
    const int grid = seg_grid(h->seg.nseg, kThreads / 32, kSMs * 4);
    if (vel_dtype == FS_F32) {
        FS_DISPATCH(h, visc3d_store_active_kernel<T, float><<<grid, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, FS_VEC_X), h->act, h->seg.list, h->seg.nseg_dev, (float*)vx, (float*)vy, (float*)vz, own_lo, own_hi));
    } else {
        FS_DISPATCH(h, visc3d_store_active_kernel<T, double><<<grid, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, FS_VEC_X), h->act, h->seg.list, h->seg.nseg_dev, (double*)vx, (double*)vy, (double*)vz, own_lo, own_hi));
    }
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaStreamSynchronize(s));
    return FS_OK;
}

int fs_visc3d_solve(fs_visc3d* h, double dt, double mu, double rho, double cell_vol,
                    void* vx, void* vy, void* vz, int vel_dtype, const double* sphi, const double* lvol,
                    double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (max_iter < 0) return fail(FS_ERR_ARG, "fs_visc3d_solve: max_iter < 0");
    if (vel_dtype != FS_F32 && vel_dtype != FS_F64) return fail(FS_ERR_ARG, "fs_visc3d_solve: bad velocity dtype");
    cudaStream_t s = (cudaStream_t)stream;
    const double scale = dt / cell_vol / rho;                                   // :567
    const double sm = scale * mu;
    FS_TRY(fs_visc3d_pack(h, sphi, lvol, cell_vol * 0.125, stream));            // :568 (+ active segment list)
    FS_TRY(visc3d_load_extrapolate(h, vx, vy, vz, vel_dtype, stream));          // :569-573 (where the solve looks)
    FS_TRY(visc3d_cg_begin_sparse(h, scale, mu, tol, max_iter, s));             // :574-587 on the active set
    // :588-612 (single-reduction CG: one extra slot whose apply evaluates r.r of the last iterate and sets the final status)
    int status = cg_drive(h->cg, [&](cudaStream_t ss, long long nb) { return visc3d_iterations(h, sm, nb, ss); }, (long long)max_iter + (visc3d_use_sr(h) ? 1 : 0), stats, s,
                          visc3d_batch(h));
    if (status != FS_OK) return status;                                         // the reference raises before write-back
    {
        // :613 — only rows of the active set can differ from what was loaded from these very arrays; every other fluid face
        // keeps the caller's value untouched (also when the solver stores fp32 and the caller fp64)
        const int grid = seg_grid(h->seg.nseg, kThreads / 32, kSMs * 4);
        if (vel_dtype == FS_F32) {
            FS_DISPATCH(h, visc3d_store_active_kernel<T, float><<<grid, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, FS_VEC_X), h->act, h->seg.list, h->seg.nseg_dev, (float*)vx, (float*)vy, (float*)vz, 0, h->L.X));
        } else {
            FS_DISPATCH(h, visc3d_store_active_kernel<T, double><<<grid, kThreads, 0, s>>>(h->L, vec_ptr<T>(h, FS_VEC_X), h->act, h->seg.list, h->seg.nseg_dev, (double*)vx, (double*)vy, (double*)vz, 0, h->L.X));
        }
        FS_LAUNCH_CHECK();
    }
    FS_CUDA(cudaStreamSynchronize(s));
    FS_TRY(h->xseg.finish());                                                   // (region list of a sparse set-up: its count has arrived long ago)
    return FS_OK;
}

}  // extern "C"
