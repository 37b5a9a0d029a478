// Pressure projection (ghost-fluid variable-coefficient Poisson CG) and solid face fractions, sm_100a.
//
// Replaces PressureCGSolver3D.py (:6-226), PressureCGSolver2D.py (:6-179), SolidFraction3D.py (:6-32),
// SolidFraction2D.py (:6-26), SolidFractionCommon.py (:4-60).
//
// The pressure path works directly on the reference's dense C-order fp64 arrays (cells (nx,ny[,nz]),
// face weights with +1 extent on their own axis, CGSolverBuffer's d,r,q,b): they are already SoA with
// the fastest axis contiguous, so no repacking is needed — one thread per cell with the fastest axis on
// threadIdx.x gives coalesced 8-byte accesses.  All arithmetic follows the reference's association with
// explicit round-to-nearest intrinsics (no FMA contraction), so fp64 results are bit-identical to it.
#include <stdlib.h>
#include <type_traits>

#include "fs_common.cuh"

namespace fs {

template <int D> struct Grid {
    int n[D];          // cells per axis
    long long cs[D];   // cell strides
    long long ncells;
};

template <int D> __host__ __device__ inline Grid<D> make_grid(int nx, int ny, int nz) {
    Grid<D> g;
    g.n[0] = nx; g.n[1] = ny;
    if (D == 3) g.n[2] = nz;
    long long s = 1;
    for (int a = D - 1; a >= 0; --a) { g.cs[a] = s; s *= g.n[a]; }
    g.ncells = s;
    return g;
}

template <int D> __device__ __forceinline__ void decode(const Grid<D>& g, long long i, int* c) {
#pragma unroll
    for (int a = D - 1; a >= 0; --a) { c[a] = (int)(i % g.n[a]); i /= g.n[a]; }
}

// index into the face array of axis `a` (shape n + e_a) at cell coordinates c (+ off along a)
template <int D> __device__ __forceinline__ long long face_idx(const Grid<D>& g, int a, const int* c, int off) {
    long long idx = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int ext = g.n[k] + (k == a ? 1 : 0);
        idx = idx * ext + c[k] + (k == a ? off : 0);
    }
    return idx;
}

// index into the (2n+1)^D fine grid at node 2*c + o
template <int D> __device__ __forceinline__ long long fine_idx(const Grid<D>& g, const int* c, const int* o) {
    long long idx = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) idx = idx * (2 * g.n[k] + 1) + 2 * c[k] + o[k];
    return idx;
}

template <int D> __device__ __forceinline__ bool interior(const Grid<D>& g, const int* c) {
    bool in = true;
#pragma unroll
    for (int k = 0; k < D; ++k) in = in && c[k] >= 1 && c[k] <= g.n[k] - 2;
    return in;
}

constexpr int kPT = 256;

template <int D> struct PressW { const double* w[D]; };
template <int D, typename S> struct Vel { S* v[D]; };

// edge_in_fraction (SolidFractionCommon.py:4-16)
__device__ __forceinline__ double edge_in_fraction(double l, double r) {
    const bool l_in = l < 0, r_in = r < 0;
    if (l_in && r_in) return 1.0;
    if (!l_in && !r_in) return 0.0;
    const double diff = -fabs(__dsub_rn(l, r));
    return l_in ? l / diff : r / diff;
}

// ---------------------------------------------------------------------------------------------
// matvecmul_kernel (PressureCGSolver3D.py:52-130 / 2D :46-100), optionally fused with d.q
// ---------------------------------------------------------------------------------------------
// Persistent grid (kPressBlocksPerSM CTAs per SM), grid-stride over the cells with the fastest axis across the warp, one
// deterministic block reduction at the very end (a reduction per 256 cells made the first version barrier-bound).
constexpr int kPressBlocksPerSM = 6;

// OP selects the operator: OP_PRESSURE = PressureCGSolver*.matvecmul_kernel (diagonal weighted by the face fractions);
// OP_DENSITY = DensityCGSolver3D.matvecmul_kernel (:117-194): unit diagonal contributions, and the -z off-diagonal term
// reads wz[x,y,z+1] (not wz[x,y,z]) exactly as the reference does (:184).
enum { OP_PRESSURE = 0, OP_DENSITY = 1 };

template <int D, bool CG, int OP>
__global__ void __launch_bounds__(kPT) press_apply_kernel(Grid<D> g, const double* __restrict__ v, double* __restrict__ out, PressW<D> W,
                                                          const double* __restrict__ lphi, CgState* st, double* partials) {
    if (CG) { if (*(volatile int*)&st->done) return; }
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.ncells; i += stride) {
        int c[D];
        decode<D>(g, i, c);
        if (!interior<D>(g, c)) continue;
        const double phi = __ldg(lphi + i);
        double res = 0.0;
        if (phi < 0) {
            double nphi[2 * D], w[2 * D], vn[2 * D];          // requested together, used in the reference's order (+a then -a)
            const double vc = __ldg(v + i);
#pragma unroll
            for (int a = 0; a < D; ++a) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const long long j = i + (t == 0 ? g.cs[a] : -g.cs[a]);
                    const int foff = (t == 0 || (OP == OP_DENSITY && a == D - 1)) ? 1 : 0;
                    nphi[2 * a + t] = __ldg(lphi + j);
                    w[2 * a + t] = __ldg(W.w[a] + face_idx<D>(g, a, c, foff));
                    vn[2 * a + t] = __ldg(v + j);
                }
            }
            double val = 0.0, diag = 0.0;
#pragma unroll
            for (int k = 0; k < 2 * D; ++k) {
                const double dw = (OP == OP_DENSITY) ? 1.0 : w[k];
                if (nphi[k] < 0) {
                    val = __dsub_rn(val, __dmul_rn(w[k], vn[k]));
                    diag = __dadd_rn(diag, dw);
                } else {
                    const double frac = fmin(1.0, fmax(0.01, phi / __dsub_rn(phi, nphi[k])));
                    diag = __dadd_rn(diag, dw / frac);
                }
            }
            res = __dadd_rn(val, __dmul_rn(diag, vc));
            if (CG) acc += vc * res;
        }
        out[i] = res;
    }
    if (CG) grid_sum_finish(acc, partials, &st->counter[0], [=](double s) { st->dq = s; });
}

// ---------------------------------------------------------------------------------------------
// Active-set CG for the pressure system.  Rows that the operator computes are the interior cells with lphi < 0; on
// every other cell b = q = r = d = 0 and x stays 0 for the whole solve.  One byte per cell marks the computed rows, the
// sorted list of active 32-cell segments is built once per solve (fs_common: SegList) and a persistent cooperative
// kernel runs whole iterations on it: apply + d.q, x/r update + r.r, d update as phases separated by grid barriers,
// the CG scalars replicated in registers (same structure as visc3d_cg_persistent_kernel).  Arithmetic and association
// are those of press_apply_kernel / cg_update_*; only the set of visited cells and the reduction order differ.
// ---------------------------------------------------------------------------------------------
template <int D> __device__ __forceinline__ void decode_fast(const Grid<D>& g, long long i, int* c) {
    if (g.ncells < 0x7fffffffLL) {
        unsigned int u = (unsigned int)i;
#pragma unroll
        for (int a = D - 1; a >= 0; --a) {
            const unsigned int n = (unsigned int)g.n[a];
            const unsigned int t = u / n;
            c[a] = (int)(u - t * n);
            u = t;
        }
        return;
    }
    decode<D>(g, i, c);
}

template <int D>
__global__ void __launch_bounds__(kPT) press_activity_kernel(Grid<D> g, const double* __restrict__ lphi, uint8_t* __restrict__ act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ncells) return;
    int c[D];
    decode_fast<D>(g, i, c);
    act[i] = (interior<D>(g, c) && lphi[i] < 0) ? 1 : 0;
}

// one pass over the active segments: out = A v on computed rows (a warp per 32-cell segment), returns this thread's share of v.out
template <int D, int OP>
__device__ __forceinline__ double press_apply_seg_body(const Grid<D>& g, const double* v, double* out, const PressW<D>& W,
                                                       const double* __restrict__ lphi, const int* __restrict__ seg, int nseg) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    int sg_n = w0 < nseg ? __ldg(seg + w0) : 0;
    for (long long k = w0; k < nseg; k += nw) {
        const long long i = (long long)sg_n * kSegPts + lane;
        {
            const long long k2 = k + nw;
            sg_n = k2 < nseg ? __ldg(seg + k2) : 0;
        }
        if (i >= g.ncells) continue;
        int c[D];
        decode_fast<D>(g, i, c);
        if (!interior<D>(g, c)) continue;
        const double phi = __ldg(lphi + i);
        if (!(phi < 0)) continue;                       // not computed: out holds 0 since the start of the solve
        // all 2D neighbour level-set values, face weights and vector entries are requested before the first use (one memory
        // latency per segment instead of one per neighbour); the arithmetic below keeps the reference's order (+a then -a)
        double nphi[2 * D], w[2 * D], vn[2 * D];
        const double vc = v[i];
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {                     // t = 0: +a, t = 1: -a
                const long long j = i + (t == 0 ? g.cs[a] : -g.cs[a]);
                const int foff = (t == 0 || (OP == OP_DENSITY && a == D - 1)) ? 1 : 0;
                nphi[2 * a + t] = __ldg(lphi + j);
                w[2 * a + t] = __ldg(W.w[a] + face_idx<D>(g, a, c, foff));
                vn[2 * a + t] = v[j];
            }
        }
        double val = 0.0, diag = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * D; ++k) {
            const double dw = (OP == OP_DENSITY) ? 1.0 : w[k];
            if (nphi[k] < 0) {
                val = __dsub_rn(val, __dmul_rn(w[k], vn[k]));
                diag = __dadd_rn(diag, dw);
            } else {
                const double frac = fmin(1.0, fmax(0.01, phi / __dsub_rn(phi, nphi[k])));
                diag = __dadd_rn(diag, dw / frac);
            }
        }
        const double res = __dadd_rn(val, __dmul_rn(diag, vc));
        acc += vc * res;
        out[i] = res;
    }
    return acc;
}

template <int D, int OP>
__global__ void __launch_bounds__(kPersistThreads, 1) press_cg_persistent_kernel(Grid<D> g, double* x, double* r, double* d, double* q, PressW<D> W,
                                                                                 const double* __restrict__ lphi, const int* __restrict__ seg,
                                                                                 const int* __restrict__ nseg_p, CgState* st, double* partials,
                                                                                 GridBar* bar, int n_iters) {
    const int nseg = *nseg_p;
    double delta = st->delta, delta_old = st->delta_old, dq = st->dq, alpha_d = st->alpha, beta_d = st->beta;
    const double tol2 = st->tol2;
    long long iter = st->iter;
    const long long max_iter = st->max_iter;
    int done = st->done;
    GridSync gs{bar, 0u};
    PeerHot hot;                                     // unused (single GPU)
    hot.has_lo = hot.has_hi = 0;
    for (int it = 0; it < n_iters && !done; ++it) {
        double acc = press_apply_seg_body<D, OP>(g, d, q, W, lphi, seg, nseg);
        dq = grid_allreduce(acc, partials, gs);
        alpha_d = delta / dq;
        acc = cg_update_xr_seg_body<double, 1, false>(g.ncells, g.ncells, seg, nseg, x, r, d, q, alpha_d, hot);
        const double rr = grid_allreduce(acc, partials, gs);
        delta_old = delta;
        delta = rr;
        iter += 1;
        if (rr < tol2) done = 1;
        else if (iter >= max_iter || !(rr == rr)) done = 2;
        if (done) break;
        beta_d = delta / delta_old;
        cg_update_d_seg_body<double, 1>(g.ncells, g.ncells, seg, nseg, d, r, beta_d);
        gs.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->delta = delta; st->delta_old = delta_old; st->dq = dq; st->alpha = alpha_d; st->beta = beta_d;
        st->iter = iter; st->done = done;
    }
}

// ---------------------------------------------------------------------------------------------
// initialize_solver_kernel (PressureCGSolver3D.py:6-50 / 2D :6-44)
// ---------------------------------------------------------------------------------------------
template <int D, typename S>
__global__ void __launch_bounds__(kPT) press_rhs_kernel(Grid<D> g, double cs0, double cs1, double cs2, Vel<D, const S> V, const double* __restrict__ sv,
                                                        const double* __restrict__ lphi, double* __restrict__ b, PressW<D> W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ncells) return;
    int c[D];
    decode<D>(g, i, c);
    if (!interior<D>(g, c)) return;
    if (!(lphi[i] < 0)) { b[i] = 0.0; return; }
    const double cs[3] = {cs0, cs1, cs2};
    double val = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        int o[D];
#pragma unroll
        for (int k = 0; k < D; ++k) o[k] = (k == a) ? 0 : 1;   // low face centre of this cell
        {   // + face
            const long long f = face_idx<D>(g, a, c, 1);
            const double w = W.w[a][f];
            val = __dadd_rn(val, __dmul_rn(w, (double)V.v[a][f]) / cs[a]);
            if (w < 1) {
                int o2[D];
#pragma unroll
                for (int k = 0; k < D; ++k) o2[k] = o[k] + (k == a ? 2 : 0);
                val = __dsub_rn(val, __dmul_rn(w, sv[fine_idx<D>(g, c, o2) * D + a]) / cs[a]);
            }
        }
        {   // - face
            const long long f = face_idx<D>(g, a, c, 0);
            const double w = W.w[a][f];
            val = __dsub_rn(val, __dmul_rn(w, (double)V.v[a][f]) / cs[a]);
            if (w < 1) val = __dadd_rn(val, __dmul_rn(w, sv[fine_idx<D>(g, c, o) * D + a]) / cs[a]);
        }
    }
    b[i] = val;
}

// ---------------------------------------------------------------------------------------------
// apply_pressure_kernel (PressureCGSolver3D.py:132-153 / 2D :102-120): indices 1..g-1 on every axis
// ---------------------------------------------------------------------------------------------
template <int D, typename S>
__global__ void __launch_bounds__(kPT) press_update_kernel(Grid<D> g, double cs0, double cs1, double cs2, Vel<D, S> V, const double* __restrict__ pv,
                                                           PressW<D> W, const double* __restrict__ sv, const double* __restrict__ lphi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ncells) return;
    int c[D];
    decode<D>(g, i, c);
#pragma unroll
    for (int k = 0; k < D; ++k) if (c[k] < 1) return;           // upper bound g-1 is the last cell
    const double cs[3] = {cs0, cs1, cs2};
    const double phi = lphi[i];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const long long j = i - g.cs[a];
        const double phim = lphi[j];
        if (phi < 0 || phim < 0) {
            const double theta = fmin(1.0, fmax(0.01, edge_in_fraction(phi, phim)));
            const long long f = face_idx<D>(g, a, c, 0);
            double nv = __dadd_rn((double)V.v[a][f], __dmul_rn(__dsub_rn(pv[i], pv[j]), cs[a]) / theta);
            const double w = W.w[a][f];
            int o[D];
#pragma unroll
            for (int k = 0; k < D; ++k) o[k] = (k == a) ? 0 : 1;
            const double s = sv[fine_idx<D>(g, c, o) * D + a];
            nv = __dadd_rn(__dmul_rn(w, nv), __dmul_rn(__dsub_rn(1.0, w), s));
            V.v[a][f] = (S)nv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// solid fractions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri_all_in(double a, double b, double c) { return (a < 0) && (b < 0) && (c < 0); }

// face_in_fraction (SolidFractionCommon.py:52-60).  tri_in_fraction's 1-in / 2-in branches evaluate
// edge_in_fraction on two same-sign vertices and therefore contribute 0; only 3-in triangles count.
__device__ __forceinline__ double face_in_fraction(double bl, double br, double tl, double tr) {
    const double ce = 0.25 * __dadd_rn(__dadd_rn(__dadd_rn(bl, br), tl), tr);
    const int n = tri_all_in(bl, br, ce) + tri_all_in(br, tr, ce) + tri_all_in(tr, tl, ce) + tri_all_in(tl, bl, ce);
    return 0.25 * (double)n;
}

__global__ void __launch_bounds__(kPT) solidfrac3d_kernel(int nx, int ny, int nz, const double* __restrict__ sphi,
                                                          double* __restrict__ wx, double* __restrict__ wy, double* __restrict__ wz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nz) return;
    const int z = (int)(i % nz);
    const int y = (int)((i / nz) % ny);
    const int x = (int)(i / ((long long)nz * ny));
    const long long fz = 1, fy = 2LL * nz + 1, fx = fy * (2LL * ny + 1);
    const double* p = sphi + 2LL * x * fx + 2LL * y * fy + 2LL * z;
    const double blb = p[0], brb = p[2 * fx], tlb = p[2 * fy], trb = p[2 * fx + 2 * fy];
    const double blf = p[2 * fz], brf = p[2 * fx + 2 * fz], tlf = p[2 * fy + 2 * fz];
    wx[((long long)x * ny + y) * nz + z] = 1.0 - face_in_fraction(tlb, blb, tlf, blf);          // SolidFraction3D.py:22
    wy[((long long)x * (ny + 1) + y) * nz + z] = 1.0 - face_in_fraction(brb, blb, brf, blf);    // :24
    wz[((long long)x * ny + y) * (nz + 1) + z] = 1.0 - face_in_fraction(trb, tlb, brb, blb);    // :26
}

// SolidFraction2D.py:6-20: threads x<=W-2, y<=H-2 write wx[x],wx[x+1],wy[.,y],wy[.,y+1] with identical values
// where they overlap; here each written entry is produced exactly once.
__global__ void __launch_bounds__(kPT) solidfrac2d_kernel(int W, int H, const double* __restrict__ sphi, double* __restrict__ wx, double* __restrict__ wy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)W * H) return;
    const int y = (int)(i % H), x = (int)(i / H);
    const long long fy = 1, fx = 2LL * H + 1;
    const double* p = sphi + 2LL * x * fx + 2LL * y * fy;
    if (y <= H - 2) wx[(long long)x * H + y] = 1.0 - edge_in_fraction(p[2 * fy], p[0]);          // wx[X,y], X in 0..W-1
    if (x <= W - 2) wy[(long long)x * (H + 1) + y] = 1.0 - edge_in_fraction(p[2 * fx], p[0]);    // wy[x,Y], Y in 0..H-1
}

}  // namespace fs

// =================================================================================================
// host side / C ABI
// =================================================================================================
using namespace fs;

struct fs_press {
    int nx, ny, nz;   // nz == 0 -> 2-D
    long long ncells;
    double* partials;
    CgState* st;
    CgHost cg;
    int grid;
    IterGraph graph;
    const void* gkey[9];   // pointer set the captured graph was built for
    uint8_t* act;          // computed-row flag per cell
    SegList seg;           // active 32-cell segments of the current solve
    GridBar* bar;
    bool use_list;         // the current solve iterates on the active list with the persistent kernel
    int op;                // OP_PRESSURE / OP_DENSITY
};

struct PressLayout { size_t st, act, seglist, segscratch, bar, total; };

static PressLayout press_layout(long long ncells) {
    PressLayout o;
    int grid = (int)((ncells + kPT - 1) / kPT);
    size_t np = (size_t)(grid > kVecGrid ? grid : kVecGrid);
    size_t p = align_up(np * sizeof(double), 256);
    o.st = p; p += align_up(sizeof(CgState), 256);
    o.act = p; p = align_up(p + (size_t)ncells + 64, 256);
    o.seglist = p; p = align_up(p + SegList::list_bytes(ncells), 256);
    o.segscratch = p; p = align_up(p + SegList::scratch_bytes(ncells), 256);
    o.bar = p; p = align_up(p + sizeof(GridBar), 256);
    o.total = p;
    return o;
}

template <int D> static PressW<D> mkW(const double* wx, const double* wy, const double* wz) {
    PressW<D> W;
    W.w[0] = wx; W.w[1] = wy;
    if (D == 3) W.w[D - 1] = wz;
    return W;
}

template <typename S, int D> static Vel<D, S> mkV(S* vx, S* vy, S* vz) {
    Vel<D, S> V;
    V.v[0] = vx; V.v[1] = vy;
    if (D == 3) V.v[D - 1] = vz;
    return V;
}

static int press_apply_launch(fs_press* h, const double* v, double* out, const double* wx, const double* wy, const double* wz,
                              const double* lphi, bool cg, cudaStream_t s) {
    const int pg = h->grid < kSMs * kPressBlocksPerSM ? h->grid : kSMs * kPressBlocksPerSM;
    if (h->nz > 0) {
        auto g = make_grid<3>(h->nx, h->ny, h->nz);
        if (h->op == OP_DENSITY) {
            if (cg) press_apply_kernel<3, true, OP_DENSITY><<<pg, kPT, 0, s>>>(g, v, out, mkW<3>(wx, wy, wz), lphi, h->st, h->partials);
            else press_apply_kernel<3, false, OP_DENSITY><<<pg, kPT, 0, s>>>(g, v, out, mkW<3>(wx, wy, wz), lphi, h->st, h->partials);
        } else {
            if (cg) press_apply_kernel<3, true, OP_PRESSURE><<<pg, kPT, 0, s>>>(g, v, out, mkW<3>(wx, wy, wz), lphi, h->st, h->partials);
            else press_apply_kernel<3, false, OP_PRESSURE><<<pg, kPT, 0, s>>>(g, v, out, mkW<3>(wx, wy, wz), lphi, h->st, h->partials);
        }
    } else {
        auto g = make_grid<2>(h->nx, h->ny, 0);
        if (cg) press_apply_kernel<2, true, OP_PRESSURE><<<pg, kPT, 0, s>>>(g, v, out, mkW<2>(wx, wy, nullptr), lphi, h->st, h->partials);
        else press_apply_kernel<2, false, OP_PRESSURE><<<pg, kPT, 0, s>>>(g, v, out, mkW<2>(wx, wy, nullptr), lphi, h->st, h->partials);
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

static bool press_same_ptrs(fs_press* h, const void* const* p) {
    for (int k = 0; k < 9; ++k) if (h->gkey[k] != p[k]) return false;
    return true;
}

static int press_iteration(fs_press* h, double* x, double* d, double* r, double* q, const double* wx, const double* wy, const double* wz,
                           const double* lphi, cudaStream_t s) {
    FS_TRY(press_apply_launch(h, d, q, wx, wy, wz, lphi, true, s));
    FS_TRY((cg_launch_update_xr<double>(h->ncells, x, r, d, q, h->st, h->partials, s)));
    FS_TRY((cg_launch_update_d<double>(h->ncells, d, r, h->st, s)));
    return FS_OK;
}

// Build the computed-row map and the active segment list of this solve; decide whether the iterations run on the list
// (persistent kernel; FLUIDSOLVER_B200_PERSISTENT=0/1 forces off/on, default: active working set <= 256 MB) or on the
// dense three-kernel path.  The list kernels use 16-byte accesses: an odd cell count or unaligned arrays stay dense.
static int press_prepare_list(fs_press* h, const double* x, const double* d, const double* r, const double* q, const double* lphi, cudaStream_t s) {
    h->use_list = false;
    static int mode = -2;
    if (mode == -2) {
        const char* e = getenv("FLUIDSOLVER_B200_PERSISTENT");
        mode = !e ? -1 : (e[0] == '0' ? 0 : 1);
    }
    if (mode == 0) return FS_OK;
    if ((h->ncells & 1) || !aligned16(x) || !aligned16(d) || !aligned16(r) || !aligned16(q)) return FS_OK;
    if (h->nz > 0) press_activity_kernel<3><<<h->grid, kPT, 0, s>>>(make_grid<3>(h->nx, h->ny, h->nz), lphi, h->act);
    else press_activity_kernel<2><<<h->grid, kPT, 0, s>>>(make_grid<2>(h->nx, h->ny, 0), lphi, h->act);
    FS_LAUNCH_CHECK();
    FS_TRY(h->seg.build(h->act, s));
    const double ws = (double)h->seg.nseg * kSegPts * 80.0;     // x,r,d,q,lphi + weights, 8 bytes each
    h->use_list = (mode == 1) || ws <= 256e6;
    return FS_OK;
}

static int press_persistent(fs_press* h, double* x, double* d, double* r, double* q, const double* wx, const double* wy, const double* wz,
                            const double* lphi, long long n, cudaStream_t s) {
    const void* fn3 = h->op == OP_DENSITY ? (const void*)press_cg_persistent_kernel<3, OP_DENSITY> : (const void*)press_cg_persistent_kernel<3, OP_PRESSURE>;
    const void* fn = h->nz > 0 ? fn3 : (const void*)press_cg_persistent_kernel<2, OP_PRESSURE>;
    const int cap = coop_max_blocks(fn, kPersistThreads);      // SMs of this context x resident CTAs per SM
    if (cap < 1) { h->use_list = false; return 1; }
    const int grid = seg_grid(h->seg.nseg, kPersistThreads / 32, cap < kSMs ? cap : kSMs);
    while (n > 0) {
        int ni = (int)(n < (1 << 20) ? n : (1 << 20));
        cudaError_t e = cudaMemsetAsync(h->bar, 0, sizeof(GridBar), s);
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        const int* seg = h->seg.list; const int* nsegp = h->seg.nseg_dev;
        CgState* st = h->st; double* partials = h->partials; GridBar* bar = h->bar;
        if (h->nz > 0) {
            Grid<3> g = make_grid<3>(h->nx, h->ny, h->nz);
            PressW<3> W = mkW<3>(wx, wy, wz);
            void* args[] = {&g, &x, &r, &d, &q, &W, &lphi, &seg, &nsegp, &st, &partials, &bar, &ni};
            e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPersistThreads), args, 0, s);
        } else {
            Grid<2> g = make_grid<2>(h->nx, h->ny, 0);
            PressW<2> W = mkW<2>(wx, wy, nullptr);
            void* args[] = {&g, &x, &r, &d, &q, &W, &lphi, &seg, &nsegp, &st, &partials, &bar, &ni};
            e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPersistThreads), args, 0, s);
        }
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
            cudaGetLastError();                      // this context cannot hold the cooperative grid: dense three-kernel path
            h->use_list = false;
            return 1;
        }
        if (e != cudaSuccess) return fail(FS_ERR_CUDA, "cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
        FS_LAUNCH_CHECK();
        n -= ni;
    }
    return FS_OK;
}

static int press_iterations(fs_press* h, double* x, double* d, double* r, double* q, const double* wx, const double* wy, const double* wz,
                            const double* lphi, long long n, cudaStream_t s) {
    if (h->use_list) {
        const int st = press_persistent(h, x, d, r, q, wx, wy, wz, lphi, n, s);
        if (st != 1) return st;                      // 1 = cooperative launch impossible here, nothing was enqueued
    }
    const void* ptrs[9] = {x, d, r, q, wx, wy, wz, lphi, nullptr};
    if (!press_same_ptrs(h, ptrs)) {                      // the graph bakes the array pointers in
        h->graph.valid = false;
        for (int k = 0; k < 9; ++k) h->gkey[k] = ptrs[k];
    }
    return cg_enqueue_iterations(h->graph, true, 1.0, n, [&](cudaStream_t ss) { return press_iteration(h, x, d, r, q, wx, wy, wz, lphi, ss); }, s);
}

extern "C" {

size_t fs_press_workspace_bytes(int nx, int ny, int nz) {
    if (nx < 1 || ny < 1 || nz < 0) return 0;
    return press_layout((long long)nx * ny * (nz > 0 ? nz : 1)).total;
}

int fs_press_create(fs_press** out, int nx, int ny, int nz, void* ws, size_t ws_bytes) {
    if (!out || !ws) return fail(FS_ERR_ARG, "fs_press_create: null argument");
    if (nx < 1 || ny < 1 || nz < 0) return fail(FS_ERR_ARG, "fs_press_create: bad grid resolution");
    if ((uintptr_t)ws % 256) return fail(FS_ERR_ARG, "fs_press_create: workspace must be 256-byte aligned");
    fs_press* h = new fs_press();
    h->nx = nx; h->ny = ny; h->nz = nz;
    h->ncells = (long long)nx * ny * (nz > 0 ? nz : 1);
    const PressLayout lay = press_layout(h->ncells);
    const size_t need = lay.total;
    if (ws_bytes < need) { delete h; return fail(FS_ERR_ARG, "fs_press_create: workspace too small"); }
    h->partials = (double*)ws;
    h->st = (CgState*)((char*)ws + lay.st);
    h->act = (uint8_t*)((char*)ws + lay.act);
    h->bar = (GridBar*)((char*)ws + lay.bar);
    h->use_list = false;
    h->grid = (int)((h->ncells + kPT - 1) / kPT);
    for (int k = 0; k < 9; ++k) h->gkey[k] = nullptr;
    int s = h->cg.init();
    if (s < 0) { delete h; return s; }
    s = h->seg.init(h->ncells, (char*)ws + lay.seglist, (char*)ws + lay.segscratch);
    if (s < 0) { h->cg.destroy(); delete h; return s; }
    h->cg.st_dev = h->st; h->cg.partials_dev = h->partials;
    cudaError_t e = cudaMemset(ws, 0, need);
    if (e != cudaSuccess) { h->cg.destroy(); h->seg.destroy(); delete h; return fail(FS_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    *out = h;
    return FS_OK;
}

void fs_press_destroy(fs_press* h) {
    if (!h) return;
    h->graph.destroy();
    h->cg.destroy();
    h->seg.destroy();
    delete h;
}

int fs_press_set_operator(fs_press* h, int op) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (op != FS_OP_PRESSURE && op != FS_OP_DENSITY) return fail(FS_ERR_ARG, "fs_press_set_operator: bad operator");
    if (op == FS_OP_DENSITY && h->nz == 0) return fail(FS_ERR_ARG, "fs_press_set_operator: the density operator is 3-D only");
    h->op = op;
    h->graph.valid = false;
    return FS_OK;
}

int fs_press_apply(fs_press* h, const double* v, double* out, const double* wx, const double* wy, const double* wz, const double* lphi, void* stream) {
    if (!h || !v || !out || !wx || !wy || !lphi || (h->nz > 0 && !wz)) return fail(FS_ERR_ARG, "fs_press_apply: null argument");
    return press_apply_launch(h, v, out, wx, wy, wz, lphi, false, (cudaStream_t)stream);
}

int fs_press_rhs(fs_press* h, const double* cell_size3, const void* vx, const void* vy, const void* vz, int vel_dtype,
                 const double* sv, const double* lphi, double* b, const double* wx, const double* wy, const double* wz, void* stream) {
    if (!h || !cell_size3 || !vx || !vy || !sv || !lphi || !b || !wx || !wy || (h->nz > 0 && (!vz || !wz)))
        return fail(FS_ERR_ARG, "fs_press_rhs: null argument");
    if (vel_dtype != FS_F32 && vel_dtype != FS_F64) return fail(FS_ERR_ARG, "fs_press_rhs: bad velocity dtype");
    cudaStream_t s = (cudaStream_t)stream;
    const double c0 = cell_size3[0], c1 = cell_size3[1], c2 = h->nz > 0 ? cell_size3[2] : 1.0;
    if (h->nz > 0) {
        auto g = make_grid<3>(h->nx, h->ny, h->nz);
        if (vel_dtype == FS_F32) press_rhs_kernel<3, float><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<const float, 3>((const float*)vx, (const float*)vy, (const float*)vz), sv, lphi, b, mkW<3>(wx, wy, wz));
        else press_rhs_kernel<3, double><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<const double, 3>((const double*)vx, (const double*)vy, (const double*)vz), sv, lphi, b, mkW<3>(wx, wy, wz));
    } else {
        auto g = make_grid<2>(h->nx, h->ny, 0);
        if (vel_dtype == FS_F32) press_rhs_kernel<2, float><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<const float, 2>((const float*)vx, (const float*)vy, nullptr), sv, lphi, b, mkW<2>(wx, wy, nullptr));
        else press_rhs_kernel<2, double><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<const double, 2>((const double*)vx, (const double*)vy, nullptr), sv, lphi, b, mkW<2>(wx, wy, nullptr));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_press_update(fs_press* h, const double* cell_size3, void* vx, void* vy, void* vz, int vel_dtype, const double* pv,
                    const double* wx, const double* wy, const double* wz, const double* sv, const double* lphi, void* stream) {
    if (!h || !cell_size3 || !vx || !vy || !pv || !sv || !lphi || !wx || !wy || (h->nz > 0 && (!vz || !wz)))
        return fail(FS_ERR_ARG, "fs_press_update: null argument");
    if (vel_dtype != FS_F32 && vel_dtype != FS_F64) return fail(FS_ERR_ARG, "fs_press_update: bad velocity dtype");
    cudaStream_t s = (cudaStream_t)stream;
    const double c0 = cell_size3[0], c1 = cell_size3[1], c2 = h->nz > 0 ? cell_size3[2] : 1.0;
    if (h->nz > 0) {
        auto g = make_grid<3>(h->nx, h->ny, h->nz);
        if (vel_dtype == FS_F32) press_update_kernel<3, float><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<float, 3>((float*)vx, (float*)vy, (float*)vz), pv, mkW<3>(wx, wy, wz), sv, lphi);
        else press_update_kernel<3, double><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<double, 3>((double*)vx, (double*)vy, (double*)vz), pv, mkW<3>(wx, wy, wz), sv, lphi);
    } else {
        auto g = make_grid<2>(h->nx, h->ny, 0);
        if (vel_dtype == FS_F32) press_update_kernel<2, float><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<float, 2>((float*)vx, (float*)vy, nullptr), pv, mkW<2>(wx, wy, nullptr), sv, lphi);
        else press_update_kernel<2, double><<<h->grid, kPT, 0, s>>>(g, c0, c1, c2, mkV<double, 2>((double*)vx, (double*)vy, nullptr), pv, mkW<2>(wx, wy, nullptr), sv, lphi);
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_press_cg(fs_press* h, double* x, double* d, double* r, double* q, const double* b,
                const double* wx, const double* wy, const double* wz, const double* lphi,
                double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h || !x || !d || !r || !q || !b || !wx || !wy || !lphi || (h->nz > 0 && !wz)) return fail(FS_ERR_ARG, "fs_press_cg: null argument");
    if (max_iter < 0) return fail(FS_ERR_ARG, "fs_press_cg: max_iter < 0");
    cudaStream_t s = (cudaStream_t)stream;
    cg_state_init_kernel<<<1, 1, 0, s>>>(h->st, tol * tol, (long long)max_iter);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaMemsetAsync(x, 0, h->ncells * sizeof(double), s));                       // self.x *= 0.0   (:198)
    FS_TRY(press_apply_launch(h, x, q, wx, wy, wz, lphi, false, s));                     // q = A x         (:201)
    FS_TRY((cg_launch_residual_init<double>(h->ncells, b, q, d, r, h->st, h->partials, s)));   // d = b - q ; r = d ; delta (:202-204)
    FS_TRY(press_prepare_list(h, x, d, r, q, lphi, s));
    return cg_drive(h->cg, [&](cudaStream_t ss, long long nb) { return press_iterations(h, x, d, r, q, wx, wy, wz, lphi, nb, ss); },
                    (long long)max_iter, stats, s, h->use_list ? kCgBatchPersistent : kCgBatch);
}

int fs_press_cg_enqueue(fs_press* h, double* x, double* d, double* r, double* q,
                        const double* wx, const double* wy, const double* wz, const double* lphi, int64_t n, void* stream) {
    if (!h || !x || !d || !r || !q || !wx || !wy || !lphi || (h->nz > 0 && !wz)) return fail(FS_ERR_ARG, "fs_press_cg_enqueue: null argument");
    cg_state_unlimit_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->st);
    FS_LAUNCH_CHECK();
    FS_TRY(press_prepare_list(h, x, d, r, q, lphi, (cudaStream_t)stream));
    return press_iterations(h, x, d, r, q, wx, wy, wz, lphi, n, (cudaStream_t)stream);
}

int fs_solidfrac3d(int nx, int ny, int nz, const double* sphi, double* wx, double* wy, double* wz, void* stream) {
    if (!sphi || !wx || !wy || !wz) return fail(FS_ERR_ARG, "fs_solidfrac3d: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_solidfrac3d: bad grid resolution");
    const long long n = (long long)nx * ny * nz;
    solidfrac3d_kernel<<<(unsigned)((n + kPT - 1) / kPT), kPT, 0, (cudaStream_t)stream>>>(nx, ny, nz, sphi, wx, wy, wz);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_solidfrac2d(int W, int H, const double* sphi, double* wx, double* wy, void* stream) {
    if (!sphi || !wx || !wy) return fail(FS_ERR_ARG, "fs_solidfrac2d: null argument");
    if (W < 1 || H < 1) return fail(FS_ERR_ARG, "fs_solidfrac2d: bad grid resolution");
    if (W < 2 || H < 2) return FS_OK;        // the reference launches no writing thread (x >= W-1 or y >= H-1 return)
    const long long n = (long long)W * H;
    solidfrac2d_kernel<<<(unsigned)((n + kPT - 1) / kPT), kPT, 0, (cudaStream_t)stream>>>(W, H, sphi, wx, wy);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

}  // extern "C"
