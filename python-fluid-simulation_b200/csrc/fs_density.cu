// Particle-density (volume-conservation) solve: the grid-side and particle-side kernels around the density CG, sm_100a.
//
// Replaces DensityCGSolver3D.py: initialize_density_kernel :8-36 (particle -> cell scatter of mass and volume),
// fix_volume_kernel :38-92, initialize_solver_kernel :94-125, compute_displacement_kernel :206-219 and
// apply_displacement_kernel :221-248 (cell-face displacement -> particle gather).  The operator apply and the CG loop
// (:117-204, :313-343) run through fs_press with FS_OP_DENSITY (fs_press.cu), i.e. on the active cell set with the
// persistent whole-iteration kernel.
//
// Arrays are the reference's dense C-order fp64 arrays (cells (nx,ny,nz), face arrays with +1 extent on their own
// axis, fine grid (2n+1)^3); particles px (P,3) fp32/fp64, pm (P) fp32/fp64.  Arithmetic follows the reference's
// association with explicit round-to-nearest intrinsics (no FMA contraction).  The scatter accumulates with fp64
// atomics like the reference, so cell mass / volume are reproducible to summation-order rounding only.
#include "fs_common.cuh"

namespace fs {

constexpr int kDT = 256;

struct Dens3 {
    int nx, ny, nz;
    long long ncells;
    double bmin[3], cs[3];
};

// trilinear stencil of one particle on a grid with per-axis bias (0.5 = cell centres, 0 = faces of that axis)
struct Tri {
    long long gi[3];
    double w[3];
};

template <typename S>
__device__ __forceinline__ Tri tri_setup(const S* __restrict__ px, long long P, const double* bmin, const double* cs, const double* bias) {
    Tri t;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double x = (double)px[P * 3 + d];
        const double g = floor(__dsub_rn(__dsub_rn(x, bmin[d]) / cs[d], bias[d]));            // :14 / :232
        t.gi[d] = (long long)g;
        const double gx = __dadd_rn(__dmul_rn(__dadd_rn(g, bias[d]), cs[d]), bmin[d]);         // :15 / :233
        t.w[d] = fabs(__dsub_rn(gx, x)) / cs[d];                                               // :16 / :234
    }
    return t;
}

__device__ __forceinline__ double corner_w(int i, double w) {            // i + (-1)^i (1 - w)   (:24-26)
    const double o = __dsub_rn(1.0, w);
    return i ? __dadd_rn(1.0, -o) : o;
}

__device__ __forceinline__ long long clampll(long long v, long long lo, long long hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename SX, typename SM>
__global__ void __launch_bounds__(kDT) dens_scatter_kernel(Dens3 G, const SX* __restrict__ px, const SM* __restrict__ pm, long long np, double pvol,
                                                           double* __restrict__ gm, double* __restrict__ gvol) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= np) return;
    const double bias[3] = {0.5, 0.5, 0.5};
    const Tri t = tri_setup(px, P, G.bmin, G.cs, bias);
    const double m = (double)pm[P];
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
                const long long cx = clampll(t.gi[0] + ix, 0, G.nx - 1), cy = clampll(t.gi[1] + iy, 0, G.ny - 1), cz = clampll(t.gi[2] + iz, 0, G.nz - 1);
                const double weight = __dmul_rn(__dmul_rn(corner_w(ix, t.w[0]), corner_w(iy, t.w[1])), corner_w(iz, t.w[2]));
                const long long c = (cx * G.ny + cy) * G.nz + cz;
                atomicAdd(gm + c, __dmul_rn(weight, m));
                atomicAdd(gvol + c, __dmul_rn(weight, pvol));
            }
}

__device__ __forceinline__ bool dens_decode_interior(const Dens3& G, long long i, int& x, int& y, int& z) {
    z = (int)(i % G.nz);
    const long long t = i / G.nz;
    y = (int)(t % G.ny);
    x = (int)(t / G.ny);
    return x >= 1 && x <= G.nx - 2 && y >= 1 && y <= G.ny - 2 && z >= 1 && z <= G.nz - 2;
}

// (wx[x]+wx[x+1]+wy[y]+wy[y+1]+wz[z]+wz[z+1]) / 6, left to right (:82-89, :106-113)
__device__ __forceinline__ double open_fraction(const Dens3& G, long long i, int x, int y, const double* __restrict__ wx,
                                                const double* __restrict__ wy, const double* __restrict__ wz) {
    const long long iy = i + (long long)x * G.nz;            // wy has ny+1 rows per x-plane
    const long long iz = i + ((long long)x * G.ny + y);      // wz has nz+1 entries per row
    double s = __dadd_rn(wx[i], wx[i + (long long)G.ny * G.nz]);
    s = __dadd_rn(s, wy[iy]);
    s = __dadd_rn(s, wy[iy + G.nz]);
    s = __dadd_rn(s, wz[iz]);
    s = __dadd_rn(s, wz[iz + 1]);
    return s / 6.0;
}

__global__ void __launch_bounds__(kDT) dens_fix_volume_kernel(Dens3 G, double cvol, double dx, double* __restrict__ gvol, const double* __restrict__ sphi,
                                                              const double* __restrict__ lphi, const double* __restrict__ wx,
                                                              const double* __restrict__ wy, const double* __restrict__ wz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.ncells) return;
    int x, y, z;
    if (!dens_decode_interior(G, i, x, y, z)) return;
    double fluid_vol = gvol[i];
    const long long fy = 2LL * G.nz + 1, fx = fy * (2LL * G.ny + 1);
    const bool near_solid = sphi[(2LL * x + 1) * fx + (2LL * y + 1) * fy + (2LL * z + 1)] < dx;
    const long long sx = (long long)G.ny * G.nz, sy = G.nz;
    const bool internal = lphi[i] < 0 && lphi[i + sx] < 0 && lphi[i - sx] < 0 && lphi[i + sy] < 0 && lphi[i - sy] < 0 && lphi[i + 1] < 0 && lphi[i - 1] < 0;
    if (internal && !near_solid) fluid_vol = cvol;
    const double frac = open_fraction(G, i, x, y, wx, wy, wz);
    gvol[i] = fmin(fluid_vol, __dmul_rn(cvol, frac));
}

__global__ void __launch_bounds__(kDT) dens_rhs_kernel(Dens3 G, double rho0, double cvol, double dt, const double* __restrict__ gm,
                                                       const double* __restrict__ gvol, const double* __restrict__ lphi, const double* __restrict__ wx,
                                                       const double* __restrict__ wy, const double* __restrict__ wz, double* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.ncells) return;
    int x, y, z;
    if (!dens_decode_interior(G, i, x, y, z)) return;
    if (!(lphi[i] < 0)) { b[i] = 0.0; return; }
    const double frac = open_fraction(G, i, x, y, wx, wy, wz);
    const double solid_vol = __dmul_rn(__dsub_rn(1.0, frac), cvol);
    const double solid_mass = __dmul_rn(rho0, solid_vol);
    const double cell_mass = __dadd_rn(gm[i], solid_mass);
    const double cell_vol = __dadd_rn(gvol[i], solid_vol);
    double dens = (cell_mass / fmax(cell_vol, 1e-10)) / rho0;
    if (cell_mass < 1e-10) dens = 1.0;
    dens = fmax(0.5, fmin(1.5, dens));
    b[i] = __dsub_rn(1.0, dens) / dt;
}

__device__ __forceinline__ double dens_edge_in_fraction(double l, double r) {      // SolidFractionCommon.py:4-16
    const bool l_in = l < 0, r_in = r < 0;
    if (l_in && r_in) return 1.0;
    if (!l_in && !r_in) return 0.0;
    const double diff = -fabs(__dsub_rn(l, r));
    return l_in ? l / diff : r / diff;
}

// indices 1..g-1 on all axes, for all three components (:208-219)
__global__ void __launch_bounds__(kDT) dens_displacement_kernel(Dens3 G, double dt, double* __restrict__ dx, double* __restrict__ dy, double* __restrict__ dz,
                                                                const double* __restrict__ pv, const double* __restrict__ lphi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.ncells) return;
    const int z = (int)(i % G.nz);
    const long long t = i / G.nz;
    const int y = (int)(t % G.ny), x = (int)(t / G.ny);
    if (x < 1 || y < 1 || z < 1) return;
    const long long sx = (long long)G.ny * G.nz, sy = G.nz;
    const double phi = lphi[i], p = pv[i];
    const double tx = fmin(1.0, fmax(0.01, dens_edge_in_fraction(phi, lphi[i - sx])));
    const double ty = fmin(1.0, fmax(0.01, dens_edge_in_fraction(phi, lphi[i - sy])));
    const double tz = fmin(1.0, fmax(0.01, dens_edge_in_fraction(phi, lphi[i - 1])));
    dx[i] = __dmul_rn(__dmul_rn(__dsub_rn(p, pv[i - sx]), dt), G.cs[0]) / tx;
    dy[i + (long long)x * G.nz] = __dmul_rn(__dmul_rn(__dsub_rn(p, pv[i - sy]), dt), G.cs[1]) / ty;
    dz[i + ((long long)x * G.ny + y)] = __dmul_rn(__dmul_rn(__dsub_rn(p, pv[i - 1]), dt), G.cs[2]) / tz;
}

template <typename SX>
__global__ void __launch_bounds__(kDT) dens_gather_kernel(SX* __restrict__ px, long long np, const double* __restrict__ d, int s0, int s1, int s2,
                                                          double b0, double b1, double b2, double c0, double c1, double c2,
                                                          double g0, double g1, double g2, int axis) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= np) return;
    const double bmin[3] = {b0, b1, b2}, cs[3] = {c0, c1, c2}, bias[3] = {g0, g1, g2};
    const Tri t = tri_setup(px, P, bmin, cs, bias);
    SX acc = px[P * 3 + axis];
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
                const long long cx = clampll(t.gi[0] + ix, 0, s0 - 1), cy = clampll(t.gi[1] + iy, 0, s1 - 1), cz = clampll(t.gi[2] + iz, 0, s2 - 1);
                const double weight = __dmul_rn(__dmul_rn(corner_w(ix, t.w[0]), corner_w(iy, t.w[1])), corner_w(iz, t.w[2]));
                // px[P,axis] += weight * dx[...] : accumulated in the particle array's own type, one corner at a time (:248)
                acc = (SX)__dadd_rn((double)acc, __dmul_rn(weight, d[(cx * s1 + cy) * s2 + cz]));
            }
    px[P * 3 + axis] = acc;
}

static Dens3 mkdens(int nx, int ny, int nz, const double* bmin, const double* cs) {
    Dens3 G;
    G.nx = nx; G.ny = ny; G.nz = nz;
    G.ncells = (long long)nx * ny * nz;
    for (int d = 0; d < 3; ++d) { G.bmin[d] = bmin ? bmin[d] : 0.0; G.cs[d] = cs ? cs[d] : 1.0; }
    return G;
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_dens3d_scatter(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, const void* px, int px_dtype,
                      const void* pm, int pm_dtype, int64_t np, double pvol, double* gm, double* gvol, void* stream) {
    if (!bound_min3 || !cell_size3 || !gm || !gvol || (np > 0 && (!px || !pm))) return fail(FS_ERR_ARG, "fs_dens3d_scatter: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || np < 0) return fail(FS_ERR_ARG, "fs_dens3d_scatter: bad sizes");
    if ((px_dtype != FS_F32 && px_dtype != FS_F64) || (pm_dtype != FS_F32 && pm_dtype != FS_F64)) return fail(FS_ERR_ARG, "fs_dens3d_scatter: bad dtype");
    if (np == 0) return FS_OK;
    const Dens3 G = mkdens(nx, ny, nz, bound_min3, cell_size3);
    const unsigned grid = (unsigned)((np + kDT - 1) / kDT);
    cudaStream_t s = (cudaStream_t)stream;
    if (px_dtype == FS_F32 && pm_dtype == FS_F32) dens_scatter_kernel<float, float><<<grid, kDT, 0, s>>>(G, (const float*)px, (const float*)pm, np, pvol, gm, gvol);
    else if (px_dtype == FS_F32) dens_scatter_kernel<float, double><<<grid, kDT, 0, s>>>(G, (const float*)px, (const double*)pm, np, pvol, gm, gvol);
    else if (pm_dtype == FS_F32) dens_scatter_kernel<double, float><<<grid, kDT, 0, s>>>(G, (const double*)px, (const float*)pm, np, pvol, gm, gvol);
    else dens_scatter_kernel<double, double><<<grid, kDT, 0, s>>>(G, (const double*)px, (const double*)pm, np, pvol, gm, gvol);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_dens3d_fix_volume(int nx, int ny, int nz, const double* cell_size3, double* gvol, const double* sphi, const double* lphi,
                         const double* wx, const double* wy, const double* wz, void* stream) {
    if (!cell_size3 || !gvol || !sphi || !lphi || !wx || !wy || !wz) return fail(FS_ERR_ARG, "fs_dens3d_fix_volume: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_dens3d_fix_volume: bad sizes");
    const Dens3 G = mkdens(nx, ny, nz, nullptr, cell_size3);
    const double cvol = cell_size3[0] * cell_size3[1] * cell_size3[2];                     // cp.prod(cell_size)
    const double dx = fmin(cell_size3[0], fmin(cell_size3[1], cell_size3[2]));            // cp.min(cell_size)
    dens_fix_volume_kernel<<<(unsigned)((G.ncells + kDT - 1) / kDT), kDT, 0, (cudaStream_t)stream>>>(G, cvol, dx, gvol, sphi, lphi, wx, wy, wz);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_dens3d_rhs(int nx, int ny, int nz, double rho0, double dt, const double* cell_size3, const double* gm, const double* gvol,
                  const double* lphi, const double* wx, const double* wy, const double* wz, double* b, void* stream) {
    if (!cell_size3 || !gm || !gvol || !lphi || !wx || !wy || !wz || !b) return fail(FS_ERR_ARG, "fs_dens3d_rhs: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_dens3d_rhs: bad sizes");
    const Dens3 G = mkdens(nx, ny, nz, nullptr, cell_size3);
    const double cvol = cell_size3[0] * cell_size3[1] * cell_size3[2];
    dens_rhs_kernel<<<(unsigned)((G.ncells + kDT - 1) / kDT), kDT, 0, (cudaStream_t)stream>>>(G, rho0, cvol, dt, gm, gvol, lphi, wx, wy, wz, b);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_dens3d_displacement(int nx, int ny, int nz, double dt, const double* cell_size3, double* dx, double* dy, double* dz,
                           const double* pv, const double* lphi, void* stream) {
    if (!cell_size3 || !dx || !dy || !dz || !pv || !lphi) return fail(FS_ERR_ARG, "fs_dens3d_displacement: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_dens3d_displacement: bad sizes");
    const Dens3 G = mkdens(nx, ny, nz, nullptr, cell_size3);
    dens_displacement_kernel<<<(unsigned)((G.ncells + kDT - 1) / kDT), kDT, 0, (cudaStream_t)stream>>>(G, dt, dx, dy, dz, pv, lphi);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_dens3d_gather(void* px, int px_dtype, int64_t np, const double* d, int s0, int s1, int s2, const double* bound_min3,
                     const double* cell_size3, const double* grid_bias3, int axis, void* stream) {
    if (!d || !bound_min3 || !cell_size3 || !grid_bias3 || (np > 0 && !px)) return fail(FS_ERR_ARG, "fs_dens3d_gather: null argument");
    if (s0 < 1 || s1 < 1 || s2 < 1 || np < 0 || axis < 0 || axis > 2) return fail(FS_ERR_ARG, "fs_dens3d_gather: bad sizes");
    if (px_dtype != FS_F32 && px_dtype != FS_F64) return fail(FS_ERR_ARG, "fs_dens3d_gather: bad dtype");
    if (np == 0) return FS_OK;
    const unsigned grid = (unsigned)((np + kDT - 1) / kDT);
    cudaStream_t s = (cudaStream_t)stream;
    const double* b = bound_min3; const double* c = cell_size3; const double* g = grid_bias3;
    if (px_dtype == FS_F32) dens_gather_kernel<float><<<grid, kDT, 0, s>>>((float*)px, np, d, s0, s1, s2, b[0], b[1], b[2], c[0], c[1], c[2], g[0], g[1], g[2], axis);
    else dens_gather_kernel<double><<<grid, kDT, 0, s>>>((double*)px, np, d, s0, s1, s2, b[0], b[1], b[2], c[0], c[1], c[2], g[0], g[1], g[2], axis);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

}  // extern "C"
