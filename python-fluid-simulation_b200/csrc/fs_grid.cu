// Grid-side steps of the APIC time loop either side of the implicit solves (sm_100a):
// particle -> grid transfer, grid -> particle transfer, fluid level set, fluid volume splat, velocity extrapolation with
// mass validity, solid boundary condition.
//
// Reference: code cells of 3D_viscous_fluid_sim.ipynb — p2g_particle / p2g_grid (:279-344), g2p_particle (:352-393),
// compute_fls_kernel (:94-136), compute_fluid_volume_kernel / constrain_fluid_volume_kernel (:224-268),
// extrapolate_kernel / extrapolate (:501-557), boundary_condition_{x,y,z} / apply_boundary_condition (:405-565, cell 5).
// They produce the lvol / lphi / velocity / mass arrays the solvers consume and take their output back to the particles.
//
// Data types are the notebook's: particle arrays fp64 (x, v, c* as (P,3), m as (P)), MAC grids fp32, level-set and
// volume grids fp64; particle-local quantities are fp32 exactly where the notebook declares fp32 local arrays.
// Not a translation: one fused launch handles all three MAC components (the notebook launches every kernel three times
// and synchronises in between), the extrapolation runs in place with a generation byte instead of ping-pong copies of
// six arrays per sweep, and the fp64 atomic min of the level set is a single integer atomic per sample.
#include "fs_common.cuh"

namespace fs {

struct GridGeom {
    int n[3];            // cells
    float bmin[3];       // bound_min (fp32 in the notebook)
    double cell[3];      // cell_size = bound_size / gres (fp64 in the notebook: float32 / int64)
};

__device__ __forceinline__ int clampi(long long v, int hi) { return v < 0 ? 0 : (v > hi ? hi : (int)v); }

// particle -> (base index, fp32 offsets, fp32 weights) on the lattice with bias `bias` (p2g_particle :287-299)
struct Stencil {
    long long gi[3];
    float disp[3], w[3];
};
__device__ __forceinline__ Stencil particle_stencil(const GridGeom& G, const double* __restrict__ px, long long P, const float* bias) {
    Stencil s;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float x = (float)px[P * 3 + d];
        const double t = (double)(x - G.bmin[d]) / G.cell[d] - (double)bias[d];
        s.gi[d] = (long long)floor(t);
        const float gx = (float)(((double)s.gi[d] + (double)bias[d]) * G.cell[d] + (double)G.bmin[d]);
        s.disp[d] = gx - x;
        s.w[d] = (float)((double)fabsf(s.disp[d]) / G.cell[d]);
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// P2G: one thread per (particle, MAC component).  gm += w m ; gv += w m (v_axis + C_axis . (x_node - x_p))
// ---------------------------------------------------------------------------------------------
struct MacF32 { float* m[3]; float* v[3]; };

__global__ void __launch_bounds__(256) grid_p2g_kernel(GridGeom G, long long np, const double* __restrict__ px, const double* __restrict__ pm,
                                                       const double* __restrict__ pv, const double* __restrict__ cx, const double* __restrict__ cy,
                                                       const double* __restrict__ cz, MacF32 g) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np * 3) return;
    const long long P = t / 3;
    const int axis = (int)(t - P * 3);
    float bias[3] = {0.5f, 0.5f, 0.5f};
    bias[axis] = 0.0f;
    const Stencil s = particle_stencil(G, px, P, bias);
    const double* pca = axis == 0 ? cx : (axis == 1 ? cy : cz);
    const double m = pm[P];
    const float va = (float)pv[P * 3 + axis];
    const double c0 = pca[P * 3 + 0], c1 = pca[P * 3 + 1], c2 = pca[P * 3 + 2];
    int sh[3] = {G.n[0], G.n[1], G.n[2]};
    sh[axis] += 1;
    float* gm = g.m[axis];
    float* gv = g.v[axis];
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
                // index clamp uses gres-1 on every axis, like the reference (:303-305) — the extra plane of the staggered
                // component is never a clamp target
                const int gix = clampi(s.gi[0] + ix, G.n[0] - 1), giy = clampi(s.gi[1] + iy, G.n[1] - 1), giz = clampi(s.gi[2] + iz, G.n[2] - 1);
                const float wx = ix ? s.w[0] : 1.0f - s.w[0];
                const float wy = iy ? s.w[1] : 1.0f - s.w[1];
                const float wz = iz ? s.w[2] : 1.0f - s.w[2];
                const double cv = ((double)s.disp[0] + ix * G.cell[0]) * c0 + ((double)s.disp[1] + iy * G.cell[1]) * c1 + ((double)s.disp[2] + iz * G.cell[2]) * c2;
                const float weight = wx * wy * wz;
                const long long idx = ((long long)gix * sh[1] + giy) * sh[2] + giz;
                atomicAdd(gm + idx, (float)((double)weight * m));
                atomicAdd(gv + idx, (float)((double)weight * m * ((double)va + cv)));
            }
}

// p2g_grid (:322-330): v = v / m where m > 0, all three components in one launch
__global__ void __launch_bounds__(256) grid_normalise_kernel(long long n0, long long n1, long long n2, MacF32 g) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n[3] = {n0, n1, n2};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (i < n[a]) {
            const float m = g.m[a][i];
            if (m > 0.0f) g.v[a][i] = g.v[a][i] / m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// G2P: one thread per (particle, component): v_axis and the affine row C_axis (:352-385)
// ---------------------------------------------------------------------------------------------
struct MacF32c { const float* v[3]; };

__global__ void __launch_bounds__(256) grid_g2p_kernel(GridGeom G, long long np, const double* __restrict__ px, double* __restrict__ pv,
                                                       double* __restrict__ cx, double* __restrict__ cy, double* __restrict__ cz, MacF32c g) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np * 3) return;
    const long long P = t / 3;
    const int axis = (int)(t - P * 3);
    float bias[3] = {0.5f, 0.5f, 0.5f};
    bias[axis] = 0.0f;
    const Stencil s = particle_stencil(G, px, P, bias);
    int sh[3] = {G.n[0], G.n[1], G.n[2]};
    sh[axis] += 1;
    const float* gv = g.v[axis];
    double v = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;          // accumulated in the particle arrays' precision (fp64), like pv[P,axis] += ...
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
                const int gix = clampi(s.gi[0] + ix, G.n[0] - 1), giy = clampi(s.gi[1] + iy, G.n[1] - 1), giz = clampi(s.gi[2] + iz, G.n[2] - 1);
                const float wx = ix ? s.w[0] : 1.0f - s.w[0];
                const float wy = iy ? s.w[1] : 1.0f - s.w[1];
                const float wz = iz ? s.w[2] : 1.0f - s.w[2];
                const float val = gv[((long long)gix * sh[1] + giy) * sh[2] + giz];
                v += (double)(wx * wy * wz * val);
                c0 += (double)((float)(2 * ix - 1) * wy * wz * val) / G.cell[0];
                c1 += (double)(wx * (float)(2 * iy - 1) * wz * val) / G.cell[1];
                c2 += (double)(wx * wy * (float)(2 * iz - 1) * val) / G.cell[2];
            }
    pv[P * 3 + axis] = v;
    double* pca = axis == 0 ? cx : (axis == 1 ? cy : cz);
    pca[P * 3 + 0] = c0; pca[P * 3 + 1] = c1; pca[P * 3 + 2] = c2;
}

// ---------------------------------------------------------------------------------------------
// Fluid level set (:94-136): phi = min over particles within a 5^3 cell neighbourhood of |x_cell - x_p| - r.
// fp64 atomic min as ONE integer atomic: non-negative doubles order like signed integers, negative ones inversely like
// unsigned integers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    if (v >= 0.0) atomicMin(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
    else atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(128) grid_levelset_kernel(GridGeom G, long long np, const double* __restrict__ px, double r, double* __restrict__ phi) {
    // one thread per (particle, dx-plane of the 5^3 neighbourhood): 25 samples each
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np * 5) return;
    const long long P = t / 5;
    const int ddx = (int)(t - P * 5) - 2;
    float x[3];
    long long gi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        x[d] = (float)px[P * 3 + d];
        gi[d] = (long long)floor((double)(x[d] - G.bmin[d]) / G.cell[d]);
    }
    const int gx = clampi(gi[0] + ddx, G.n[0] - 1);
    const float px0 = (float)(((double)gx + 0.5) * G.cell[0] + (double)G.bmin[0] - (double)x[0]);
    for (int dy = -2; dy <= 2; ++dy) {
        const int gy = clampi(gi[1] + dy, G.n[1] - 1);
        const float py0 = (float)(((double)gy + 0.5) * G.cell[1] + (double)G.bmin[1] - (double)x[1]);
        for (int dz = -2; dz <= 2; ++dz) {
            const int gz = clampi(gi[2] + dz, G.n[2] - 1);
            const float pz0 = (float)(((double)gz + 0.5) * G.cell[2] + (double)G.bmin[2] - (double)x[2]);
            double n2 = 0.0;                      // norm(): n += v[d]*v[d] with fp32 products, fp64 sum
            n2 += (double)(px0 * px0);
            n2 += (double)(py0 * py0);
            n2 += (double)(pz0 * pz0);
            atomic_min_f64(phi + ((long long)gx * G.n[1] + gy) * G.n[2] + gz, sqrt(n2) - r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fluid volume splat on the (2n+1)^3 node grid (:224-252) and the clamp to the cell volume (:253-258)
// G describes the NODE grid here: n = resolution (2*gres+1), cell = bound_size / (2*gres), no bias.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_volume_kernel(GridGeom G, long long np, const double* __restrict__ px, double pvol, double* __restrict__ gvol) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= np) return;
    const float bias[3] = {0.0f, 0.0f, 0.0f};
    const Stencil s = particle_stencil(G, px, P, bias);
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int iz = 0; iz < 2; ++iz) {
                const int gix = clampi(s.gi[0] + ix, G.n[0] - 1), giy = clampi(s.gi[1] + iy, G.n[1] - 1), giz = clampi(s.gi[2] + iz, G.n[2] - 1);
                const float wx = ix ? s.w[0] : 1.0f - s.w[0];
                const float wy = iy ? s.w[1] : 1.0f - s.w[1];
                const float wz = iz ? s.w[2] : 1.0f - s.w[2];
                atomicAdd(gvol + ((long long)gix * G.n[1] + giy) * G.n[2] + giz, (double)(wx * wy * wz) * pvol);
            }
}

__global__ void __launch_bounds__(256) grid_fill_min_kernel(long long n, double* __restrict__ a, double fill, double cap, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) a[i] = fill;
    else a[i] = a[i] < cap ? a[i] : cap;      // min(gvol, cell_vol)
}

// ---------------------------------------------------------------------------------------------
// Extrapolation with mass validity (:501-557), IN PLACE: generation byte 0 = invalid, 1 = valid from the start (m > 0),
// k+1 = filled by sweep k; sweep k reads a neighbour only if 1 <= generation <= k, i.e. exactly the faces that were
// valid before the sweep — the reference's Jacobi sweeps without its twelve array copies per sweep.
// ---------------------------------------------------------------------------------------------
struct Ext3 { float* v[3]; const float* m[3]; uint8_t* gen[3]; };

__global__ void __launch_bounds__(256) grid_validity_kernel(long long n0, long long n1, long long n2, Ext3 e) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n[3] = {n0, n1, n2};
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (i < n[a]) e.gen[a][i] = e.m[a][i] > 0.0f ? 1 : 0;
}

__global__ void __launch_bounds__(256) grid_extrapolate_kernel(int nx, int ny, int nz, Ext3 e, int sweep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int sw = (unsigned int)sweep;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int s0 = nx + (a == 0), s1 = ny + (a == 1), s2 = nz + (a == 2);
        if (i >= (long long)s0 * s1 * s2) continue;
        const int z = (int)(i % s2), y = (int)((i / s2) % s1), x = (int)(i / ((long long)s1 * s2));
        if (x == 0 || x >= s0 - 1 || y == 0 || y >= s1 - 1 || z == 0 || z >= s2 - 1) continue;
        uint8_t* gen = e.gen[a];
        float* v = e.v[a];
        if (gen[i] != 0) continue;
        const long long sx = (long long)s1 * s2, sy = s2;
        auto ok = [&](long long j) { const unsigned int g = gen[j]; return g >= 1u && g <= sw; };
        double val = 0.0;                          // the reference accumulates in fp64 (val = 0.0) and stores fp32
        int count = 0;
        if (ok(i + sx)) { val += (double)v[i + sx]; ++count; }
        if (ok(i - sx)) { val += (double)v[i - sx]; ++count; }
        if (ok(i + sy)) { val += (double)v[i + sy]; ++count; }
        if (ok(i - sy)) { val += (double)v[i - sy]; ++count; }
        if (ok(i + 1)) { val += (double)v[i + 1]; ++count; }
        if (ok(i - 1)) { val += (double)v[i - 1]; ++count; }
        if (count > 0) {
            v[i] = (float)(val / count);
            gen[i] = (uint8_t)(sweep + 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Solid boundary condition (cell 5): dv = -(1 - d/dx) * [min(0, n.(v - v_solid))] n_axis / |n|^2 on interior faces closer than
// dx to the solid; 0 elsewhere (also on the array boundary, which is what the reference's in-bounds threads write).
// The two other velocity components are mass-weighted means of the four surrounding faces.
// ---------------------------------------------------------------------------------------------
struct Bc3 { const float* v[3]; const float* m[3]; float* dv[3]; };

template <int A>
__device__ __forceinline__ void bc_face(int nx, int ny, int nz, const Bc3& b, const double* __restrict__ sphi, const double* __restrict__ sv, double dx, long long i) {
    const int n[3] = {nx, ny, nz};
    int sh[3] = {nx, ny, nz};
    sh[A] += 1;
    if (i >= (long long)sh[0] * sh[1] * sh[2]) return;
    int c[3];
    c[2] = (int)(i % sh[2]); c[1] = (int)((i / sh[2]) % sh[1]); c[0] = (int)(i / ((long long)sh[1] * sh[2]));
    float* dv = b.dv[A];
    if (c[0] == 0 || c[0] >= sh[0] - 1 || c[1] == 0 || c[1] >= sh[1] - 1 || c[2] == 0 || c[2] >= sh[2] - 1) { dv[i] = 0.0f; return; }
    const long long fs2 = 1, fs1 = 2LL * nz + 1, fs0 = fs1 * (2LL * ny + 1);
    const long long fst[3] = {fs0, fs1, fs2};
    long long f = 0;                                  // fine node of the face: 2c + (1 - e_A)
#pragma unroll
    for (int d = 0; d < 3; ++d) f += (2LL * c[d] + (d == A ? 0 : 1)) * fst[d];
    const double ndist = sphi[f] / dx;
    if (ndist >= 1.0) { dv[i] = 0.0f; return; }
    // own component, and the mass-weighted mean of the four faces of each other component around this face
    double vel[3];
    vel[A] = (double)b.v[A][i];
#pragma unroll
    for (int B = 0; B < 3; ++B) {
        if (B == A) continue;
        int shb[3] = {n[0], n[1], n[2]};
        shb[B] += 1;
        double ms = 0.0, vs = 0.0;
#pragma unroll
        for (int ia = 0; ia < 2; ++ia)
#pragma unroll
            for (int ib = 0; ib < 2; ++ib) {
                int q[3] = {c[0], c[1], c[2]};
                q[A] -= ia;                           // the two cells either side of the face along its own axis
                q[B] += ib;                           // the two B-faces of each of those cells
                const long long j = ((long long)q[0] * shb[1] + q[1]) * shb[2] + q[2];
                const float mm = b.m[B][j];
                ms += (double)mm;
                vs += (double)(b.v[B][j] * mm);
            }
        vel[B] = vs / ms;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) vel[d] -= sv[f * 3 + d];
    double sn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) sn[d] = sphi[f + fst[d]] - sphi[f - fst[d]];
    const double sn_inv = 1.0 / (sn[0] * sn[0] + sn[1] * sn[1] + sn[2] * sn[2]);
    const double dot = sn[0] * vel[0] + sn[1] * vel[1] + sn[2] * vel[2];
    const double g = (dot < 0.0 ? dot : 0.0) * sn[A] * sn_inv;
    dv[i] = (float)(-g * (1.0 - ndist));
}

__global__ void __launch_bounds__(256) grid_boundary_kernel(int nx, int ny, int nz, Bc3 b, const double* __restrict__ sphi, const double* __restrict__ sv, double dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bc_face<0>(nx, ny, nz, b, sphi, sv, dx, i);
    bc_face<1>(nx, ny, nz, b, sphi, sv, dx, i);
    bc_face<2>(nx, ny, nz, b, sphi, sv, dx, i);
}

__global__ void __launch_bounds__(256) grid_add_dv_kernel(long long n0, long long n1, long long n2, float* v0, float* v1, float* v2,
                                                          const float* d0, const float* d1, const float* d2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0) v0[i] += d0[i];
    if (i < n1) v1[i] += d1[i];
    if (i < n2) v2[i] += d2[i];
}

static GridGeom make_geom(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3) {
    GridGeom G;
    G.n[0] = nx; G.n[1] = ny; G.n[2] = nz;
    for (int d = 0; d < 3; ++d) { G.bmin[d] = (float)bound_min3[d]; G.cell[d] = cell_size3[d]; }
    return G;
}

static inline unsigned int blocks_for(long long n, int threads) { return (unsigned int)((n + threads - 1) / threads); }

}  // namespace fs

using namespace fs;

extern "C" {

int fs_grid_p2g(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np,
                const double* px, const double* pm, const double* pv, const double* cx, const double* cy, const double* cz,
                float* mx, float* vx, float* my, float* vy, float* mz, float* vz, void* stream) {
    if (!bound_min3 || !cell_size3 || !px || !pm || !pv || !cx || !cy || !cz || !mx || !vx || !my || !vy || !mz || !vz) return fail(FS_ERR_ARG, "fs_grid_p2g: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || np < 0) return fail(FS_ERR_ARG, "fs_grid_p2g: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const GridGeom G = make_geom(nx, ny, nz, bound_min3, cell_size3);
    MacF32 g;
    g.m[0] = mx; g.m[1] = my; g.m[2] = mz; g.v[0] = vx; g.v[1] = vy; g.v[2] = vz;
    if (np > 0) {
        grid_p2g_kernel<<<blocks_for(np * 3, 256), 256, 0, s>>>(G, (long long)np, px, pm, pv, cx, cy, cz, g);
        FS_LAUNCH_CHECK();
    }
    const long long n0 = (long long)(nx + 1) * ny * nz, n1 = (long long)nx * (ny + 1) * nz, n2 = (long long)nx * ny * (nz + 1);
    const long long nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    grid_normalise_kernel<<<blocks_for(nmax, 256), 256, 0, s>>>(n0, n1, n2, g);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_grid_g2p(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np,
                const double* px, double* pv, double* cx, double* cy, double* cz, const float* vx, const float* vy, const float* vz, void* stream) {
    if (!bound_min3 || !cell_size3 || !px || !pv || !cx || !cy || !cz || !vx || !vy || !vz) return fail(FS_ERR_ARG, "fs_grid_g2p: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || np < 0) return fail(FS_ERR_ARG, "fs_grid_g2p: bad sizes");
    if (np == 0) return FS_OK;
    const GridGeom G = make_geom(nx, ny, nz, bound_min3, cell_size3);
    MacF32c g;
    g.v[0] = vx; g.v[1] = vy; g.v[2] = vz;
    grid_g2p_kernel<<<blocks_for(np * 3, 256), 256, 0, (cudaStream_t)stream>>>(G, (long long)np, px, pv, cx, cy, cz, g);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_grid_levelset(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np, const double* px,
                     double radius, double fill, double* phi, void* stream) {
    if (!bound_min3 || !cell_size3 || !px || !phi) return fail(FS_ERR_ARG, "fs_grid_levelset: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || np < 0) return fail(FS_ERR_ARG, "fs_grid_levelset: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = (long long)nx * ny * nz;
    grid_fill_min_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, phi, fill, 0.0, 0);       // ls.phi[:] = gdx * 3
    FS_LAUNCH_CHECK();
    if (np > 0) {
        grid_levelset_kernel<<<blocks_for(np * 5, 128), 128, 0, s>>>(make_geom(nx, ny, nz, bound_min3, cell_size3), (long long)np, px, radius, phi);
        FS_LAUNCH_CHECK();
    }
    return FS_OK;
}

int fs_grid_fluid_volume(int rx, int ry, int rz, const double* bound_min3, const double* cell_size3, int64_t np, const double* px,
                         double pvol, double cell_vol, double* vol, void* stream) {
    if (!bound_min3 || !cell_size3 || !px || !vol) return fail(FS_ERR_ARG, "fs_grid_fluid_volume: null argument");
    if (rx < 1 || ry < 1 || rz < 1 || np < 0) return fail(FS_ERR_ARG, "fs_grid_fluid_volume: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = (long long)rx * ry * rz;
    FS_CUDA(cudaMemsetAsync(vol, 0, (size_t)n * sizeof(double), s));                      // fv.vol[:] = 0.0
    if (np > 0) {
        grid_volume_kernel<<<blocks_for(np, 256), 256, 0, s>>>(make_geom(rx, ry, rz, bound_min3, cell_size3), (long long)np, px, pvol, vol);
        FS_LAUNCH_CHECK();
    }
    grid_fill_min_kernel<<<blocks_for(n, 256), 256, 0, s>>>(n, vol, 0.0, cell_vol, 1);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

size_t fs_grid_extrapolate_workspace_bytes(int nx, int ny, int nz) {
    if (nx < 1 || ny < 1 || nz < 1) return 0;
    return (size_t)(nx + 1) * ny * nz + (size_t)nx * (ny + 1) * nz + (size_t)nx * ny * (nz + 1) + 768;
}

int fs_grid_extrapolate(int nx, int ny, int nz, int sweeps, float* vx, float* vy, float* vz, const float* mx, const float* my, const float* mz,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!vx || !vy || !vz || !mx || !my || !mz || !workspace) return fail(FS_ERR_ARG, "fs_grid_extrapolate: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_grid_extrapolate: bad sizes");
    if (sweeps > 250) return fail(FS_ERR_ARG, "fs_grid_extrapolate: at most 250 sweeps");
    if (workspace_bytes < fs_grid_extrapolate_workspace_bytes(nx, ny, nz)) return fail(FS_ERR_ARG, "fs_grid_extrapolate: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n0 = (long long)(nx + 1) * ny * nz, n1 = (long long)nx * (ny + 1) * nz, n2 = (long long)nx * ny * (nz + 1);
    const long long nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    Ext3 e;
    e.v[0] = vx; e.v[1] = vy; e.v[2] = vz; e.m[0] = mx; e.m[1] = my; e.m[2] = mz;
    uint8_t* w = (uint8_t*)workspace;
    e.gen[0] = w; e.gen[1] = w + align_up((size_t)n0, 256); e.gen[2] = e.gen[1] + align_up((size_t)n1, 256);
    grid_validity_kernel<<<blocks_for(nmax, 256), 256, 0, s>>>(n0, n1, n2, e);
    FS_LAUNCH_CHECK();
    for (int k = 1; k <= sweeps; ++k) {
        grid_extrapolate_kernel<<<blocks_for(nmax, 256), 256, 0, s>>>(nx, ny, nz, e, k);
        FS_LAUNCH_CHECK();
    }
    return FS_OK;
}

int fs_grid_boundary(int nx, int ny, int nz, double dx, float* vx, float* vy, float* vz, const float* mx, const float* my, const float* mz,
                     const double* sphi, const double* sv, float* dvx, float* dvy, float* dvz, void* stream) {
    if (!vx || !vy || !vz || !mx || !my || !mz || !sphi || !sv || !dvx || !dvy || !dvz) return fail(FS_ERR_ARG, "fs_grid_boundary: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_grid_boundary: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n0 = (long long)(nx + 1) * ny * nz, n1 = (long long)nx * (ny + 1) * nz, n2 = (long long)nx * ny * (nz + 1);
    const long long nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    Bc3 b;
    b.v[0] = vx; b.v[1] = vy; b.v[2] = vz; b.m[0] = mx; b.m[1] = my; b.m[2] = mz; b.dv[0] = dvx; b.dv[1] = dvy; b.dv[2] = dvz;
    grid_boundary_kernel<<<blocks_for(nmax, 256), 256, 0, s>>>(nx, ny, nz, b, sphi, sv, dx);     // all dv from the OLD velocities ...
    FS_LAUNCH_CHECK();
    grid_add_dv_kernel<<<blocks_for(nmax, 256), 256, 0, s>>>(n0, n1, n2, vx, vy, vz, dvx, dvy, dvz);   // ... then v += dv
    FS_LAUNCH_CHECK();
    return FS_OK;
}

}  // extern "C"
