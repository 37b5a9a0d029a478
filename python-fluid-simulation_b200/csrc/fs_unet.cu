// UNet surrogate, the parts around the network (sm_100a): the 11-channel input tensor built straight from the MAC
// velocities / solid SDF / liquid volume, and the gather of the three velocity increments out of the network output.
//
// Replaces the array code of grad_v and unet_solve (3D_viscous_fluid_sim.ipynb:844-913): the notebook scatters the five
// inputs into fourteen zero-padded fp64 volumes, forms nine masked central differences with ~40 sliced CuPy operations,
// masks the SDF in place, concatenates, transposes and converts to fp32 — about 60 launches and ~3 GB of traffic for the
// 112x176x112 volume of its 48x80x48 grid.  Here ONE kernel evaluates the eleven channels of a padded voxel from the source
// arrays (each padded velocity sample is an index computation, never stored) and writes the fp32 NCDHW tensor once.
// Same arithmetic: differences in fp64 of fp32 samples, exact-zero tests on the samples, lvol / gdx^3 in fp64, one rounding.
#include "fs_common.cuh"

namespace fs {

struct UnetGeom {
    int n[3];        // grid cells
    int X, Y, Z;     // padded volume (data_size)
    int p[3];        // pad_l per axis: int((data_size - (2n+1)) / 2)
};

// padded staggered sample of component c at padded voxel (i,j,k): the notebook's v*_sympad, 0 where nothing was scattered
__device__ __forceinline__ double vpad(const UnetGeom& G, int c, const float* __restrict__ v, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= G.X || j >= G.Y || k >= G.Z) return 0.0;
    const int q[3] = {i - G.p[0], j - G.p[1], k - G.p[2]};
    int idx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int off = (d == c) ? 0 : 1;                 // component c sits on even fine indices along its own axis, odd along the others
        const int t = q[d] - off;
        if (t < 0 || (t & 1) || q[d] >= 2 * G.n[d] + 1) return 0.0;
        idx[d] = t >> 1;
    }
    const int s1 = G.n[1] + (c == 1), s2 = G.n[2] + (c == 2);
    return (double)v[((long long)idx[0] * s1 + idx[1]) * s2 + idx[2]];
}

__global__ void __launch_bounds__(256) unet_features_kernel(UnetGeom G, const float* __restrict__ vx, const float* __restrict__ vy, const float* __restrict__ vz,
                                                            const double* __restrict__ sphi, const double* __restrict__ lvol, double cell_vol /*gdx^3, as the caller computed it*/,
                                                            float pad_solid, float* __restrict__ out /*[11][X][Y][Z]*/) {
    const long long nvox = (long long)G.X * G.Y * G.Z;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nvox) return;
    const int k = (int)(t % G.Z), j = (int)((t / G.Z) % G.Y), i = (int)(t / ((long long)G.Y * G.Z));
    const float* v[3] = {vx, vy, vz};
    const int dim[3] = {G.X, G.Y, G.Z};
    const int pos[3] = {i, j, k};
    // g[c][ax] = v_c(p - e_ax) - v_c(p + e_ax), 0 on the first / last slice of the axis and where either sample is exactly 0
    double g[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            double r = 0.0;
            if (pos[ax] >= 1 && pos[ax] <= dim[ax] - 2) {
                const double a = vpad(G, c, v[c], i - (ax == 0), j - (ax == 1), k - (ax == 2));
                const double b = vpad(G, c, v[c], i + (ax == 0), j + (ax == 1), k + (ax == 2));
                if (a != 0.0 && b != 0.0) r = a - b;
            }
            g[c][ax] = r;
        }
    const int q[3] = {i - G.p[0], j - G.p[1], k - G.p[2]};
    const bool inside = q[0] >= 0 && q[1] >= 0 && q[2] >= 0 && q[0] <= 2 * G.n[0] && q[1] <= 2 * G.n[1] && q[2] <= 2 * G.n[2];
    float solid = pad_solid, lv = 0.0f;
    if (inside) {
        const long long f = ((long long)q[0] * (2 * G.n[1] + 1) + q[1]) * (2 * G.n[2] + 1) + q[2];
        solid = sphi[f] > 0.0 ? 0.0f : 1.0f;           // >0 -> 2 -> 0 ; <=0 -> 1   (the three in-place masking lines)
        lv = (float)(lvol[f] / cell_vol);
    }
    // dxdx dydy dzdz dxdy dxdz dydx dydz dzdx dzdy solid lvol
    out[0 * nvox + t] = (float)g[0][0];
    out[1 * nvox + t] = (float)g[1][1];
    out[2 * nvox + t] = (float)g[2][2];
    out[3 * nvox + t] = (float)g[0][1];
    out[4 * nvox + t] = (float)g[0][2];
    out[5 * nvox + t] = (float)g[1][0];
    out[6 * nvox + t] = (float)g[1][2];
    out[7 * nvox + t] = (float)g[2][0];
    out[8 * nvox + t] = (float)g[2][1];
    out[9 * nvox + t] = solid;
    out[10 * nvox + t] = lv;
}

// delv_c[x,y,z] = net[c][staggered position of face (x,y,z)] / steps_per_second
__global__ void __launch_bounds__(256) unet_gather_kernel(UnetGeom G, const float* __restrict__ net /*[3][X][Y][Z]*/, float inv_scale_is_div /*divisor*/,
                                                          float* __restrict__ dvx, float* __restrict__ dvy, float* __restrict__ dvz) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nvox = (long long)G.X * G.Y * G.Z;
    float* dst[3] = {dvx, dvy, dvz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s0 = G.n[0] + (c == 0), s1 = G.n[1] + (c == 1), s2 = G.n[2] + (c == 2);
        if (t >= (long long)s0 * s1 * s2) continue;
        const int z = (int)(t % s2), y = (int)((t / s2) % s1), x = (int)(t / ((long long)s1 * s2));
        const int i = G.p[0] + 2 * x + (c == 0 ? 0 : 1), j = G.p[1] + 2 * y + (c == 1 ? 0 : 1), k = G.p[2] + 2 * z + (c == 2 ? 0 : 1);
        dst[c][t] = net[c * nvox + ((long long)i * G.Y + j) * G.Z + k] / inv_scale_is_div;
    }
}

static int make_unet_geom(UnetGeom& G, int nx, int ny, int nz, int X, int Y, int Z) {
    G.n[0] = nx; G.n[1] = ny; G.n[2] = nz; G.X = X; G.Y = Y; G.Z = Z;
    const int dim[3] = {X, Y, Z};
    for (int d = 0; d < 3; ++d) {
        const int stg = 2 * G.n[d] + 1;
        if (dim[d] < stg) return fail(FS_ERR_ARG, "fs_unet: the padded volume is smaller than the (2n+1) fine grid");
        G.p[d] = (dim[d] - stg) / 2;                     // int((data_size - stg_size) / 2)
    }
    return FS_OK;
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_unet_features(int nx, int ny, int nz, int X, int Y, int Z, const float* vx, const float* vy, const float* vz,
                     const double* sphi, const double* lvol, double cell_vol, float pad_solid, float* out, void* stream) {
    if (!vx || !vy || !vz || !sphi || !lvol || !out) return fail(FS_ERR_ARG, "fs_unet_features: null argument");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FS_ERR_ARG, "fs_unet_features: bad grid");
    UnetGeom G;
    FS_TRY(make_unet_geom(G, nx, ny, nz, X, Y, Z));
    const long long nvox = (long long)X * Y * Z;
    if (!(cell_vol > 0.0)) return fail(FS_ERR_ARG, "fs_unet_features: cell_vol must be positive");
    unet_features_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, (cudaStream_t)stream>>>(G, vx, vy, vz, sphi, lvol, cell_vol, pad_solid, out);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_unet_gather(int nx, int ny, int nz, int X, int Y, int Z, const float* net_out, double divisor, float* dvx, float* dvy, float* dvz, void* stream) {
    if (!net_out || !dvx || !dvy || !dvz) return fail(FS_ERR_ARG, "fs_unet_gather: null argument");
    if (nx < 1 || ny < 1 || nz < 1 || !(divisor != 0.0)) return fail(FS_ERR_ARG, "fs_unet_gather: bad arguments");
    UnetGeom G;
    FS_TRY(make_unet_geom(G, nx, ny, nz, X, Y, Z));
    const long long n0 = (long long)(nx + 1) * ny * nz, n1 = (long long)nx * (ny + 1) * nz, n2 = (long long)nx * ny * (nz + 1);
    const long long nmax = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    unet_gather_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, (cudaStream_t)stream>>>(G, net_out, (float)divisor, dvx, dvy, dvz);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

}  // extern "C"
