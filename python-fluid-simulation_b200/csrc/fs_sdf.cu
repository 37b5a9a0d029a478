// Rigid-body signed distance field: evaluate (sd + solid velocity at arbitrary points) and project (push particles out of
// the bodies) — sm_100a.  Replaces solver/sdf3D.py:218-279 (evaluate_kernel, project_kernel) and the per-shape device
// functions it calls (:12-215).  Body table layout (generate_rb, :294-327): rb_d[n][10][4] fp64 — row 0 = [2*shape + flipped,
// p1, p2, p3] (shape 0 sphere: radius; 1 box: full sizes; 2 cylinder: radius, height), rows 1-4 translation matrix,
// rows 5-8 rotation matrix, row 9 velocity.
//
// The arithmetic follows the reference's association term by term with non-contracted fp64 operations, so results match the
// reference (run under Numba's simulator) to the last bit wherever it uses + - * /; its `x ** 0.5` becomes sqrt.
// Reference quirks kept: box_project tests `rb[0,0] % 2 and ~(in_out)` — `~` of an int is never 0, so a FLIPPED box always
// clamps the point into the box and transforms back (also for points already inside); min_sd starts at 100.
// One deliberate repair: cylinder_eval reads y_clip before assignment for points between the end planes (UnboundLocalError in
// Python, an uninitialised value on a GPU); here y_clip = y there, as cylinder_project does.
#include "fs_common.cuh"

namespace fs {

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }

struct Body {
    const double* rb;                 // [10][4]
    __device__ __forceinline__ double at(int r, int c) const { return __ldg(rb + r * 4 + c); }
    __device__ __forceinline__ int shape() const { return (int)floor(at(0, 0) / 2.0); }
    __device__ __forceinline__ bool flipped() const { return fmod(at(0, 0), 2.0) != 0.0; }
    // pos_rb = inv(T R) p     (inv_rigid :32-41 + matvecmul4 :20-29)
    __device__ __forceinline__ void to_body(const double* p, double* q) const {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double t = 0.0;                                   // inv_tr[i,3]: tmp -= R[j,i] * T[j,3]
#pragma unroll
            for (int j = 0; j < 3; ++j) t = ds(t, dm(at(5 + j, i), at(1 + j, 3)));
            double acc = 0.0;                                 // tmp += inv_tr[i,j] * p[j] ; tmp += inv_tr[i,3]
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = da(acc, dm(at(5 + j, i), p[j]));
            q[i] = da(acc, t);
        }
    }
    // p = (T R) q             (mat_TR :12-17 + matvecmul4)
    __device__ __forceinline__ void to_world(const double* q, double* p) const {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) acc = da(acc, dm(at(5 + i, j), q[j]));
            p[i] = da(acc, at(1 + i, 3));
        }
    }
};

__device__ __forceinline__ double sq3(double a, double b, double c) { return da(da(dm(a, a), dm(b, b)), dm(c, c)); }

__device__ double sphere_sd(const Body& B, const double* p) {                      // :52-66
    const double d0 = ds(p[0], B.at(1, 3)), d1 = ds(p[1], B.at(2, 3)), d2 = ds(p[2], B.at(3, 3));
    double sd = ds(sqrt(sq3(d0, d1, d2)), B.at(0, 1));
    return B.flipped() ? -sd : sd;
}

__device__ double box_sd(const Body& B, const double* p) {                         // :86-109
    double q[3];
    B.to_body(p, q);
    double tmp = 0.0, max_disp = -100.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d = ds(fabs(q[i]), B.at(0, 1 + i) / 2.0);
        if (d > 0.0) tmp = da(tmp, dm(d, d));
        if (max_disp < d) max_disp = d;
    }
    double sd = sqrt(tmp);
    if (max_disp < 0.0) sd = da(sd, max_disp);
    return B.flipped() ? -sd : sd;
}

__device__ double cylinder_sd(const Body& B, const double* p) {                    // :154-178
    double q[3];
    B.to_body(p, q);
    const double hh = B.at(0, 2) / 2.0;
    double y_clip = q[1];
    if (q[1] < -hh) y_clip = -hh;
    else if (q[1] > hh) y_clip = hh;
    double sd = ds(sqrt(da(dm(q[0], q[0]), dm(q[2], q[2]))), B.at(0, 1));
    const bool capped = (y_clip == hh || y_clip == -hh);
    if (sd < 0.0) {
        if (capped) sd = fabs(ds(y_clip, q[1]));
        else sd = fmax(sd, fmax(ds(q[1], hh), -da(q[1], hh)));
    } else if (capped) {
        const double dy = fabs(ds(y_clip, q[1]));
        sd = sqrt(da(dm(sd, sd), dm(dy, dy)));
    }
    return B.flipped() ? -sd : sd;
}

__global__ void __launch_bounds__(256) sdf3d_evaluate_kernel(const double* __restrict__ rb_d, int nb, long long np, const double* __restrict__ pos,
                                                             double* __restrict__ sd_out, double* __restrict__ vel) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= np) return;
    const double p[3] = {pos[P * 3], pos[P * 3 + 1], pos[P * 3 + 2]};
    double min_sd = 100.0;
    int idx = 0;
    for (int i = 0; i < nb; ++i) {
        Body B{rb_d + (size_t)i * 40};
        const int sh = B.shape();
        double d;
        if (sh == 0) d = sphere_sd(B, p);
        else if (sh == 1) d = box_sd(B, p);
        else if (sh == 2) d = cylinder_sd(B, p);
        else continue;
        if (d < min_sd) { min_sd = d; idx = i; }
    }
    sd_out[P] = min_sd;
    const bool in = min_sd <= 0.0;                              // vel *= 0 first (:258), then the nearest body's velocity inside it
#pragma unroll
    for (int k = 0; k < 3; ++k) vel[P * 3 + k] = in ? __ldg(rb_d + (size_t)idx * 40 + 36 + k) : 0.0;
}

__device__ void sphere_project(const Body& B, double* p) {                         // :68-84
    const double d0 = ds(p[0], B.at(1, 3)), d1 = ds(p[1], B.at(2, 3)), d2 = ds(p[2], B.at(3, 3));
    const double dist = sqrt(sq3(d0, d1, d2));
    double sd = ds(dist, B.at(0, 1));
    if (B.flipped()) sd = -sd;
    if (sd < 0.0) {
        const double n[3] = {d0 / dist, d1 / dist, d2 / dist};
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = da(dm(n[i], B.at(0, 1)), B.at(1 + i, 3));
    }
}

__device__ void box_project(const Body& B, double* p) {                            // :111-152
    double q[3];
    B.to_body(p, q);
    int in_out = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double h = B.at(0, 1 + i) / 2.0;
        if (q[i] > h || q[i] < -h) ++in_out;
    }
    if (B.flipped()) {                     // `rb[0,0] % 2 and ~(in_out)`: ~int is never 0 -> every point is clamped into the box
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double h = B.at(0, 1 + i) / 2.0;
            if (q[i] < -h) q[i] = -h;
            else if (q[i] > h) q[i] = h;
        }
        B.to_world(q, p);
    } else if (in_out == 0) {              // inside a solid box: out through the nearest face
        int index = 0;
        double dist = 100.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double h = B.at(0, 1 + i) / 2.0;
            if (ds(h, q[i]) < dist) { dist = ds(h, q[i]); index = i * 2; }
            if (da(q[i], h) < dist) { dist = da(q[i], h); index = i * 2 + 1; }
        }
        const double step = (index % 2) ? -dist : dist;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i == index / 2) q[i] = da(q[i], step);
        B.to_world(q, p);
    }
}

__device__ void cylinder_project(const Body& B, double* p) {                       // :180-226
    double q[3];
    B.to_body(p, q);
    const double hh = B.at(0, 2) / 2.0, rad = B.at(0, 1);
    double y_clip = q[1];
    if (q[1] < -hh) y_clip = -hh;
    else if (q[1] > hh) y_clip = hh;
    const double dist = sqrt(da(dm(q[0], q[0]), dm(q[2], q[2])));
    const double sd = ds(dist, rad);
    if (B.flipped()) {
        if (fabs(y_clip) == hh || sd > 0.0) {
            if (sd < 0.0) q[1] = y_clip;
            else { q[0] = dm(q[0] / dist, rad); q[2] = dm(q[2] / dist, rad); q[1] = y_clip; }
        }
        B.to_world(q, p);
    } else if (sd < 0.0 && fabs(y_clip) != hh) {
        const double a = ds(q[1], hh), b = -da(q[1], hh);
        const double mx = fmax(sd, fmax(a, b));
        if (mx == sd) { q[0] = dm(q[0] / dist, rad); q[2] = dm(q[2] / dist, rad); }
        else if (mx == a) q[1] = hh;
        else q[1] = -hh;
        B.to_world(q, p);
    }
}

__global__ void __launch_bounds__(256) sdf3d_project_kernel(const double* __restrict__ rb_d, int nb, long long np, double* __restrict__ pos) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= np) return;
    double p[3] = {pos[P * 3], pos[P * 3 + 1], pos[P * 3 + 2]};
    for (int i = 0; i < nb; ++i) {           // bodies in table order, each acting on the result of the previous one
        Body B{rb_d + (size_t)i * 40};
        const int sh = B.shape();
        if (sh == 0) sphere_project(B, p);
        else if (sh == 1) box_project(B, p);
        else if (sh == 2) cylinder_project(B, p);
    }
    pos[P * 3] = p[0]; pos[P * 3 + 1] = p[1]; pos[P * 3 + 2] = p[2];
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_sdf3d_evaluate(const double* rb_d, int nbodies, int64_t npos, const double* pos, double* sd, double* vel, void* stream) {
    if (!pos || !sd || !vel || (nbodies > 0 && !rb_d)) return fail(FS_ERR_ARG, "fs_sdf3d_evaluate: null argument");
    if (nbodies < 0 || npos < 0) return fail(FS_ERR_ARG, "fs_sdf3d_evaluate: bad sizes");
    if (npos == 0) return FS_OK;
    sdf3d_evaluate_kernel<<<(unsigned)((npos + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rb_d, nbodies, (long long)npos, pos, sd, vel);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_sdf3d_project(const double* rb_d, int nbodies, int64_t npos, double* pos, void* stream) {
    if (!pos || (nbodies > 0 && !rb_d)) return fail(FS_ERR_ARG, "fs_sdf3d_project: null argument");
    if (nbodies < 0 || npos < 0) return fail(FS_ERR_ARG, "fs_sdf3d_project: bad sizes");
    if (npos == 0 || nbodies == 0) return FS_OK;
    sdf3d_project_kernel<<<(unsigned)((npos + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rb_d, nbodies, (long long)npos, pos);
    FS_LAUNCH_CHECK();
    return FS_OK;
}

}  // extern "C"
