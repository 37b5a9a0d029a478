// 2-D variational viscosity solver on a packed MAC lattice (sm_100a).
//
// Replaces ViscosityCGSolver2D.py (kernels :6-219, launchers :222-244, solve :275-318).
// Same design as fs_visc3d.cu with D=2: lattice X=W+1, Yp=roundup(H+1,4), flat index i=x*Yp+y;
//   coef[0..1] Vu,Vv (NaN on rows never computed), coef[2] Vc, coef[3] Nn (node volume),
//   mask[0..1] fluid flags — NOTE the 2-D reference treats sphi <= 0 as solid (fluid iff sphi > 0,
//   ViscosityCGSolver2D.py:13,28,112,129) whereas 3-D uses sphi < 0 solid.  No extrapolation step in 2-D.
#include <type_traits>

#include "fs_common.cuh"
#include "fs_visc_rows.cuh"

namespace fs {

struct Lat2 {
    int W, H;
    int X, Yp;
    long long NL;
};

template <typename T> struct Visc2Dev {
    Lat2 L;
    const T* coef[4];
    const uint8_t* mask[2];
};

constexpr int kT2 = 256;

__device__ __forceinline__ void comp_shape2(const Lat2& L, int c, int& s0, int& s1) {
    s0 = L.W + (c == 0);
    s1 = L.H + (c == 1);
}

template <typename T>
__global__ void __launch_bounds__(kT2) visc2d_pack_kernel(Lat2 L, const double* __restrict__ sphi, const double* __restrict__ lvol, double vol_norm,
                                                          T* __restrict__ coef, uint8_t* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    const int y = (int)(i % L.Yp), x = (int)(i / L.Yp);
    const long long fy = 1, fx = 2LL * L.H + 1;
    const long long f0 = 2LL * x * fx + 2LL * y * fy;
    const bool ix = x < L.W, iy = y < L.H, iny = y <= L.H;
    const T nan = (T)__longlong_as_double(0x7ff8000000000000LL);
    auto vol = [&](long long off) { return (T)(lvol[f0 + off] / vol_norm); };
    {   // u face (2x, 2y+1)
        bool fluid = false;
        T v = nan;
        if (iy) {
            fluid = sphi[f0 + fy] > 0.0;
            const bool interior = x >= 1 && x <= L.W - 1 && y >= 1 && y <= L.H - 2;
            if (fluid && interior) v = vol(fy);
        }
        coef[0 * L.NL + i] = v;
        mask[0 * L.NL + i] = fluid;
    }
    {   // v face (2x+1, 2y)
        bool fluid = false;
        T v = nan;
        if (ix && iny) {
            fluid = sphi[f0 + fx] > 0.0;
            const bool interior = x >= 1 && x <= L.W - 2 && y >= 1 && y <= L.H - 1;
            if (fluid && interior) v = vol(fx);
        }
        coef[1 * L.NL + i] = v;
        mask[1 * L.NL + i] = fluid;
    }
    coef[2 * L.NL + i] = (ix && iy) ? vol(fx + fy) : T(0);   // cell centre
    coef[3 * L.NL + i] = (iny) ? vol(0) : T(0);              // node
}

template <typename T, typename S>
__global__ void __launch_bounds__(kT2) visc2d_load_kernel(Lat2 L, const S* __restrict__ a0, const S* __restrict__ a1, T* __restrict__ vec) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    const int y = (int)(i % L.Yp), x = (int)(i / L.Yp);
    const S* src[2] = {a0, a1};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int s0, s1;
        comp_shape2(L, c, s0, s1);
        T v = T(0);
        if (x < s0 && y < s1) v = (T)src[c][(long long)x * s1 + y];
        vec[c * L.NL + i] = v;
    }
}

template <typename T, typename S>
__global__ void __launch_bounds__(kT2) visc2d_store_kernel(Lat2 L, const T* __restrict__ vec, const uint8_t* __restrict__ mask,
                                                           S* __restrict__ a0, S* __restrict__ a1, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    const int y = (int)(i % L.Yp), x = (int)(i / L.Yp);
    S* dst[2] = {a0, a1};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int s0, s1;
        comp_shape2(L, c, s0, s1);
        if (!(x < s0 && y < s1)) continue;
        bool w;
        if (mode == FS_STORE_ALL) w = true;
        else if (mode == FS_STORE_INTERIOR) w = x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2;
        else w = x >= 1 && x <= L.W - 1 && y >= 1 && y <= L.H - 1 && mask[c * L.NL + i];   // :209-219
        if (w) dst[c][(long long)x * s1 + y] = (S)vec[c * L.NL + i];
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kT2) visc2d_general_kernel(Visc2Dev<T> P, T s, T s2, const T* __restrict__ src, T* __restrict__ dst) {
    const Lat2& L = P.L;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NL) return;
    const int y = (int)(i % L.Yp), x = (int)(i / L.Yp);
    const long long st[2] = {L.Yp, 1};
    const long long NL = L.NL;
    const uint8_t* const* mask = P.mask;
    auto nb = [&](int comp, long long j) -> T {
        const bool fluid = mask[comp][j] != 0;
        const bool keep = (MODE == ROW_APPLY) ? fluid : !fluid;
        return keep ? src[comp * NL + j] : T(0);
    };
    auto row = [&](auto Atag) {
        constexpr int A = decltype(Atag)::value;
        int s0, s1;
        comp_shape2(L, A, s0, s1);
        if (!(x >= 1 && x <= s0 - 2 && y >= 1 && y <= s1 - 2)) return;
        T out = T(0);
        if (mask[A][i]) out = visc_row<T, 2, A, true, MODE>(P.coef, i, st, P.coef[A][i], src[A * NL + i], s, s2, nb);
        dst[A * NL + i] = out;
    };
    row(std::integral_constant<int, 0>{});
    row(std::integral_constant<int, 1>{});
}

// Persistent grid-stride version (same reasoning as the 3-D K1): one block reduction at the end instead of one per 256 points.
constexpr int kV2BlocksPerSM = 4;

template <typename T>
__global__ void __launch_bounds__(kT2) visc2d_apply_dot_kernel(Visc2Dev<T> P, T s, T s2, const T* __restrict__ d, T* __restrict__ q,
                                                               CgState* st_, double* partials) {
    if (*(volatile int*)&st_->done) return;
    const Lat2& L = P.L;
    const long long NL = L.NL;
    const long long st[2] = {L.Yp, 1};
    double acc = 0.0;
    auto nb = [&](int comp, long long j) -> T { return __ldg(d + comp * NL + j); };
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NL; i += stride) {
        const T cu = __ldg(P.coef[0] + i), cv = __ldg(P.coef[1] + i);
        T ou = T(0), ov = T(0);
        if (cu == cu) {          // computed rows are interior, so every neighbour index is inside the lattice
            const T own = __ldg(d + i);
            ou = visc_row<T, 2, 0, false, ROW_APPLY>(P.coef, i, st, cu, own, s, s2, nb);
            acc += (double)own * (double)ou;
        }
        if (cv == cv) {
            const T own = __ldg(d + NL + i);
            ov = visc_row<T, 2, 1, false, ROW_APPLY>(P.coef, i, st, cv, own, s, s2, nb);
            acc += (double)own * (double)ov;
        }
        q[i] = ou;
        q[NL + i] = ov;
    }
    grid_sum_finish(acc, partials, &st_->counter[0], [=](double sum) { st_->dq = sum; });
}

}  // namespace fs

using namespace fs;

struct fs_visc2d {
    Lat2 L;
    int dtype;
    size_t esz;
    char* ws;
    char* coef;     // [4][NL]
    char* vecs;     // [5][2][NL]
    uint8_t* mask;  // [2][NL]
    double* partials;
    CgState* st;
    CgHost cg;
    int grid_pts;
    bool packed;
    IterGraph graph;
};

static Lat2 make_lat2(int W, int H) {
    Lat2 L;
    L.W = W; L.H = H; L.X = W + 1; L.Yp = (H + 1 + 3) / 4 * 4;
    L.NL = (long long)L.X * L.Yp;
    return L;
}

struct V2Layout { size_t coef, vecs, mask, partials, st, total; int grid_pts; };

static V2Layout v2_layout(const Lat2& L, size_t esz) {
    V2Layout o;
    size_t p = 0;
    o.grid_pts = (int)((L.NL + kT2 - 1) / kT2);
    o.coef = p; p = align_up(p + 4 * L.NL * esz, 256);
    o.vecs = p; p = align_up(p + 10 * L.NL * esz, 256);
    o.mask = p; p = align_up(p + 2 * L.NL, 256);
    size_t np = (size_t)(o.grid_pts > kVecGrid ? o.grid_pts : kVecGrid);
    o.partials = p; p = align_up(p + np * sizeof(double), 256);
    o.st = p; p = align_up(p + sizeof(CgState), 256);
    o.total = p;
    return o;
}

template <typename T> static Visc2Dev<T> dev_view2(const fs_visc2d* h) {
    Visc2Dev<T> P;
    P.L = h->L;
    for (int k = 0; k < 4; ++k) P.coef[k] = reinterpret_cast<const T*>(h->coef) + k * h->L.NL;
    for (int k = 0; k < 2; ++k) P.mask[k] = h->mask + k * h->L.NL;
    return P;
}

template <typename T> static T* vec2(const fs_visc2d* h, int v) { return reinterpret_cast<T*>(h->vecs) + (long long)v * 2 * h->L.NL; }

#define FS_DISPATCH2(h, ...)                                     \
    do {                                                         \
        if ((h)->dtype == FS_F32) { using T = float; __VA_ARGS__; } \
        else { using T = double; __VA_ARGS__; }                  \
    } while (0)

static int v2_general(fs_visc2d* h, double scale, double mu, int src, int dst, int mode, cudaStream_t s) {
    if (!h->packed) return fail(FS_ERR_STATE, "viscosity operator used before fs_visc2d_pack");
    if (src < 0 || src >= FS_NUM_VECS || dst < 0 || dst >= FS_NUM_VECS || src == dst) return fail(FS_ERR_ARG, "bad src/dst vector ids");
    const double sm = scale * mu;
    if (mode == ROW_APPLY) {
        FS_DISPATCH2(h, visc2d_general_kernel<T, ROW_APPLY><<<h->grid_pts, kT2, 0, s>>>(dev_view2<T>(h), (T)sm, (T)(2 * sm), vec2<T>(h, src), vec2<T>(h, dst)));
    } else {
        FS_DISPATCH2(h, visc2d_general_kernel<T, ROW_RHS><<<h->grid_pts, kT2, 0, s>>>(dev_view2<T>(h), (T)sm, (T)(2 * sm), vec2<T>(h, src), vec2<T>(h, dst)));
    }
    FS_LAUNCH_CHECK();
    return FS_OK;
}

static int v2_iteration(fs_visc2d* h, double sm, cudaStream_t s) {
    const long long n = 2 * h->L.NL;
    const int grid = h->grid_pts < kSMs * kV2BlocksPerSM ? h->grid_pts : kSMs * kV2BlocksPerSM;
    FS_DISPATCH2(h, visc2d_apply_dot_kernel<T><<<grid, kT2, 0, s>>>(dev_view2<T>(h), (T)sm, (T)(2 * sm), vec2<T>(h, FS_VEC_D), vec2<T>(h, FS_VEC_Q), h->st, h->partials));
    FS_LAUNCH_CHECK();
    FS_DISPATCH2(h, FS_TRY(cg_launch_update_xr<T>(n, vec2<T>(h, FS_VEC_X), vec2<T>(h, FS_VEC_R), vec2<T>(h, FS_VEC_D), vec2<T>(h, FS_VEC_Q), h->st, h->partials, s)));
    FS_DISPATCH2(h, FS_TRY(cg_launch_update_d<T>(n, vec2<T>(h, FS_VEC_D), vec2<T>(h, FS_VEC_R), h->st, s)));
    return FS_OK;
}

extern "C" {

size_t fs_visc2d_workspace_bytes(int W, int H, int dtype) {
    if (W < 1 || H < 1 || (dtype != FS_F32 && dtype != FS_F64)) return 0;
    return v2_layout(make_lat2(W, H), dtype == FS_F32 ? 4 : 8).total;
}

int fs_visc2d_create(fs_visc2d** out, int W, int H, int dtype, void* ws, size_t ws_bytes) {
    if (!out || !ws) return fail(FS_ERR_ARG, "fs_visc2d_create: null argument");
    if (W < 1 || H < 1) return fail(FS_ERR_ARG, "fs_visc2d_create: grid resolution must be >= 1");
    if (dtype != FS_F32 && dtype != FS_F64) return fail(FS_ERR_ARG, "fs_visc2d_create: dtype must be FS_F32 or FS_F64");
    if ((uintptr_t)ws % 256) return fail(FS_ERR_ARG, "fs_visc2d_create: workspace must be 256-byte aligned");
    fs_visc2d* h = new fs_visc2d();
    h->L = make_lat2(W, H);
    h->dtype = dtype;
    h->esz = dtype == FS_F32 ? 4 : 8;
    V2Layout lay = v2_layout(h->L, h->esz);
    if (ws_bytes < lay.total) { delete h; return fail(FS_ERR_ARG, "fs_visc2d_create: workspace too small"); }
    h->ws = (char*)ws;
    h->coef = h->ws + lay.coef; h->vecs = h->ws + lay.vecs; h->mask = (uint8_t*)(h->ws + lay.mask);
    h->partials = (double*)(h->ws + lay.partials); h->st = (CgState*)(h->ws + lay.st);
    h->grid_pts = lay.grid_pts;
    h->packed = false;
    int s = h->cg.init();
    if (s < 0) { delete h; return s; }
    h->cg.st_dev = h->st; h->cg.partials_dev = h->partials;
    cudaError_t e = cudaMemset(ws, 0, lay.total);
    if (e != cudaSuccess) { h->cg.destroy(); delete h; return fail(FS_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    *out = h;
    return FS_OK;
}

void fs_visc2d_destroy(fs_visc2d* h) {
    if (!h) return;
    h->graph.destroy();
    h->cg.destroy();
    delete h;
}

int fs_visc2d_lattice(const fs_visc2d* h, int* X, int* Yp, int64_t* NL) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (X) *X = h->L.X;
    if (Yp) *Yp = h->L.Yp;
    if (NL) *NL = h->L.NL;
    return FS_OK;
}

void* fs_visc2d_vector_ptr(const fs_visc2d* h, int vec, int comp) {
    if (!h || vec < 0 || vec >= FS_NUM_VECS || comp < 0 || comp > 1) return nullptr;
    return h->vecs + ((long long)vec * 2 + comp) * h->L.NL * h->esz;
}

int fs_visc2d_pack(fs_visc2d* h, const double* sphi, const double* lvol, double vol_norm, void* stream) {
    if (!h || !sphi || !lvol) return fail(FS_ERR_ARG, "fs_visc2d_pack: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    FS_DISPATCH2(h, visc2d_pack_kernel<T><<<h->grid_pts, kT2, 0, s>>>(h->L, sphi, lvol, vol_norm, reinterpret_cast<T*>(h->coef), h->mask));
    FS_LAUNCH_CHECK();
    h->packed = true;
    return FS_OK;
}

int fs_visc2d_load(fs_visc2d* h, int vec, const void* vx, const void* vy, int src_dtype, void* stream) {
    if (!h || !vx || !vy) return fail(FS_ERR_ARG, "fs_visc2d_load: null argument");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc2d_load: bad vector id");
    cudaStream_t s = (cudaStream_t)stream;
    if (src_dtype == FS_F32) {
        FS_DISPATCH2(h, visc2d_load_kernel<T, float><<<h->grid_pts, kT2, 0, s>>>(h->L, (const float*)vx, (const float*)vy, vec2<T>(h, vec)));
    } else if (src_dtype == FS_F64) {
        FS_DISPATCH2(h, visc2d_load_kernel<T, double><<<h->grid_pts, kT2, 0, s>>>(h->L, (const double*)vx, (const double*)vy, vec2<T>(h, vec)));
    } else return fail(FS_ERR_ARG, "fs_visc2d_load: bad dtype");
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_visc2d_store(fs_visc2d* h, int vec, void* vx, void* vy, int dst_dtype, int mode, void* stream) {
    if (!h || !vx || !vy) return fail(FS_ERR_ARG, "fs_visc2d_store: null argument");
    if (vec < 0 || vec >= FS_NUM_VECS) return fail(FS_ERR_ARG, "fs_visc2d_store: bad vector id");
    if (mode < FS_STORE_ALL || mode > FS_STORE_FLUID) return fail(FS_ERR_ARG, "fs_visc2d_store: bad mode");
    if (mode == FS_STORE_FLUID && !h->packed) return fail(FS_ERR_STATE, "fs_visc2d_store: FS_STORE_FLUID needs fs_visc2d_pack first");
    cudaStream_t s = (cudaStream_t)stream;
    if (dst_dtype == FS_F32) {
        FS_DISPATCH2(h, visc2d_store_kernel<T, float><<<h->grid_pts, kT2, 0, s>>>(h->L, vec2<T>(h, vec), h->mask, (float*)vx, (float*)vy, mode));
    } else if (dst_dtype == FS_F64) {
        FS_DISPATCH2(h, visc2d_store_kernel<T, double><<<h->grid_pts, kT2, 0, s>>>(h->L, vec2<T>(h, vec), h->mask, (double*)vx, (double*)vy, mode));
    } else return fail(FS_ERR_ARG, "fs_visc2d_store: bad dtype");
    FS_LAUNCH_CHECK();
    return FS_OK;
}

int fs_visc2d_rhs(fs_visc2d* h, double scale, double mu, int src_vec, int dst_vec, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    return v2_general(h, scale, mu, src_vec, dst_vec, ROW_RHS, (cudaStream_t)stream);
}

int fs_visc2d_apply(fs_visc2d* h, double scale, double mu, int src_vec, int dst_vec, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    return v2_general(h, scale, mu, src_vec, dst_vec, ROW_APPLY, (cudaStream_t)stream);
}

int fs_visc2d_cg(fs_visc2d* h, double scale, double mu, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    if (max_iter < 0) return fail(FS_ERR_ARG, "fs_visc2d_cg: max_iter < 0");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = 2 * h->L.NL;
    cg_state_init_kernel<<<1, 1, 0, s>>>(h->st, tol * tol, (long long)max_iter);
    FS_LAUNCH_CHECK();
    FS_TRY(v2_general(h, scale, mu, FS_VEC_X, FS_VEC_Q, ROW_APPLY, s));
    FS_DISPATCH2(h, FS_TRY(cg_launch_residual_init<T>(n, vec2<T>(h, FS_VEC_B), vec2<T>(h, FS_VEC_Q), vec2<T>(h, FS_VEC_D), vec2<T>(h, FS_VEC_R), h->st, h->partials, s)));
    const double sm = scale * mu;
    return cg_drive(h->cg, [&](cudaStream_t ss, long long nb) {
        return cg_enqueue_iterations(h->graph, true, sm, nb, [&](cudaStream_t s3) { return v2_iteration(h, sm, s3); }, ss);
    }, (long long)max_iter, stats, s);
}

int fs_visc2d_solve(fs_visc2d* h, double dt, double mu, double rho, double cell_vol, void* vx, void* vy, int vel_dtype,
                    const double* sphi, const double* lvol, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream) {
    if (!h) return fail(FS_ERR_ARG, "null handle");
    const double scale = dt / cell_vol / rho;                                 // ViscosityCGSolver2D.py:276
    FS_TRY(fs_visc2d_pack(h, sphi, lvol, cell_vol * 0.125, stream));          // :278 (yes, 0.125 in 2-D too)
    FS_TRY(fs_visc2d_load(h, FS_VEC_X, vx, vy, vel_dtype, stream));           // :279-280
    FS_TRY(fs_visc2d_rhs(h, scale, mu, FS_VEC_X, FS_VEC_B, stream));          // :282
    int status = fs_visc2d_cg(h, scale, mu, tol, max_iter, stats, stream);    // :283-314
    if (status != FS_OK) return status;
    FS_TRY(fs_visc2d_store(h, FS_VEC_X, vx, vy, vel_dtype, FS_STORE_FLUID, stream));   // :316
    FS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FS_OK;
}

}  // extern "C"
