// Row evaluator of the variational viscosity operator on the packed MAC lattice, generic in the
// dimension D (2 or 3) and in the component A of the row.
//
// Reference per-thread code: ViscosityCGSolver3D.py:248-456 (apply), :41-246 (RHS);
// ViscosityCGSolver2D.py:105-206, :6-102.  Here the 3x(1+14) hand-unrolled terms collapse to one
// rule, because every term of a row for component A has one of two shapes:
//   same-component neighbour along axis ax :  coef(ax) * vol_{hi|lo}(ax) * vel_A(i +/- e_ax),
//                                             coef = 2*scale*mu if ax == A else scale*mu
//   cross-component B (own axis b)         :  scale*mu * vol_hi(b) * [ +vel_B(i+e_b) - vel_B(i+e_b-e_A) ]
//                                             scale*mu * vol_lo(b) * [ -vel_B(i)     + vel_B(i-e_A)     ]
// with  vol_hi(A)=Vc(i), vol_lo(A)=Vc(i-e_A)  and, for ax != A,  vol_hi(ax)=E_{A,ax}(i+e_ax), vol_lo(ax)=E_{A,ax}(i)
// (E = edge-centred volume in 3-D, node volume in 2-D).  Term order and association follow the reference
// (+x,-x,+y,-y,+z,-z for the own component, then the other components in ascending order), so that the
// EXACT instantiation reproduces the reference's fp64 results bit for bit (no FMA contraction).
#pragma once
#include "fs_common.cuh"

namespace fs {

template <bool EXACT> struct Ar;
template <> struct Ar<true> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};
template <> struct Ar<false> {
    template <typename T> static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    template <typename T> static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    template <typename T> static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
};

// packed coefficient planes: [0..D-1] face volumes (NaN on rows that are never computed),
// [D] cell-centre volume, [D+a+b] edge/node volume between axes a<b.
template <int D> struct CoefCount { static constexpr int value = (D == 3) ? 7 : 4; };

enum RowMode { ROW_APPLY = 0, ROW_RHS = 1 };

// NB(comp, idx) returns the neighbour value to use (already masked as the mode requires).
template <typename T, int D, int A, bool EXACT, int MODE, class NB>
__device__ __forceinline__ T visc_row(const T* const* __restrict__ coef, long long i, const long long* st,
                                       T center, T own, T s, T s2, NB nb) {
    using R = Ar<EXACT>;
    T hi[D], lo[D];
#pragma unroll
    for (int ax = 0; ax < D; ++ax) {
        if (ax == A) {
            hi[ax] = __ldg(coef[D] + i);
            lo[ax] = __ldg(coef[D] + i - st[A]);
        } else {
            const T* E = coef[D + A + ax];
            hi[ax] = __ldg(E + i + st[ax]);
            lo[ax] = __ldg(E + i);
        }
    }
    T val;
    if (MODE == ROW_APPLY) {
        // diag = vol_center + scale*mu*(w*hi0 + w*lo0 + w*hi1 + ...)   left to right, w=2 on the own axis
        T sum = T(0);
#pragma unroll
        for (int ax = 0; ax < D; ++ax) {
            T h = (ax == A) ? R::mul(T(2), hi[ax]) : hi[ax];
            T l = (ax == A) ? R::mul(T(2), lo[ax]) : lo[ax];
            sum = (ax == 0) ? R::add(h, l) : R::add(R::add(sum, h), l);
        }
        T diag = R::add(center, R::mul(s, sum));
        val = R::mul(diag, own);
    } else {
        val = R::mul(own, center);
    }
    // `plus` terms are subtracted by the apply and added by the RHS; `minus` terms the other way round
    auto plus = [&](T c, T vol, T v) {
        T t = R::mul(R::mul(c, vol), v);
        val = (MODE == ROW_APPLY) ? R::sub(val, t) : R::add(val, t);
    };
    auto minus = [&](T c, T vol, T v) {
        T t = R::mul(R::mul(c, vol), v);
        val = (MODE == ROW_APPLY) ? R::add(val, t) : R::sub(val, t);
    };
#pragma unroll
    for (int ax = 0; ax < D; ++ax) {
        const T c = (ax == A) ? s2 : s;
        plus(c, hi[ax], nb(A, i + st[ax]));
        plus(c, lo[ax], nb(A, i - st[ax]));
    }
#pragma unroll
    for (int B = 0; B < D; ++B) {
        if (B == A) continue;
        plus(s, hi[B], nb(B, i + st[B]));
        minus(s, hi[B], nb(B, i + st[B] - st[A]));
        minus(s, lo[B], nb(B, i));
        plus(s, lo[B], nb(B, i - st[A]));
    }
    return val;
}

// CG-loop row on PRE-SCALED coefficients (visc3d_scale_kernel, once per solve): plane A (A < D) holds the row's diagonal
// diag_A = V_A + scale*mu*(2 hi_A + 2 lo_A + sum of the other axes' hi + lo), plane D holds 2*scale*mu*Vc and planes
// D+a+b hold scale*mu*E_ab, i.e. exactly the products (c*vol) the reference forms first in every term (`2*scale*mu*vol*v`),
// so a row is one multiply and 14 fused multiply-adds instead of ~30 fp64 operations — what takes the operator apply off
// the fp64 pipe (64 lanes per clock per SM on B200) and back onto the memory system.  Inside the loop d is zero on every
// row that is not computed, so no neighbour masks (SURVEY A-1).
// CF(plane, axis, sign) returns the coefficient of `plane` at i + sign*e_axis (sign = 0: at i itself); all three arguments
// are compile-time constants after unrolling, so an accessor may map them to a fixed slot (shared-memory resident copy).
// NBU(comp, p, m) returns component `comp` at i + e_p - e_m (p, m = axis numbers, -1 = no shift): the unit-vector form of
// the neighbour accessor, so that a caller whose data does not sit at a flat lattice index (a shared-memory tile with
// x-planes in separate ring buffers) can resolve every offset at compile time.
template <typename T, int D, int A, class CF, class NBU>
__device__ __forceinline__ T visc_row_scaled_u(T diag, T own, CF cf, NBU nbu) {
    T val = diag * own;
#pragma unroll
    for (int ax = 0; ax < D; ++ax) {
        T hi, lo;
        if (ax == A) {
            hi = cf(D, 0, 0);
            lo = cf(D, A, -1);
        } else {
            hi = cf(D + A + ax, ax, +1);
            lo = cf(D + A + ax, 0, 0);
        }
        val -= hi * nbu(A, ax, -1);
        val -= lo * nbu(A, -1, ax);
        if (ax != A) {                       // cross-component terms of component B = ax
            val -= hi * nbu(ax, ax, -1);
            val += hi * nbu(ax, ax, A);
            val += lo * nbu(ax, -1, -1);
            val -= lo * nbu(ax, -1, A);
        }
    }
    return val;
}

template <typename T, int D, int A, class CF, class NB>
__device__ __forceinline__ T visc_row_scaled_cf(long long i, const long long* st, T diag, T own, CF cf, NB nb) {
    auto nbu = [&](int comp, int p, int m) -> T { return nb(comp, i + (p >= 0 ? st[p] : 0) - (m >= 0 ? st[m] : 0)); };
    return visc_row_scaled_u<T, D, A>(diag, own, cf, nbu);
}

template <typename T, int D, int A, class NB>
__device__ __forceinline__ T visc_row_scaled(const T* const* __restrict__ cs, long long i, const long long* st, T diag, T own, NB nb) {
    auto cf = [&](int plane, int axis, int sign) -> T { return __ldg(cs[plane] + i + (long long)sign * st[axis]); };
    return visc_row_scaled_cf<T, D, A>(i, st, diag, own, cf, nb);
}

// Slot of coefficient (plane, axis, sign) in the 16-value per-point set of the 3-D scaled operator:
// 0..2 diagonals; 3: 2sVc(i), 4..6: 2sVc(i - e_x / e_y / e_z); 7: sExy(i), 8: sExy(i+e_y), 9: sExy(i+e_x);
// 10: sExz(i), 11: sExz(i+e_z), 12: sExz(i+e_x); 13: sEyz(i), 14: sEyz(i+e_z), 15: sEyz(i+e_y)
__host__ __device__ constexpr int visc3_cslot(int plane, int axis, int sign) {
    return plane == 3 ? (sign == 0 ? 3 : 4 + axis)
         : plane == 4 ? (sign == 0 ? 7 : (axis == 1 ? 8 : 9))
         : plane == 5 ? (sign == 0 ? 10 : (axis == 2 ? 11 : 12))
                      : (sign == 0 ? 13 : (axis == 2 ? 14 : 15));
}

}  // namespace fs
