/*
 * TEST INFRASTRUCTURE ONLY — fp64 C/OpenMP restatement of the reference's 3-D viscosity CG hot loop, used as
 * the CPU baseline ("port") of bench.py and as a faster checker than the NumPy oracle at 64^3..128^3.
 * Never linked into or called by the product library.
 *
 * Follows ViscosityCGSolver3D.py:248-456 (matvecmul_{x,y,z}_kernel) and :588-610 (the CG loop) of the reference,
 * on the reference's own data layout (dense C-order MAC arrays, (2n+1)^3 fine grids).  Pinned against the NumPy
 * oracle (itself pinned bit-exactly to the reference's kernels) by tests/test_c_port_cpu.py; compiled with
 * -ffp-contract=off so the apply is bit-identical to it.
 *
 * The three 15-term kernels are restated as one axis-generic rule on the fine grid.  With f the fine index of the
 * face (2c + p_A, p_A = (0,1,1),(1,0,1),(1,1,0)), s = scale*mu, M(.) = [sphi >= 0]:
 *   diag = vol[f] + s * sum_ax w_ax (vol[f+e_ax] + vol[f-e_ax]),  w_ax = 2 if ax == A else 1   (left-to-right)
 *   q    = diag*v_A[c]
 *          - sum_ax  (w_ax s) vol[f+e_ax] M(f+2e_ax) v_A[c+e_ax] - (w_ax s) vol[f-e_ax] M(f-2e_ax) v_A[c-e_ax]
 *          - sum_{B != A, b = axis of B}
 *                s vol[f+e_b] ( M(f+e_b+e_A) v_B[c+e_b] - M(f+e_b-e_A) v_B[c+e_b-e_A] )
 *              + s vol[f-e_b] ( -M(f-e_b+e_A) v_B[c]    + M(f-e_b-e_A) v_B[c-e_A]     )
 * evaluated term by term in the reference's order.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n[3];            /* cells */
    ptrdiff_t fs[3];     /* fine-grid strides */
    int sh[3][3];        /* MAC array shapes per component */
    ptrdiff_t cs[3][3];  /* MAC array strides per component */
} geom_t;

static void make_geom(geom_t* g, int nx, int ny, int nz) {
    g->n[0] = nx; g->n[1] = ny; g->n[2] = nz;
    g->fs[2] = 1; g->fs[1] = 2 * (ptrdiff_t)nz + 1; g->fs[0] = g->fs[1] * (2 * (ptrdiff_t)ny + 1);
    for (int c = 0; c < 3; ++c) {
        for (int k = 0; k < 3; ++k) g->sh[c][k] = g->n[k] + (k == c);
        g->cs[c][2] = 1; g->cs[c][1] = g->sh[c][2]; g->cs[c][0] = (ptrdiff_t)g->sh[c][1] * g->sh[c][2];
    }
}

static inline ptrdiff_t cidx(const geom_t* g, int comp, const int* c) {
    return c[0] * g->cs[comp][0] + c[1] * g->cs[comp][1] + c[2] * g->cs[comp][2];
}

/* one row of component A at coarse index c (interior, fluid) */
static inline double row_apply(const geom_t* g, int A, const int* c, double s, const double* const* v,
                               const double* sphi, const double* vol) {
    ptrdiff_t f = 0;
    for (int k = 0; k < 3; ++k) f += (2 * (ptrdiff_t)c[k] + (k == A ? 0 : 1)) * g->fs[k];
    double hi[3], lo[3];
    for (int ax = 0; ax < 3; ++ax) { hi[ax] = vol[f + g->fs[ax]]; lo[ax] = vol[f - g->fs[ax]]; }
    double sum = 0.0;
    for (int ax = 0; ax < 3; ++ax) {
        const double h = (ax == A) ? 2 * hi[ax] : hi[ax], l = (ax == A) ? 2 * lo[ax] : lo[ax];
        sum = (ax == 0) ? h + l : (sum + h) + l;
    }
    const double diag = vol[f] + s * sum;
    const double* va = v[A];
    const ptrdiff_t i = cidx(g, A, c);
    double val = diag * va[i];
    for (int ax = 0; ax < 3; ++ax) {
        const double cf = (ax == A) ? 2 * s : s;
        if (sphi[f + 2 * g->fs[ax]] >= 0) val -= cf * hi[ax] * va[i + g->cs[A][ax]];
        if (sphi[f - 2 * g->fs[ax]] >= 0) val -= cf * lo[ax] * va[i - g->cs[A][ax]];
    }
    for (int B = 0; B < 3; ++B) {
        if (B == A) continue;
        const double* vb = v[B];
        const ptrdiff_t j = cidx(g, B, c);     /* v_B[c] */
        const ptrdiff_t eb = g->cs[B][B], ea = g->cs[B][A];
        if (sphi[f + g->fs[B] + g->fs[A]] >= 0) val -= s * hi[B] * vb[j + eb];
        if (sphi[f + g->fs[B] - g->fs[A]] >= 0) val += s * hi[B] * vb[j + eb - ea];
        if (sphi[f - g->fs[B] + g->fs[A]] >= 0) val += s * lo[B] * vb[j];
        if (sphi[f - g->fs[B] - g->fs[A]] >= 0) val -= s * lo[B] * vb[j - ea];
    }
    return val;
}

void port_visc3d_matvecmul(int nx, int ny, int nz, double scale, double mu,
                           const double* vx, const double* vy, const double* vz,
                           double* ox, double* oy, double* oz, const double* sphi, const double* vol) {
    geom_t g;
    make_geom(&g, nx, ny, nz);
    const double s = scale * mu;
    const double* v[3] = {vx, vy, vz};
    double* o[3] = {ox, oy, oz};
    for (int A = 0; A < 3; ++A) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int x = 1; x <= g.sh[A][0] - 2; ++x)
            for (int y = 1; y <= g.sh[A][1] - 2; ++y)
                for (int z = 1; z <= g.sh[A][2] - 2; ++z) {
                    const int c[3] = {x, y, z};
                    ptrdiff_t f = 0;
                    for (int k = 0; k < 3; ++k) f += (2 * (ptrdiff_t)c[k] + (k == A ? 0 : 1)) * g.fs[k];
                    o[A][cidx(&g, A, c)] = (sphi[f] < 0) ? 0.0 : row_apply(&g, A, c, s, v, sphi, vol);
                }
    }
}


/* RHS row (ViscosityCGSolver3D.py:41-246): the off-diagonal terms of row_apply with the COMPLEMENTARY neighbour mask
 * (sphi < 0: the neighbour face is solid and its extrapolated value moves to the right-hand side) and the opposite sign;
 * the leading term is v_A[c]*vol[f]. */
static inline double row_rhs(const geom_t* g, int A, const int* c, double s, const double* const* v,
                             const double* sphi, const double* vol) {
    ptrdiff_t f = 0;
    for (int k = 0; k < 3; ++k) f += (2 * (ptrdiff_t)c[k] + (k == A ? 0 : 1)) * g->fs[k];
    double hi[3], lo[3];
    for (int ax = 0; ax < 3; ++ax) { hi[ax] = vol[f + g->fs[ax]]; lo[ax] = vol[f - g->fs[ax]]; }
    const double* va = v[A];
    const ptrdiff_t i = cidx(g, A, c);
    double val = va[i] * vol[f];
    for (int ax = 0; ax < 3; ++ax) {
        const double cf = (ax == A) ? 2 * s : s;
        if (sphi[f + 2 * g->fs[ax]] < 0) val += cf * hi[ax] * va[i + g->cs[A][ax]];
        if (sphi[f - 2 * g->fs[ax]] < 0) val += cf * lo[ax] * va[i - g->cs[A][ax]];
    }
    for (int B = 0; B < 3; ++B) {
        if (B == A) continue;
        const double* vb = v[B];
        const ptrdiff_t j = cidx(g, B, c);
        const ptrdiff_t eb = g->cs[B][B], ea = g->cs[B][A];
        if (sphi[f + g->fs[B] + g->fs[A]] < 0) val += s * hi[B] * vb[j + eb];
        if (sphi[f + g->fs[B] - g->fs[A]] < 0) val -= s * hi[B] * vb[j + eb - ea];
        if (sphi[f - g->fs[B] + g->fs[A]] < 0) val -= s * lo[B] * vb[j];
        if (sphi[f - g->fs[B] - g->fs[A]] < 0) val += s * lo[B] * vb[j - ea];
    }
    return val;
}

/* initialize_solver (:504-513): b on interior rows, 0 on solid rows, boundary layer untouched */
void port_visc3d_rhs(int nx, int ny, int nz, double scale, double mu,
                     const double* vx, const double* vy, const double* vz,
                     double* bx, double* by, double* bz, const double* sphi, const double* vol) {
    geom_t g;
    make_geom(&g, nx, ny, nz);
    const double s = scale * mu;
    const double* v[3] = {vx, vy, vz};
    double* o[3] = {bx, by, bz};
    for (int A = 0; A < 3; ++A) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int x = 1; x <= g.sh[A][0] - 2; ++x)
            for (int y = 1; y <= g.sh[A][1] - 2; ++y)
                for (int z = 1; z <= g.sh[A][2] - 2; ++z) {
                    const int c[3] = {x, y, z};
                    ptrdiff_t f = 0;
                    for (int k = 0; k < 3; ++k) f += (2 * (ptrdiff_t)c[k] + (k == A ? 0 : 1)) * g.fs[k];
                    o[A][cidx(&g, A, c)] = (sphi[f] < 0) ? 0.0 : row_rhs(&g, A, c, s, v, sphi, vol);
                }
    }
}

/* extrapolate (:472-502, kernel :8-39): num_iter Jacobi sweeps per component; validity = sphi >= 0 at the face's fine node.
 * tmp_v / tmp_m: caller scratch of the largest component's size (doubles / bytes x 2). */
void port_visc3d_extrapolate(int nx, int ny, int nz, int num_iter, double* vx, double* vy, double* vz, const double* sphi,
                             double* tmp_v, unsigned char* tmp_m) {
    geom_t g;
    make_geom(&g, nx, ny, nz);
    double* v[3] = {vx, vy, vz};
    for (int A = 0; A < 3; ++A) {
        const ptrdiff_t n = (ptrdiff_t)g.sh[A][0] * g.sh[A][1] * g.sh[A][2];
        unsigned char* valid = tmp_m;
        unsigned char* nvalid = tmp_m + n;
#pragma omp parallel for collapse(2) schedule(static)
        for (int x = 0; x < g.sh[A][0]; ++x)
            for (int y = 0; y < g.sh[A][1]; ++y)
                for (int z = 0; z < g.sh[A][2]; ++z) {
                    const int c[3] = {x, y, z};
                    ptrdiff_t f = 0;
                    for (int k = 0; k < 3; ++k) f += (2 * (ptrdiff_t)c[k] + (k == A ? 0 : 1)) * g.fs[k];
                    valid[cidx(&g, A, c)] = sphi[f] >= 0;
                }
        for (int it = 0; it < num_iter; ++it) {
#pragma omp parallel for schedule(static)
            for (ptrdiff_t i = 0; i < n; ++i) { tmp_v[i] = v[A][i]; nvalid[i] = valid[i]; }
#pragma omp parallel for collapse(2) schedule(static)
            for (int x = 1; x <= g.sh[A][0] - 2; ++x)
                for (int y = 1; y <= g.sh[A][1] - 2; ++y)
                    for (int z = 1; z <= g.sh[A][2] - 2; ++z) {
                        const int c[3] = {x, y, z};
                        const ptrdiff_t i = cidx(&g, A, c);
                        if (valid[i]) continue;
                        double val = 0.0;
                        int cnt = 0;
                        for (int ax = 0; ax < 3; ++ax) {          /* +x,-x,+y,-y,+z,-z (:19-36) */
                            const ptrdiff_t e = g.cs[A][ax];
                            if (valid[i + e]) { val += v[A][i + e]; ++cnt; }
                            if (valid[i - e]) { val += v[A][i - e]; ++cnt; }
                        }
                        if (cnt > 0) { tmp_v[i] = val / cnt; nvalid[i] = 1; }
                    }
#pragma omp parallel for schedule(static)
            for (ptrdiff_t i = 0; i < n; ++i) { v[A][i] = tmp_v[i]; valid[i] = nvalid[i]; }
        }
    }
}

/* Sum of a[i]*b[i] in the order of NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum:
 * below 8 elements sequential, up to 128 elements eight interleaved partial sums combined as a balanced tree, above that two
 * halves with the split rounded down to a multiple of 8) — what `cp.sum(d * q)` (ViscosityCGSolver3D.py:585, :590, :600)
 * evaluates to under the NumPy-backed cupy shim the golden fixtures were generated with.  The order depends on n only, never on
 * the thread count or on scheduling, so the CG trajectory of this port is reproducible run to run (an OpenMP `reduction(+)`
 * combines the threads' partial sums in arrival order: on the stiff 6x8x6 fixture that moved the iteration count between 95
 * and 98 from one run to the next). */
static double pw_sum(const double* a, const double* b, ptrdiff_t n) {
    if (n < 8) {
        double res = 0.0;
        for (ptrdiff_t i = 0; i < n; ++i) res += a[i] * b[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j] * b[j];
        ptrdiff_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j] * b[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i] * b[i];
        return res;
    }
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return pw_sum(a, b, n2) + pw_sum(a + n2, b + n2, n - n2);
}

/* the same tree, its upper levels as OpenMP tasks (the split points depend on n only) */
static double pw_sum_tasks(const double* a, const double* b, ptrdiff_t n) {
    if (n <= (ptrdiff_t)1 << 17) return pw_sum(a, b, n);
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    double s1 = 0.0, s2 = 0.0;
#pragma omp task shared(s1)
    s1 = pw_sum_tasks(a, b, n2);
    s2 = pw_sum_tasks(a + n2, b + n2, n - n2);
#pragma omp taskwait
    return s1 + s2;
}

static double dot3(const geom_t* g, double* const* a, double* const* b) {
    double tot = 0.0;
    for (int A = 0; A < 3; ++A) {
        const ptrdiff_t n = (ptrdiff_t)g->sh[A][0] * g->sh[A][1] * g->sh[A][2];
        double s = 0.0;
#pragma omp parallel
#pragma omp single
        s = pw_sum_tasks(a[A], b[A], n);
        tot += s;
    }
    return tot;
}

/* Runs up to max_iter CG iterations (ViscosityCGSolver3D.py:588-610) on caller-provided state x, r, d (q scratch),
 * delta_io = current r.r.  Returns the number of iterations executed; stops early when delta < tol2. */
int64_t port_visc3d_cg(int nx, int ny, int nz, double scale, double mu,
                       double* xx, double* xy, double* xz, double* rx, double* ry, double* rz,
                       double* dx, double* dy, double* dz, double* qx, double* qy, double* qz,
                       const double* sphi, const double* vol, double tol2, int64_t max_iter, double* delta_io) {
    geom_t g;
    make_geom(&g, nx, ny, nz);
    double* x[3] = {xx, xy, xz};
    double* r[3] = {rx, ry, rz};
    double* d[3] = {dx, dy, dz};
    double* q[3] = {qx, qy, qz};
    double delta = *delta_io;
    int64_t it = 0;
    while (it < max_iter) {
        port_visc3d_matvecmul(nx, ny, nz, scale, mu, dx, dy, dz, qx, qy, qz, sphi, vol);
        const double dq = dot3(&g, d, q);
        const double alpha = delta / dq;
        for (int A = 0; A < 3; ++A) {
            const ptrdiff_t n = (ptrdiff_t)g.sh[A][0] * g.sh[A][1] * g.sh[A][2];
#pragma omp parallel for schedule(static)
            for (ptrdiff_t i = 0; i < n; ++i) {
                x[A][i] += alpha * d[A][i];
                r[A][i] -= alpha * q[A][i];
            }
        }
        const double old = delta;
        delta = dot3(&g, r, r);
        ++it;
        if (delta < tol2) break;
        const double beta = delta / old;
        for (int A = 0; A < 3; ++A) {
            const ptrdiff_t n = (ptrdiff_t)g.sh[A][0] * g.sh[A][1] * g.sh[A][2];
#pragma omp parallel for schedule(static)
            for (ptrdiff_t i = 0; i < n; ++i) d[A][i] = r[A][i] + beta * d[A][i];
        }
    }
    *delta_io = delta;
    return it;
}

void port_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

int port_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
