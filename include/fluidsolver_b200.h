/*
 * fluidsolver_b200 — C ABI of the B200-native implicit-solver hot path
 * (matrix-free variational viscosity CG, pressure Poisson CG, solid fractions).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every pointer named *_dev* is a DEVICE pointer on the current CUDA device; `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  Arrays "in
 * reference layout" are dense C-order [x][y][z] (z contiguous) exactly as the reference
 * passes them (CuPy ndarrays): MAC faces u (nx+1,ny,nz), v (nx,ny+1,nz), w (nx,ny,nz+1),
 * fine grids (2nx+1,2ny+1,2nz+1), cells (nx,ny,nz).
 *
 * All functions return an int status: FS_OK (0), FS_NOT_CONVERGED (1) or a negative
 * error (argument / CUDA error; text via fs_last_error()).  They never throw.
 *
 * Reference interfaces replaced (paths relative to the reference's solver/ directory):
 *   fs_visc3d_*   ViscosityCGSolver3D.py:472-530 (extrapolate, initialize_solver, matvecmul,
 *                 apply_viscosity) and :532-613 (class ViscosityCGSolver3D, solve)
 *   fs_visc2d_*   ViscosityCGSolver2D.py:222-244, :246-318
 *   fs_press*     PressureCGSolver3D.py:155-171, :173-226; PressureCGSolver2D.py:122-138, :140-179
 *   fs_solidfrac* SolidFraction3D.py:28-32, SolidFraction2D.py:22-26 (+ SolidFractionCommon.py:4-60)
 *   CG vector algebra: the CuPy expressions inside solve() (ViscosityCGSolver3D.py:577-610,
 *                 PressureCGSolver3D.py:201-221), CGSolverBuffer.py:3-8 (caller-owned d,r,q,b)
 */
#ifndef FLUIDSOLVER_B200_H
#define FLUIDSOLVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_ABI_VERSION 1

/* status codes */
#define FS_OK 0
#define FS_NOT_CONVERGED 1
#define FS_ERR_ARG (-1)
#define FS_ERR_CUDA (-2)
#define FS_ERR_STATE (-3)

/* element types of caller arrays and of the solver's internal vectors */
#define FS_F32 0
#define FS_F64 1

/* solver vectors addressable through the API (internal lattice layout) */
#define FS_VEC_X 0 /* solution            (reference: x_x,x_y,x_z) */
#define FS_VEC_R 1 /* residual            (r_*) */
#define FS_VEC_D 2 /* search direction    (d_*) */
#define FS_VEC_Q 3 /* A*d                 (q_*) */
#define FS_VEC_B 4 /* right-hand side     (b_*) */
#define FS_NUM_VECS 5

/* which entries fs_visc*_store writes into the caller's arrays */
#define FS_STORE_ALL 0       /* every entry (debug / attribute views) */
#define FS_STORE_INTERIOR 1  /* rows a stencil kernel writes: 1..shape-2 on every axis of each component
                                (matvecmul / initialize_solver semantics, boundary layer untouched) */
#define FS_STORE_FLUID 2     /* apply_viscosity semantics: indices 1..g-1 on all axes, fluid faces only */

/* which rows the CG kernels visit (fs_visc3d_set_active_mode).  Both solve the same system (only exact zeros are dropped from the sums):
 *   FS_ACTIVE_FLUID    every row the reference's kernels compute (fluid face, interior; ViscosityCGSolver3D.py:251-258)
 *   FS_ACTIVE_NONZERO  (default) of those, only rows with at least one non-zero coefficient; an all-zero row is also an
 *                      all-zero column, its b, q, r, d stay exactly 0 and its x never changes (faces far from any liquid) */
#define FS_ACTIVE_FLUID 0
#define FS_ACTIVE_NONZERO 1

/* how the CG iterations are executed (fs_visc3d_set_cg_mode); identical arithmetic:
 *   FS_CG_KERNELS     the reference's recurrence (two reductions per iteration): three kernels per iteration (K1 apply+d.q,
 *                     K2 x/r update + r.r, K3 d update), replayed from a CUDA graph
 *   FS_CG_PERSISTENT  the same recurrence in one cooperative launch: the three kernels become phases separated by grid barriers
 *   FS_CG_KERNELS_SR / FS_CG_PERSISTENT_SR   single-reduction (Chronopoulos-Gear) form of the same iteration: w = A r fused
 *                     with both dot products (r.r, w.r), then one fused update of p, s, x, r — two kernels / two grid
 *                     barriers and ONE reduction per iteration; same iterates up to rounding, same stopping test
 *   FS_CG_AUTO        (default) single-reduction; persistent while the CG working set is small (launch-latency bound),
 *                     stand-alone kernels otherwise; the NCCL transport always uses FS_CG_KERNELS */
#define FS_CG_AUTO 0
#define FS_CG_KERNELS 1
#define FS_CG_PERSISTENT 2
#define FS_CG_KERNELS_SR 3
#define FS_CG_PERSISTENT_SR 4

/* result of a CG run */
typedef struct fs_cg_stats {
    int64_t iterations; /* CG iterations executed (reference: number of loop passes) */
    double delta;       /* final  sum r.r  (reference: self.delta) */
    double alpha;       /* last alpha      (self.alpha) */
    double beta;        /* last beta       (self.beta) */
    double delta0;      /* initial sum r.r */
    int32_t converged;  /* 1 if delta < tol^2 was reached */
    int32_t reserved;
} fs_cg_stats;

int fs_abi_version(void);
const char* fs_last_error(void);
/* number of kernels this library has launched in the calling process (bench bookkeeping) */
int64_t fs_launch_count(void);
/* Tuning switches of the kernels (the programmatic form of the FLUIDSOLVER_B200_* environment variables; they select between
 * result-equivalent implementations and never change what is computed).  value < 0 restores the default.
 *   "resident_form" : 0 = persistent CG through global memory, non-zero = shared-memory resident kernel
 *   "k1_block"      : n = trips of consecutive segments a CTA of the stand-alone K1s takes before it jumps ahead by the grid
 *                     (0 = interleaved); applies to HBM-sized active sets, 1000 + n to every size
 *   "k1_tile"       : 0 / 1 = off / on: dense lattices run the stand-alone K1s as a shared-memory tiled kernel that marches
 *                     along x (2 = on every lattice, tests)
 *   "sparse_setup"  : 0 / 1 = fs_visc3d_solve loads and extrapolates the velocities on the whole lattice / only around the
 *                     active set (everything the solve reads is identical; the lattice vector x keeps stale values elsewhere)
 * Returns FS_OK, or FS_ERR_ARG for an unknown name. */
int fs_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------
 * Viscosity, 3-D  (ViscosityCGSolver3D)
 *
 * The solver object works on an internal padded lattice (X=nx+1, Y=ny+1, Zp=roundup(nz+1,4)) that
 * lives in a caller-provided device workspace; it never allocates device memory itself.
 * ---------------------------------------------------------------------------------------- */
typedef struct fs_visc3d fs_visc3d;

size_t fs_visc3d_workspace_bytes(int nx, int ny, int nz, int dtype);
int fs_visc3d_create(fs_visc3d** out, int nx, int ny, int nz, int dtype, void* workspace_dev, size_t workspace_bytes);
void fs_visc3d_destroy(fs_visc3d* h);
/* lattice geometry: extents X,Y,Zp and elements per component plane-set NL = X*Y*Zp */
int fs_visc3d_lattice(const fs_visc3d* h, int* X, int* Y, int* Zp, int64_t* NL);
/* device pointer of component `comp` (0,1,2) of solver vector `vec` (FS_VEC_*) inside the workspace */
void* fs_visc3d_vector_ptr(const fs_visc3d* h, int vec, int comp);

/* Debug: copy `bytes` of an internal byte buffer to the host (what = 0: extrapolation validity / phase-timeline scratch,
 * 1: activity map). */
int fs_visc3d_debug_read(fs_visc3d* h, int what, void* out_host, size_t bytes);
/* Select how iterations are launched (FS_CG_*). */
int fs_visc3d_set_cg_mode(fs_visc3d* h, int mode);
/* After fs_visc3d_pack: the mode in use for the current active set (FS_CG_KERNELS, FS_CG_PERSISTENT, FS_CG_KERNELS_SR or FS_CG_PERSISTENT_SR). */
int fs_visc3d_cg_mode_in_use(fs_visc3d* h);
/* Select the active-row set (FS_ACTIVE_*); takes effect at the next fs_visc3d_pack. */
int fs_visc3d_set_active_mode(fs_visc3d* h, int mode);
/* After fs_visc3d_pack: number of active 32-point lattice segments the CG kernels walk, segments in the whole lattice,
 * and computed rows (any pointer may be NULL). */
int fs_visc3d_active_info(fs_visc3d* h, int64_t* segments, int64_t* segments_total, int64_t* rows, void* stream);

/* De-interleave the reference's fine-grid inputs into SoA coefficient planes + face masks.
 * vol = lvol / vol_norm (ViscosityCGSolver3D.py:568 uses vol_norm = cell_vol*0.125; pass 1.0 when
 * `lvol_dev` already holds the normalised `vol`).  Fluid test: sphi >= 0 (:255). */
int fs_visc3d_pack(fs_visc3d* h, const double* sphi_dev, const double* lvol_dev, double vol_norm, void* stream);
/* caller MAC arrays (reference layout, FS_F32/FS_F64) -> solver vector; padding is zeroed */
int fs_visc3d_load(fs_visc3d* h, int vec, const void* vx_dev, const void* vy_dev, const void* vz_dev, int src_dtype, void* stream);
/* solver vector -> caller MAC arrays, entries selected by `mode` (FS_STORE_*) */
int fs_visc3d_store(fs_visc3d* h, int vec, void* vx_dev, void* vy_dev, void* vz_dev, int dst_dtype, int mode, void* stream);
/* extrapolate(gres, num_iter, ...) (:472-502) in place on solver vector `vec`; uses Q as scratch */
int fs_visc3d_extrapolate(fs_visc3d* h, int vec, int sweeps, void* stream);
/* initialize_solver (:504-513): dst = RHS built from src (solid-neighbour terms moved over) */
int fs_visc3d_rhs(fs_visc3d* h, double scale, double mu, int src_vec, int dst_vec, void* stream);
/* matvecmul (:515-524): dst = A*src with the reference's neighbour masks; arbitrary src */
int fs_visc3d_apply(fs_visc3d* h, double scale, double mu, int src_vec, int dst_vec, void* stream);
/* CG on the loaded problem: expects X (extrapolated start) and B; runs q=Ax, d=r=b-q, then the loop
 * (:575-612).  Blocks until finished.  Returns FS_OK or FS_NOT_CONVERGED. */
int fs_visc3d_cg(fs_visc3d* h, double scale, double mu, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);
/* Bench hook: enqueue exactly `n` CG iterations on the current state without any host sync or
 * convergence stop (the timed window of BASELINE.json's "fixed 200 iterations" configs). */
int fs_visc3d_cg_enqueue(fs_visc3d* h, double scale, double mu, int64_t n, void* stream);
/* Profiling hook: enqueue `n` back-to-back launches of ONE kernel of the iteration on the current state
 * (which = 1: K1 apply+d.q, 2: K2 x/r update + r.r, 3: K3 d update) so bench.py can time it with CUDA events. */
int fs_visc3d_kernel_enqueue(fs_visc3d* h, int which, double scale, double mu, int64_t n, void* stream);
int fs_visc3d_read_stats(fs_visc3d* h, fs_cg_stats* stats, void* stream);
/* One call = ViscosityCGSolver3D.solve (:566-613): pack, load, 3-sweep extrapolation, RHS, CG,
 * masked write-back into vx,vy,vz (caller dtype `vel_dtype`). */
int fs_visc3d_solve(fs_visc3d* h, double dt, double mu, double rho, double cell_vol,
                    void* vx_dev, void* vy_dev, void* vz_dev, int vel_dtype,
                    const double* sphi_dev, const double* lvol_dev,
                    double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (new work; the reference is single-GPU): one process per GPU, the grid cut into x-slabs
 * (x is the storage-slowest axis, so halo planes are contiguous).  A rank builds an ordinary fs_visc3d on
 * its slab EXTENDED by one cell towards each existing neighbour and attaches a communicator; every
 * fs_visc3d_* call then exchanges one plane of u,v,w per neighbour before each stencil apply and
 * all-reduces the two CG scalars, so all ranks iterate in lock-step and stop on the same iteration.
 * Bootstrap: rank 0 calls fs_comm_unique_id, the 128 bytes are broadcast by the host program
 * (torch.distributed), every rank calls fs_comm_create.
 * ---------------------------------------------------------------------------------------- */
typedef struct fs_comm fs_comm;
int fs_comm_unique_id(void* out128);
int fs_comm_create(fs_comm** out, int rank, int nranks, const void* id128);
void fs_comm_destroy(fs_comm* c);
int fs_comm_rank(const fs_comm* c);
int fs_comm_size(const fs_comm* c);
/* Declare `h` to be a slab with a lower and/or higher neighbour (rank-1 / rank+1 of `comm`). */
int fs_visc3d_set_slab(fs_visc3d* h, fs_comm* comm, int has_lo, int has_hi);
/* Peer-memory transport (all GPUs on one NVSwitch box): device buffers that other ranks map with CUDA IPC.
 * The slab's workspace and a 1 KiB mailbox are allocated with fs_shared_alloc, their 64-byte handles travel
 * over the host program's process group, and every rank maps its neighbours' workspaces and all mailboxes. */
void* fs_shared_alloc(size_t bytes);
void fs_shared_free(void* p);
int fs_shared_get_handle(void* p, void* out64);
void* fs_shared_open(const void* handle64);
void fs_shared_close(void* mapped);
/* Switch the slab to kernel-fused collectives: K1 stores its boundary q planes straight into the neighbours'
 * halo planes and the d.q / r.r all-reduces run inside the K1 / K2 tails through the mailboxes (no NCCL call per
 * iteration).  lo_ws/hi_ws: IPC mappings of the neighbours' workspaces (NULL where there is none), lo_nx/hi_nx their
 * extended slab thickness in cells; mailboxes[r]: mapping of rank r's mailbox (own entry = local pointer). */
int fs_visc3d_set_peers(fs_visc3d* h, void* lo_ws, int lo_nx, void* hi_ws, int hi_nx, void* const* mailboxes);
/* 1 if a peer failed to answer inside a fused all-reduce (the solve then ends as not converged) */
int fs_visc3d_peer_error(fs_visc3d* h);

/* ------------------------------------------------------------------------------------------
 * Gathered multi-GPU solve — for active sets small enough to live in one GPU's L2 (the benchmark scene: 0.45 % of the
 * rows).  Splitting such a CG across GPUs only adds NVLink latency to every iteration, so here the once-per-solve passes
 * are sharded instead: every rank builds a handle for the GLOBAL grid, declares the x-window of cells its input arrays
 * cover (fs_visc3d_set_window: owned cells extended by 4 = sweeps + 1 towards each neighbour), packs / loads /
 * extrapolates only that window, and publishes one record per lattice segment the CG will touch on the planes it owns
 * (fs_visc3d_gather_export).  The host program all-gathers the records (torch.distributed / NCCL), every rank scatters
 * its peers' records into its lattice (fs_visc3d_gather_import) and runs the whole CG locally — no traffic between the
 * GPUs inside the iteration — then writes the rows of its own planes back (fs_visc3d_solve_packed).  Also usable on one
 * GPU to pack a host upload chunk by chunk.
 * ---------------------------------------------------------------------------------------- */
/* The caller's arrays hold the cells [cell_lo, cell_hi) only: vx has cell_hi-cell_lo+1 x-planes, vy/vz cell_hi-cell_lo,
 * sphi/lvol the fine planes 2*cell_lo .. 2*cell_hi.  (0, nx) restores whole-grid inputs. */
int fs_visc3d_set_window(fs_visc3d* h, int cell_lo, int cell_hi);
size_t fs_visc3d_gather_record_bytes(const fs_visc3d* h);
/* pack + load + 3-sweep extrapolation of the window, then one record per segment that holds a computed row of the
 * lattice planes [own_lo, own_hi) or is read by such a row's stencil.  *count = records written (<= cap).  Blocks. */
int fs_visc3d_gather_export(fs_visc3d* h, const void* vx_dev, const void* vy_dev, const void* vz_dev, int vel_dtype,
                            const double* sphi_dev, const double* lvol_dev, double vol_norm, int own_lo, int own_hi,
                            void* records_dev, int64_t cap, int64_t* count, void* stream);
/* *count > cap: nothing was written; grow the buffer and write the same records with this call */
int fs_visc3d_gather_reexport(fs_visc3d* h, void* records_dev, int64_t cap, void* stream);
/* records of rank r at records_dev + r*stride*record_bytes, counts_dev[r] of them (device array of nranks int64);
 * block `skip_rank` (the local one) is not re-read.  Rebuilds the active list over the whole lattice. */
int fs_visc3d_gather_import(fs_visc3d* h, const void* records_dev, const int64_t* counts_dev, int nranks, int64_t stride,
                            int skip_rank, void* stream);
/* RHS + CG on the packed lattice, write-back of the computed rows on planes [own_lo, own_hi) into the windowed arrays */
int fs_visc3d_solve_packed(fs_visc3d* h, double dt, double mu, double rho, double cell_vol,
                           void* vx_dev, void* vy_dev, void* vz_dev, int vel_dtype, int own_lo, int own_hi,
                           double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Viscosity, 2-D  (ViscosityCGSolver2D) — fluid test is sphi > 0, no extrapolation, tol 1e-4
 * lattice X=W+1, Y'=roundup(H+1,4); MAC faces u (W+1,H), v (W,H+1); fine grid (2W+1,2H+1)
 * ---------------------------------------------------------------------------------------- */
typedef struct fs_visc2d fs_visc2d;

size_t fs_visc2d_workspace_bytes(int W, int H, int dtype);
int fs_visc2d_create(fs_visc2d** out, int W, int H, int dtype, void* workspace_dev, size_t workspace_bytes);
void fs_visc2d_destroy(fs_visc2d* h);
int fs_visc2d_lattice(const fs_visc2d* h, int* X, int* Yp, int64_t* NL);
void* fs_visc2d_vector_ptr(const fs_visc2d* h, int vec, int comp);
int fs_visc2d_pack(fs_visc2d* h, const double* sphi_dev, const double* lvol_dev, double vol_norm, void* stream);
int fs_visc2d_load(fs_visc2d* h, int vec, const void* vx_dev, const void* vy_dev, int src_dtype, void* stream);
int fs_visc2d_store(fs_visc2d* h, int vec, void* vx_dev, void* vy_dev, int dst_dtype, int mode, void* stream);
int fs_visc2d_rhs(fs_visc2d* h, double scale, double mu, int src_vec, int dst_vec, void* stream);
int fs_visc2d_apply(fs_visc2d* h, double scale, double mu, int src_vec, int dst_vec, void* stream);
int fs_visc2d_cg(fs_visc2d* h, double scale, double mu, double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);
int fs_visc2d_solve(fs_visc2d* h, double dt, double mu, double rho, double cell_vol,
                    void* vx_dev, void* vy_dev, int vel_dtype,
                    const double* sphi_dev, const double* lvol_dev,
                    double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Solid fractions  (SolidFraction3D / SolidFraction2D) — fp64 in, fp64 out, reference layout.
 * Only the entries the reference writes are written (far planes keep their previous contents).
 * ---------------------------------------------------------------------------------------- */
int fs_solidfrac3d(int nx, int ny, int nz, const double* sphi_dev, double* wx_dev, double* wy_dev, double* wz_dev, void* stream);
int fs_solidfrac2d(int W, int H, const double* sphi_dev, double* wx_dev, double* wy_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pressure  (PressureCGSolver3D / 2D) — fp64 state in reference layout, operating directly on the
 * caller's CGSolverBuffer arrays (d,r,q,b) and the solver's x, all of `ncells` doubles.
 * nz == 0 selects the 2-D solver (arrays (nx,ny); vz/wz ignored; sv has 2 components).
 * ---------------------------------------------------------------------------------------- */
typedef struct fs_press fs_press;

size_t fs_press_workspace_bytes(int nx, int ny, int nz);
int fs_press_create(fs_press** out, int nx, int ny, int nz, void* workspace_dev, size_t workspace_bytes);
void fs_press_destroy(fs_press* h);
/* Operator the handle's apply / CG use: the pressure Poisson operator (default) or the density (volume-conservation)
 * operator of DensityCGSolver3D.py:117-204 — same 7-point ghost-fluid stencil with UNIT diagonal contributions and the
 * reference's -z quirk (the -z off-diagonal term reads wz[x,y,z+1]).  Same CG driver, same CGSolverBuffer arrays. */
#define FS_OP_PRESSURE 0
#define FS_OP_DENSITY 1
int fs_press_set_operator(fs_press* h, int op);
/* initialize_solver (PressureCGSolver3D.py:155-159): weighted divergence RHS into b */
int fs_press_rhs(fs_press* h, const double* cell_size3, const void* vx_dev, const void* vy_dev, const void* vz_dev, int vel_dtype,
                 const double* sv_dev, const double* lphi_dev, double* b_dev,
                 const double* wx_dev, const double* wy_dev, const double* wz_dev, void* stream);
/* matvecmul (:161-165): out = A*v on interior fluid cells, 0 on interior non-fluid, boundary untouched */
int fs_press_apply(fs_press* h, const double* v_dev, double* out_dev,
                   const double* wx_dev, const double* wy_dev, const double* wz_dev, const double* lphi_dev, void* stream);
/* apply_pressure (:167-171): velocity update + solid blend in place */
int fs_press_update(fs_press* h, const double* cell_size3, void* vx_dev, void* vy_dev, void* vz_dev, int vel_dtype,
                    const double* pv_dev, const double* wx_dev, const double* wy_dev, const double* wz_dev,
                    const double* sv_dev, const double* lphi_dev, void* stream);
/* CG loop (:198-223): x=0, q=Ax, d=r=b-q, iterate.  raise_on_fail semantics are the caller's. */
int fs_press_cg(fs_press* h, double* x_dev, double* d_dev, double* r_dev, double* q_dev, const double* b_dev,
                const double* wx_dev, const double* wy_dev, const double* wz_dev, const double* lphi_dev,
                double tol, int64_t max_iter, fs_cg_stats* stats, void* stream);
int fs_press_cg_enqueue(fs_press* h, double* x_dev, double* d_dev, double* r_dev, double* q_dev,
                        const double* wx_dev, const double* wy_dev, const double* wz_dev, const double* lphi_dev,
                        int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Density (volume-conservation) solve  (DensityCGSolver3D, SURVEY "next" row f-1) — the kernels around the CG;
 * the operator apply and the CG itself are fs_press_apply / fs_press_cg on a handle set to FS_OP_DENSITY.
 * Cells (nx,ny,nz), faces (+1 on their own axis), fine grid (2n+1)^3: fp64, reference layout.  Particles: px (P,3) and
 * pm (P), FS_F32 or FS_F64.  bound_min3 / cell_size3 / grid_bias3 are HOST arrays of three doubles.
 * ---------------------------------------------------------------------------------------- */
/* initialize_density (DensityCGSolver3D.py:8-36, :250-255): gm += w*pm, gvol += w*pvol (trilinear, fp64 atomics) */
int fs_dens3d_scatter(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, const void* px_dev, int px_dtype,
                      const void* pm_dev, int pm_dtype, int64_t num_particles, double pvol, double* gm_dev, double* gvol_dev, void* stream);
/* fix_volume (:38-92, :257-264): in place on gvol, interior cells */
int fs_dens3d_fix_volume(int nx, int ny, int nz, const double* cell_size3, double* gvol_dev, const double* sphi_dev, const double* lphi_dev,
                         const double* wx_dev, const double* wy_dev, const double* wz_dev, void* stream);
/* initialize_solver (:94-125, :266-272): b on interior cells (0 on non-fluid ones), boundary layer untouched */
int fs_dens3d_rhs(int nx, int ny, int nz, double rho0, double dt, const double* cell_size3, const double* gm_dev, const double* gvol_dev,
                  const double* lphi_dev, const double* wx_dev, const double* wy_dev, const double* wz_dev, double* b_dev, void* stream);
/* compute_displacement (:206-219, :280-284): dx,dy,dz[x,y,z] for x,y,z in 1..g-1 */
int fs_dens3d_displacement(int nx, int ny, int nz, double dt, const double* cell_size3, double* dx_dev, double* dy_dev, double* dz_dev,
                           const double* pv_dev, const double* lphi_dev, void* stream);
/* apply_displacement (:221-248, :286-291): px[:,axis] += trilinear gather of d (shape s0,s1,s2) */
int fs_dens3d_gather(void* px_dev, int px_dtype, int64_t num_particles, const double* d_dev, int s0, int s1, int s2,
                     const double* bound_min3, const double* cell_size3, const double* grid_bias3, int axis, void* stream);

/* ------------------------------------------------------------------------------------------
 * Grid-side steps of the APIC time loop either side of the implicit solves: the kernels the notebook
 * defines inline (3D_viscous_fluid_sim.ipynb code cells 2-7).  They produce the lvol / lphi / velocity / mass arrays the
 * solvers consume and carry the result back to the particles, so a whole time step can stay on the device.
 * Types are the notebook's: particle arrays fp64 (x, v, c* as (P,3) row-major; m as (P)), MAC grids fp32, level-set /
 * volume grids fp64.  bound_min3 / cell_size3: 3 host doubles (bound_min is rounded to fp32 like the notebook's array).
 * ---------------------------------------------------------------------------------------- */
/* p2g (:279-344): zero-initialised mass / momentum grids in, mass and mass-weighted velocity (v / m where m > 0) out */
int fs_grid_p2g(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np,
                const double* px, const double* pm, const double* pv, const double* cx, const double* cy, const double* cz,
                float* mx, float* vx, float* my, float* vy, float* mz, float* vz, void* stream);
/* g2p (:352-393): particle velocity and the three affine rows from the grid velocities */
int fs_grid_g2p(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np,
                const double* px, double* pv, double* cx, double* cy, double* cz, const float* vx, const float* vy, const float* vz, void* stream);
/* compute_fluid_levelset (:94-136): phi = fill, then min over particles of |x_cell - x_p| - radius in a 5^3 neighbourhood */
int fs_grid_levelset(int nx, int ny, int nz, const double* bound_min3, const double* cell_size3, int64_t np, const double* px,
                     double radius, double fill, double* phi, void* stream);
/* compute_fluid_volume (:224-268) on the (rx,ry,rz) = 2*gres+1 node grid: trilinear splat of pvol, clamped to cell_vol */
int fs_grid_fluid_volume(int rx, int ry, int rz, const double* bound_min3, const double* cell_size3, int64_t np, const double* px,
                         double pvol, double cell_vol, double* vol, void* stream);
/* extrapolate (:501-557): `sweeps` Jacobi sweeps into faces with zero mass, in place; workspace = generation bytes */
size_t fs_grid_extrapolate_workspace_bytes(int nx, int ny, int nz);
int fs_grid_extrapolate(int nx, int ny, int nz, int sweeps, float* vx, float* vy, float* vz, const float* mx, const float* my, const float* mz,
                        void* workspace_dev, size_t workspace_bytes, void* stream);
/* apply_boundary_condition (cell 5): dv from the old velocities (all three components), then v += dv */
int fs_grid_boundary(int nx, int ny, int nz, double dx, float* vx, float* vy, float* vz, const float* mx, const float* my, const float* mz,
                     const double* sphi, const double* sv, float* dvx, float* dvy, float* dvz, void* stream);

/* ------------------------------------------------------------------------------------------
 * Rigid-body signed distance field (solver/sdf3D.py:218-279).  rb_d: device table [nbodies][10][4] fp64 as
 * built by sdf3D.generate_rb (:294-327); positions (npos,3) fp64 row-major.
 * ---------------------------------------------------------------------------------------- */
/* sd[p] = min over bodies (starting from 100, like the reference); vel[p] = velocity of that body where sd <= 0, else 0 */
int fs_sdf3d_evaluate(const double* rb_d, int nbodies, int64_t npos, const double* pos_dev, double* sd_dev, double* vel_dev, void* stream);
/* push every position out of (or, for flipped bodies, into) each body in table order, in place */
int fs_sdf3d_project(const double* rb_d, int nbodies, int64_t npos, double* pos_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * UNet surrogate (3D_viscous_fluid_sim.ipynb:844-913): the network input built in one pass and the gather
 * of the velocity increments.  (X,Y,Z) = padded volume ("data_size"), pads = int((data_size - (2n+1)) / 2) like the notebook.
 * ---------------------------------------------------------------------------------------- */
/* out: fp32 [11][X][Y][Z] = dxdx dydy dzdz dxdy dxdz dydx dydz dzdx dzdy, solid flag (sphi <= 0; pad_solid outside the grid),
 * lvol / cell_vol (the notebook passes 0.0125**3).  vx, vy, vz: fp32 MAC arrays; sphi, lvol: fp64 (2n+1)^3. */
int fs_unet_features(int nx, int ny, int nz, int X, int Y, int Z, const float* vx, const float* vy, const float* vz,
                     const double* sphi, const double* lvol, double cell_vol, float pad_solid, float* out, void* stream);
/* dv{x,y,z}: fp32 MAC arrays <- net_out[3][X][Y][Z] at the staggered positions, divided by `divisor` (= int(1/DT)) */
int fs_unet_gather(int nx, int ny, int nz, int X, int Y, int Z, const float* net_out, double divisor,
                   float* dvx, float* dvy, float* dvz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDSOLVER_B200_H */
