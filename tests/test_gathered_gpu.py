"""GPU: the gathered multi-GPU solve (solver/distributed.py::GatheredViscosityCGSolver3D) with the ranks EMULATED in one
process on one GPU — the record exchange becomes a concatenation, everything else (windowed pack / load / extrapolation
of each rank's x-window of the global lattice, publication of the segments the CG touches, import, CG on the complete
active set, write-back of the owned planes) is the code the NCCL path runs.  Checked against the single-GPU solver on the
whole grid: identical iteration counts (+-1: the active lists are equal, so the arithmetic is the same up to the order in
which equal partial sums are formed) and velocities to 1e-9.  The real multi-process run is tests/dist_check.py (needs N GPUs)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _solve_both(gres, N, mu, world, dtype=torch.float64, active_set="nonzero", max_iter=None, tol=1e-3, lvol_cut=None):
    import scenes
    from solver.distributed import GatheredViscosityCGSolver3D, SlabPartition, emulate_gathered_solve, scatter_scene
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    full = scenes.buckling(N, device="cuda", mu=mu, gres=gres)
    g = full["gres"]
    parts = [SlabPartition(g, world, r, ext=4) for r in range(world)]
    solvers = [GatheredViscosityCGSolver3D(g, full["bound_size"], dtype=dtype, partition=p, active_set=active_set) for p in parts]
    ref = ViscosityCGSolver3D(g, full["bound_size"], dtype=dtype, active_set=active_set)
    if max_iter is not None:
        ref.max_iter = max_iter
        for s in solvers:
            s.max_iter = max_iter
    results = []
    for step in range(2 if lvol_cut else 1):
        if step == 1:                                   # second solve on the same objects with a different liquid region
            lv = full["lvol"].clone()
            lv[:, : int(lvol_cut * lv.shape[1])] = 0.0  # the pool at the bottom disappears: the active set shrinks
            full["lvol"] = lv
        per_rank = [scatter_scene(full, p) for p in parts]
        for sc in per_rank:
            for k in ("vx", "vy", "vz"):
                sc[k] = sc[k].clone()
        rv = [full[k].clone() for k in ("vx", "vy", "vz")]
        err_ref = err_g = None
        try:
            ref.solve(full["dt"], mu, full["rho"], *rv, full["sphi"], None, None, full["lvol"], tol=tol)
        except ValueError as e:
            err_ref = e
        try:
            emulate_gathered_solve(solvers, per_rank, full["dt"], mu, full["rho"], tol=tol)
        except ValueError as e:
            err_g = e
        assert (err_ref is None) == (err_g is None)
        results.append((full, parts, solvers, per_rank, ref, rv))
    return results[-1]


@pytest.mark.parametrize("gres,N,mu,world,aset", [(None, 48, 100.0, 2, "nonzero"), (None, 48, 100.0, 3, "nonzero"), ((37, 24, 28), 32, 10.0, 4, "nonzero"),
                                                  ((64, 20, 24), 32, 100.0, 8, "nonzero"), ((37, 24, 28), 32, 10.0, 3, "fluid")])
def test_gathered_matches_single_gpu(gres, N, mu, world, aset):
    full, parts, solvers, per_rank, ref, rv = _solve_both(gres, N, mu, world, active_set=aset)
    assert ref.iterations > 10
    for s in solvers:
        assert abs(s.iterations - ref.iterations) <= 1, (s.iterations, ref.iterations)
        assert s.iterations == solvers[0].iterations          # every rank takes the same decisions
        assert s.active_info() == ref.active_info()            # the imported lattice holds the same active set
    # assemble the global result from the owned planes of every rank's window arrays
    for k, kind, r in (("vx", "u", 0), ("vy", "v", 1), ("vz", "w", 2)):
        out = torch.empty_like(rv[r])
        for p, sc in zip(parts, per_rank):
            lo, hi = p.owned_planes(kind)
            out[p.e0 + lo: p.e0 + hi] = sc[k][lo:hi]
        assert rel_l2(out.cpu().numpy(), rv[r].cpu().numpy()) < 1e-9, k
        # rows outside a rank's owned planes are left as passed in
        for p, sc in zip(parts, per_rank):
            lo, hi = p.owned_planes(kind)
            orig = p.slab(full[k], kind)
            assert torch.equal(sc[k][:lo], orig[:lo]) and torch.equal(sc[k][hi:], orig[hi:])


def test_gathered_second_solve_with_changed_liquid():
    """stale-state handling: the active set shrinks between two solves on the same objects"""
    full, parts, solvers, per_rank, ref, rv = _solve_both(None, 32, 10.0, 3, lvol_cut=0.2)
    for s in solvers:
        assert abs(s.iterations - ref.iterations) <= 1
        assert s.active_info() == ref.active_info()
    out = torch.empty_like(rv[1])
    for p, sc in zip(parts, per_rank):
        lo, hi = p.owned_planes("v")
        out[p.e0 + lo: p.e0 + hi] = sc["vy"][lo:hi]
    assert rel_l2(out.cpu().numpy(), rv[1].cpu().numpy()) < 1e-9


def test_gathered_fixed_window_raises_on_every_rank():
    full, parts, solvers, per_rank, ref, rv = _solve_both(None, 32, 100.0, 2, max_iter=25, tol=0.0)
    for s in solvers:
        assert s.iterations == 25 == ref.iterations
        assert abs(s.delta - ref.delta) <= 1e-9 * ref.delta
