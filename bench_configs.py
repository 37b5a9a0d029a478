#!/usr/bin/env python
"""Measurements for every BASELINE.json config (SURVEY.md §8 d); bench.py stays the single-line driver contract for
config 4.  Prints one JSON object per config and a markdown table (paste into BASELINE.md).

    python bench_configs.py [--configs 1,2,3,4,5] [--dtype f64]
    torchrun --nproc-per-node 8 bench_configs.py --configs 5w      # weak scaling, 512^3 cells per GPU

All timings: CUDA events around the public solve() calls with device-resident inputs, after one warm-up solve.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "python-fluid-simulation_b200"), REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402


def _timed(fn, reps=3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        ev0.record()
        out = fn()
        ev1.record()
        torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1)
        best = t if best is None else min(best, t)
    return best, out


def counts3(n):
    F = 3 * n * n * (n + 1)
    V7 = F + n ** 3 + 3 * (n + 1) * (n + 1) * n
    return F, V7


def peak():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def fixed_window(solver, sc, mu, iters):
    solver.max_iter = iters

    def run():
        try:
            solver.solve(sc["dt"], mu, sc["rho"], sc["vx"], sc["vy"], sc["vz"], sc["sphi"], None, None, sc["lvol"], tol=0.0)
        except ValueError:
            pass
        return solver.iterations
    run()
    return _timed(run)


def cpu_window(sc, mu, iters):
    from oracle import c_port
    s = c_port.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    st = s.prepare(sc["dt"], mu, sc["rho"], sc["vx"].cpu().numpy(), sc["vy"].cpu().numpy(), sc["vz"].cpu().numpy(),
                   sc["sphi"].cpu().numpy(), sc["lvol"].cpu().numpy())
    c_port.cg(sc["gres"], st["scale"], mu, st["x"], st["r"], st["d"], st["q"], st["sphi"], st["vol"], 0.0, 1, st["delta"])
    t0 = time.perf_counter()
    it, _ = c_port.cg(sc["gres"], st["scale"], mu, st["x"], st["r"], st["d"], st["q"], st["sphi"], st["vol"], 0.0, iters, st["delta"])
    return it / (time.perf_counter() - t0), c_port.num_threads()


def config1(tdtype, esz):
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    n = 64
    sc = scenes.buckling(n, device="cuda", mu=1.0)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], dtype=tdtype)
    ms, its = fixed_window(s, sc, 1.0, 200)
    F, V7 = counts3(n)
    gbs = (11 * F + V7) * esz * 200 / (ms * 1e-3) / 1e9
    cpu, cores = cpu_window(sc, 1.0, 40)
    return {"config": "1: 64^3 buckling, 200 fixed viscosity-CG iterations", "ms_per_solve": ms, "iterations": its, "it_per_s": 200e3 / ms,
            "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak(), "note": "working set (~80 MB) is L2-resident: latency/launch-bound, not HBM-bound",
            "cpu_port_it_per_s": cpu, "cpu_cores": cores}


def config2(tdtype, esz):
    import scenes
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver2D import PressureCGSolver2D
    from solver.ViscosityCGSolver2D import ViscosityCGSolver2D
    W = 1024
    sc = scenes.box2d(W, device="cuda")
    visc = ViscosityCGSolver2D(sc["gres"], sc["bound_size"], dtype=tdtype)
    press = PressureCGSolver2D(CGSolverBuffer(sc["gres"]), sc["gres"], sc["bound_size"])

    def step():
        v = [sc["vx"].clone(), sc["vy"].clone()]
        visc.solve(sc["dt"], sc["mu"], sc["rho"], *v, sc["sphi"], sc["sv"], sc["lphi"], sc["lvol"])
        press.solve(*v, sc["sphi"], sc["sv"], sc["lphi"])
        return visc.iterations, press.iterations
    step()
    ms, (iv, ip) = _timed(step)

    def only_visc():
        v = [sc["vx"].clone(), sc["vy"].clone()]
        visc.solve(sc["dt"], sc["mu"], sc["rho"], *v, sc["sphi"], sc["sv"], sc["lphi"], sc["lvol"])
    msv, _ = _timed(only_visc)
    return {"config": "2: 2-D 1024^2 box, ViscosityCGSolver2D (tol 1e-4) + SolidFraction2D + PressureCGSolver2D (tol 1e-3) per timestep",
            "ms_per_step": ms, "visc_ms": msv, "press_ms": ms - msv, "visc_iterations": iv, "press_iterations": ip,
            "visc_us_per_iter": 1e3 * msv / max(iv, 1), "press_us_per_iter": 1e3 * (ms - msv) / max(ip, 1),
            "note": "L2-resident (~0.2 GB): latency-bound; iterations run from CUDA graphs"}


def config3(tdtype, esz):
    import scenes
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver3D import PressureCGSolver3D
    from solver.SolidFraction3D import compute_solid_frac
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    n = 128
    sc = scenes.buckling(n, device="cuda", mu=1.0, with_sv=True)
    g = sc["gres"]
    visc = ViscosityCGSolver3D(g, sc["bound_size"], dtype=tdtype)
    press = PressureCGSolver3D(CGSolverBuffer(g), g, sc["dx"])          # bound_size = GDX, as the notebook does
    w = [torch.zeros(s, dtype=torch.float64, device="cuda") for s in ((g[0] + 1, g[1], g[2]), (g[0], g[1] + 1, g[2]), (g[0], g[1], g[2] + 1))]
    stage = {}

    def step():
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        compute_solid_frac(g, sc["sphi"], *w)
        e[1].record()
        visc.solve(sc["dt"], sc["mu"], sc["rho"], *v, sc["sphi"], sc["sv"], sc["lphi"], sc["lvol"])
        e[2].record()
        press.solve(*v, sc["sphi"], sc["sv"], sc["lphi"], wx=w[0], wy=w[1], wz=w[2])
        e[3].record()
        torch.cuda.synchronize()
        stage.update(solidfrac_ms=e[0].elapsed_time(e[1]), visc_ms=e[1].elapsed_time(e[2]), press_ms=e[2].elapsed_time(e[3]))
        return visc.iterations, press.iterations
    step()
    ms, (iv, ip) = _timed(step)
    F, V7 = counts3(n)
    return {"config": "3: 128^3 buckling full step (SolidFraction3D + ViscosityCGSolver3D + PressureCGSolver3D), tol 1e-3, mu=1",
            "ms_per_step": ms, **stage, "visc_iterations": iv, "press_iterations": ip,
            "visc_it_per_s": 1e3 * iv / stage["visc_ms"], "press_it_per_s": 1e3 * ip / stage["press_ms"],
            "visc_algorithmic_GBps": (11 * F + V7) * esz * iv / (stage["visc_ms"] * 1e-3) / 1e9,
            "press_algorithmic_GBps": 15 * n ** 3 * 8 * ip / (stage["press_ms"] * 1e-3) / 1e9}


def config4(tdtype, esz):
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    n = 256
    sc = scenes.buckling(n, device="cuda", mu=100.0)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], dtype=tdtype)

    def full():
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
        return s.iterations, s.delta
    ms, (its, delta) = _timed(full, reps=2)
    F, V7 = counts3(n)
    gbs = (11 * F + V7) * esz * its / (ms * 1e-3) / 1e9
    return {"config": "4: 256^3 buckling, mu=100, full ViscosityCGSolver3D.solve to tol=1e-3 (1 GPU)", "solve_wall_ms": ms, "iterations": its,
            "final_delta": delta, "it_per_s_over_whole_solve": 1e3 * its / ms, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak()}


def config5(tdtype, esz):
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    n = 512
    sc = scenes.viscous_column((n, n, n), device="cuda", mu=100.0)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], dtype=tdtype)
    ms, its = fixed_window(s, sc, 100.0, 100)
    F, V7 = counts3(n)
    gbs = (11 * F + V7) * esz * 100 / (ms * 1e-3) / 1e9
    return {"config": "5 (1-GPU leg): 512^3 viscous column, 100 fixed viscosity-CG iterations", "ms_per_solve": ms, "iterations": its,
            "it_per_s": 100e3 / ms, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak(),
            "note": "dense-fluid scene: K1 computes almost every row (no solid skipping)"}


def config5_weak(tdtype, esz):
    """weak scaling: 512^3 cells per GPU, global 512G x 512 x 512, slab-generated scene"""
    import torch.distributed as dist
    import scenes
    from solver.distributed import SlabPartition, SlabViscosityCGSolver3D
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fluidsolver_b200_nccl_%h_%p.log")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(os.environ.get("FS_WEAK_N", "512"))
    g = (n * world, n, n)
    part = SlabPartition(g, world, rank)
    sc = scenes.viscous_column(part.local_gres, device="cuda", mu=100.0, x0=part.e0, gx_total=g[0])
    bound = tuple(k * sc["dx"] for k in g)
    s = SlabViscosityCGSolver3D(g, bound, dtype=tdtype, partition=part)
    s.max_iter = 100

    def run():
        try:
            s.solve(sc["dt"], 100.0, sc["rho"], sc["vx"], sc["vy"], sc["vz"], sc["sphi"], None, None, sc["lvol"], tol=0.0)
        except ValueError:
            pass
        return s.iterations
    run()
    dist.barrier()
    ms, its = _timed(run, reps=2)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        F = sum((g[0] + (a == 0)) * (g[1] + (a == 1)) * (g[2] + (a == 2)) for a in range(3))
        V7 = F + g[0] * g[1] * g[2] + (g[0] + 1) * (g[1] + 1) * g[2] + (g[0] + 1) * g[1] * (g[2] + 1) + g[0] * (g[1] + 1) * (g[2] + 1)
        gbs = (11 * F + V7) * esz * 100 / (float(t.item()) * 1e-3) / 1e9
        out = {"config": f"5: weak scaling, {n}^3 cells per GPU, {world} GPU(s), 100 fixed iterations", "n_gpus": world, "ms_per_solve": float(t.item()),
               "iterations": its, "it_per_s": 100e3 / float(t.item()), "algorithmic_GBps_all_gpus": gbs, "frac_of_aggregate_hbm_peak": gbs / (peak() * world)}
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    args = ap.parse_args()
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    fns = {"1": config1, "2": config2, "3": config3, "4": config4, "5": config5, "5w": config5_weak}
    for c in args.configs.split(","):
        r = fns[c](tdtype, esz)
        if r is not None:
            r["dtype"] = args.dtype
            print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
