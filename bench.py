#!/usr/bin/env python
"""Benchmark of the north-star path: 3-D viscosity CG iterations/s on the 256^3 high-viscosity buckling scene
(BASELINE.json config 4; SURVEY.md §8 d), 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (C/OpenMP port of the reference's loop)

A "step" is one fixed window of `--iters` (default 200) CG iterations of ViscosityCGSolver3D on the scene,
started from scratch every step exactly the way config 1 prescribes it for the reference (max_iter=200, tol=0:
the solve packs the inputs, extrapolates, builds the RHS, runs 200 iterations and raises "Failed to converge!").
`value`  = CG iterations/s with every input already resident in HBM (device tensors passed to solve()).
`e2e`    = the same call with HOST (pinned) buffers: H2D of vx,vy,vz,sphi,lvol and D2H of the solution inside the
           timed region.
`roofline` = the dominant kernel of the iteration timed alone with CUDA events (fs_visc3d_kernel_enqueue),
           algorithmic bytes per launch / duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline` = oracle/c_port (fp64 C/OpenMP restatement) timed on this box's host cores on a bounded sample.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "python-fluid-simulation_b200")
for _p in (PKG, REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "viscosity_cg_iters_per_s_256^3"
UNIT = "CG iterations/s"


def counts(n):
    """faces F and used fine-grid volume classes V7 of an n^3 grid (SURVEY §8 d)"""
    F = 3 * n * n * (n + 1)
    V7 = F + n ** 3 + 3 * (n + 1) * (n + 1) * n
    return F, V7


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML in-process every ~1 ms when available (the timed
    region of the default run is only tens of milliseconds), else `nvidia-smi` every 0.2 s."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index = index
        self.rows = []          # [sm_mhz, sm_max_mhz, set(reasons)]
        self.source = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        self._h = None
        self._max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-free boxes: LOCAL_RANK indexes the physical devices the driver gave us
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        if self._max is None:                      # constant: asked once, so that a sample is two NVML calls (the timed region is ~10 ms)
            self._max = float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM))
        mx = self._max
        try:
            bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        self.rows.append([sm, mx, {k for k, b in self.NVML_BITS.items() if bits & b}])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [v.strip() for v in out.split(",")]
            self.rows.append([float(c[0]), float(c[1]), {n for n, v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")}])

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._nvml is not None:          # NVML hiccup: fall back to nvidia-smi for the rest of the run
                    self._nvml = None
                    self.source = "nvidia-smi"
            self._stop.wait(0.001 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = sorted(set().union(*[r[2] for r in self.rows])) if self.rows else []
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------------------------------------------------
# reference arm: the CPU port on the host cores
# ------------------------------------------------------------------------------------------------------------

def run_reference(args):
    """CPU arm: the fp64 C/OpenMP restatement of the reference's solve() on this box's host cores (the reference itself
    has no CPU implementation).  One solve is prepared exactly like the GPU step (extrapolation, RHS, first apply: timed
    once); every timed step then runs a bounded sample of `--ref-iters` CG iterations of the same 200-iteration window, and
    the set-up time is charged pro rata, so the figure is iterations/s of the WHOLE step like the GPU arm's.  With
    `--ref-full-step` (default at N=1) one complete 200-iteration step is also timed end to end, once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import scenes
    from oracle import c_port
    cores = c_port.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1: size the team to the host cores explicitly
    n = args.size
    sc = scenes.buckling(n, device="cpu", mu=args.mu)
    s = c_port.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    arrs = [sc[k].numpy() for k in ("vx", "vy", "vz", "sphi", "lvol")]
    t0 = time.perf_counter()
    st = s.prepare(sc["dt"], args.mu, sc["rho"], *arrs)
    t_prep = time.perf_counter() - t0
    sample = args.ref_iters
    delta = st["delta"]

    def step(delta):
        _, d = c_port.cg(sc["gres"], st["scale"], args.mu, st["x"], st["r"], st["d"], st["q"], st["sphi"], st["vol"], 0.0, sample, delta)
        return d

    for _ in range(args.warmup):
        delta = step(delta)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        delta = step(delta)
    dt = time.perf_counter() - t0
    per_iter = dt / (sample * args.steps)
    step_s = t_prep + args.iters * per_iter                 # one whole step: set-up + the fixed iteration window
    value = args.iters / step_s
    full = None
    if args.ref_full_step:
        t0 = time.perf_counter()
        st2 = s.prepare(sc["dt"], args.mu, sc["rho"], *arrs)
        c_port.cg(sc["gres"], st2["scale"], args.mu, st2["x"], st2["r"], st2["d"], st2["q"], st2["sphi"], st2["vol"], 0.0, args.iters, st2["delta"])
        tf = time.perf_counter() - t0
        full = {"seconds": tf, "value": args.iters / tf, "what": f"one complete step (extrapolation + RHS + first apply + {args.iters} CG iterations), timed once"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n, "gpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"set-up of one solve timed once ({t_prep:.2f} s) + {sample} CG iterations per timed step of the same {n}^3 "
                                   f"buckling scene ({1e3 * per_iter:.1f} ms per iteration); value = {args.iters} / (set-up + {args.iters} x per-iteration time); "
                                   "fp64 C/OpenMP restatement (oracle/c_port)",
                         "setup_s": t_prep, "ms_per_iteration": 1e3 * per_iter, "full_step": full},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU implementation (Numba-CUDA only); this arm times a line-by-line C/OpenMP port of its solve()",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n, where):
    if getattr(args, "scene", "buckling") != "buckling":
        return {"workload": f"viscous-column-{n}^3 (dense fluid, kernel study; NOT the BASELINE config)", "grid": [n, n, n], "mu": args.mu,
                "iters_per_step": args.iters, "l2": "working set >> L2"}
    return {"workload": f"buckling-{n}^3 high-viscosity (mu={args.mu:g}), ViscosityCGSolver3D fixed {args.iters}-iteration CG window per step"
                        if where != "cpu" else f"buckling-{n}^3 high-viscosity (mu={args.mu:g}), ViscosityCGSolver3D CG iterations",
            "grid": [n, n, n], "mu": args.mu, "dt": 1.0 / 300, "rho": 1000.0, "iters_per_step": args.iters,
            "partition": f"x-slabs over {args.gpus} GPU(s)" if args.gpus > 1 else "single GPU",
            "active_set": getattr(args, "active_set", "nonzero") + " (the CG kernels visit only rows with a non-zero operator row; same iterates, "
                          "iteration counts and velocities as the dense loop: tests/test_active_set_gpu.py)"
                          if getattr(args, "active_set", "nonzero") == "nonzero" else "fluid (every row the reference's kernels compute)",
            "l2": "working set >> L2 (inputs larger than L2, no explicit flush)"}


# ------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------

def run_native(args):
    # keep stdout to the one JSON line: anything the native libraries print (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    args._json_out = os.fdopen(json_fd, "w")
    import numpy as np
    import torch
    import torch.distributed as dist

    import scenes
    from solver import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fluidsolver_b200_nccl_%h_%p.log")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    n = args.size
    F, V7 = counts(n)
    words_iter = 11 * F + V7
    peak, peak_src = hbm_peak()

    if world > 1:
        from solver.distributed import bench_distributed
        return bench_distributed(args, METRIC, UNIT, workload_config(args, n, "gpu"), peak, peak_src)

    lib = N.load()
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(n, device="cuda", mu=args.mu) if args.scene == "buckling" else scenes.viscous_column((n, n, n), device="cuda", mu=args.mu)
    solver = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], dtype=tdtype, active_set=args.active_set, cg_mode=args.cg_mode)
    solver.max_iter = args.iters
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def window(scene, steps, warm):
        """`steps` fixed-window solves with device-resident inputs: (ms per step, launches per step)"""
        dev_in = [scene[k] for k in ("vx", "vy", "vz")]

        def step():
            try:
                solver.solve(scene["dt"], args.mu, scene["rho"], *dev_in, scene["sphi"], None, None, scene["lvol"], tol=0.0)
            except ValueError:
                pass                              # "Failed to converge!" after exactly max_iter iterations (reference :611-612)
            assert solver.iterations == args.iters, solver.iterations

        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        l0 = N.launch_count()
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps, (N.launch_count() - l0) / steps

    with ClockSampler(local) as clocks:
        ms_step, launches_step = window(sc, args.steps, max(args.warmup, 3))
    ms = ms_step * args.steps
    launches = int(round(launches_step * args.steps))
    value = args.iters / (ms_step * 1e-3)

    # ---- e2e: the same call with host (pinned) buffers -------------------------------------------------------
    host = {k: sc[k].cpu().pin_memory() for k in ("vx", "vy", "vz", "sphi", "lvol")}
    # the step's result in the caller's own precision: the API writes fp32 vx, vy, vz (reference :613), so that is what comes back
    out_host = [torch.empty(tuple(a.shape), dtype=host["vx"].dtype).pin_memory() for a in (solver.x_x, solver.x_y, solver.x_z)]
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = sum(a.numel() * a.element_size() for a in out_host)

    def step_e2e():
        try:
            solver.solve(sc["dt"], args.mu, sc["rho"], host["vx"], host["vy"], host["vz"], host["sphi"], None, None, host["lvol"], tol=0.0)
        except ValueError:
            pass
        for o, x in zip(out_host, (solver.x_x, solver.x_y, solver.x_z)):
            o.copy_(x, non_blocking=True)         # the step's result: the solution after the window
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks_e2e:
        ev0.record()
        for _ in range(e2e_steps):
            step_e2e()
        ev1.record()
        torch.cuda.synchronize()
    e2e_ms = ev0.elapsed_time(ev1)
    e2e_value = args.iters * e2e_steps / (e2e_ms * 1e-3)
    del host

    # ---- roofline of the default leg: the CG window alone, and each kernel of the iteration alone ------------------
    scale = sc["dt"] / solver.cell_vol / sc["rho"]
    kinfo = kernel_study(solver, lib, N, torch, scale, args, esz)
    persistent = kinfo["persistent"]
    iter_ms, kern, kbytes, iter_bytes = kinfo["iter_ms"], kinfo["kernel_ms"], kinfo["kernel_bytes"], kinfo["iter_bytes"]
    l2_resident = kinfo["working_set"] < 100e6
    traffic, traffic_src, ipl = None, None, min(args.iters, 256)
    key = None
    try:                                              # dram bytes per launch from the committed ncu --set full capture of this command
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        dom_key = ("persistent_sr" if persistent and kinfo["sr"] else "persistent" if persistent else max(kern, key=kern.get).split()[0])
        key = args.scene + ":" + args.active_set + ":" + dom_key
        traffic = tj.get(key)
        traffic_src = tj.get("_source", {}).get(key) if isinstance(tj.get("_source"), dict) else None
        ipl = int(tj.get("_iterations_per_launch", {}).get(key, ipl)) if isinstance(tj.get("_iterations_per_launch"), dict) else ipl
    except Exception:
        pass
    if persistent:
        # the iteration runs as ONE persistent kernel (its phases are the kernels above): that kernel is the dominant one of the step
        dom = kinfo["persistent_name"]
        achieved = iter_bytes / (iter_ms * 1e-3) / 1e9
        dom_share = iter_ms * args.iters / ms_step
        per_launch = {"iterations_per_launch": ipl, "algorithmic_bytes_per_launch": ipl * iter_bytes,
                      "note": f"achieved = active-set bytes / duration, both per launch of {ipl} iterations (one launch runs the whole {args.iters}-iteration "
                              "window plus the closing evaluation of r.r); traffic = dram bytes of one such launch (ncu)"}
    else:
        dom = max(kern, key=kern.get)
        achieved = kbytes[dom] / (kern[dom] * 1e-3) / 1e9
        dom_share = kern[dom] * args.iters / ms_step
        per_launch = {"iterations_per_launch": 1, "algorithmic_bytes_per_launch": kbytes[dom]}
    iter_gbs = words_iter * esz * value / 1e9
    if l2_resident:
        # The active set of this scene (0.45 % of the face rows) is an L2-resident problem: the kernel is bound by grid-barrier
        # and L2 latency, not by HBM.  `achieved` is the rate at which it walks its (L2-resident) working set and is NOT
        # compared with the HBM peak; the DRAM-side figure is traffic / duration.  The HBM-bound regimes of the SAME kernels
        # are measured below (hbm_leg), against the HBM roofline.
        dram_gbs = (traffic / (ipl * iter_ms * 1e-3) / 1e9) if (traffic and persistent) else None
        roofline = {"bound": "l2-latency", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": None,
                    "achieved_is": "active-set bytes per second served from L2 (not DRAM); no HBM fraction is claimed for this leg",
                    "dram_GBps": dram_gbs, "dram_frac_of_peak": (dram_gbs / peak) if dram_gbs else None}
    else:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak}
    roofline.update({
        "traffic": traffic, "traffic_source": traffic_src or "profiles/ncu_traffic.json (ncu --set full capture, per launch)",
        "peak_source": peak_src, "share_of_step": dom_share, "launch": per_launch,
        "cg_iteration_us": iter_ms * 1e3, "cg_iteration_bytes": iter_bytes, "setup_ms_per_step": ms_step - iter_ms * args.iters,
        "per_kernel_ms_standalone": kern, "per_kernel_GBps_standalone": {k: kbytes[k] / (kern[k] * 1e-3) / 1e9 for k in kern},
        "bytes_per_launch": kbytes, "cg_mode": kinfo["mode"],
        "active_set": dict(kinfo["active"], faces=F, row_fraction=kinfo["active"]["computed_rows"] / F, l2_resident=l2_resident),
        "dense_equivalent": {"algorithmic_GB_per_iter": words_iter * esz / 1e9, "equivalent_GBps": iter_gbs,
                             "note": "what a kernel streaming every face row (11F+V7 words, SURVEY 8d) would need to move to match this "
                                     "iteration rate; the active-set kernels do not move these bytes, so this is a speed-up figure, not a bandwidth"}})

    # ---- CPU baseline (bounded sample) on this box's host cores ----------------------------------------------
    cpu = cpu_baseline(args, sc, solver, lib, scale)

    # ---- HBM-bound regimes of the same kernels, in the same run ----------------------------------------------------
    hbm_variant = None
    if args.hbm_leg and args.scene == "buckling":
        # (1) the same scene with every fluid row visited (what the reference's kernels compute): 627 MB CG working set
        solver._e.set_active_mode("fluid")
        with ClockSampler(local) as cl:
            ms_f, _ = window(sc, max(2, min(args.steps, 5)), 2)
        k2 = kernel_study(solver, lib, N, torch, scale, args, esz)
        hbm_variant = {"what": "same scene and window, active_set='fluid' (every row the reference's kernels compute)", "value": args.iters / (ms_f * 1e-3),
                       "unit": UNIT, "ms_per_step": ms_f, "cg_iteration_us": k2["iter_ms"] * 1e3, "cg_mode": k2["mode"], "active_set": k2["active"],
                       "per_kernel_ms": k2["kernel_ms"], "per_kernel_frac_of_hbm_peak": {k: k2["kernel_bytes"][k] / (k2["kernel_ms"][k] * 1e-3) / 1e9 / peak for k in k2["kernel_ms"]},
                       "iteration_frac_of_hbm_peak_active_bytes": k2["iter_bytes"] / (k2["iter_ms"] * 1e-3) / 1e9 / peak, "clocks": cl.summary()}
        solver._e.set_active_mode(args.active_set)
        # (2) dense liquid (viscous column, 93 % of the rows are fluid): the regime SURVEY 8d's 11F+V7 accounting describes
        del sc
        torch.cuda.empty_cache()
        col = scenes.viscous_column((n, n, n), device="cuda", mu=args.mu)
        solver._e.set_active_mode("fluid")
        with ClockSampler(local) as cl:
            ms_c, _ = window(col, 2, 1)
            k3 = kernel_study(solver, lib, N, torch, col["dt"] / solver.cell_vol / col["rho"], args, esz, win=30)
        kb = {"K1": (2 * F + V7) * esz, "K2": 9 * F * esz} if k3["sr"] else {"K1": (2 * F + V7) * esz, "K2": 6 * F * esz, "K3": 3 * F * esz}
        per_k = {}
        for name, t in k3["kernel_ms"].items():
            per_k[name] = {"ms": t, "algorithmic_bytes": kb[name.split()[0][:2]], "GBps": kb[name.split()[0][:2]] / (t * 1e-3) / 1e9,
                           "frac": kb[name.split()[0][:2]] / (t * 1e-3) / 1e9 / peak}
        it_gbs = words_iter * esz / (k3["iter_ms"] * 1e-3) / 1e9
        roofline["hbm_leg"] = {"what": f"viscous-column-{n}^3 (dense liquid), active_set='fluid': the HBM-bound regime of the same kernels, 30-iteration window",
                               "bound": "hbm", "achieved": it_gbs, "peak": peak, "unit": "GB/s", "frac": it_gbs / peak,
                               "accounting": "(11F+V7) words per iteration, SURVEY 8d: K1 2F+V7, update kernels 9F", "bytes_per_iteration": words_iter * esz,
                               "cg_iteration_ms": k3["iter_ms"], "iters_per_s": 1e3 / k3["iter_ms"], "cg_mode": k3["mode"], "per_kernel": per_k,
                               "active_set": k3["active"], "step_ms": ms_c, "clocks": cl.summary()}
        del col

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, n, "gpu"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "clocks": clocks_e2e.summary()},
        "gpu_launches": launches, "roofline": roofline, "hbm_variant": hbm_variant, "cpu_baseline": cpu, "clocks": clocks.summary(),
    }
    print(json.dumps(line), file=args._json_out, flush=True)
    return 0


def kernel_study(solver, lib, N, torch, scale, args, esz, win=None):
    """CUDA-event timings on the solver's CURRENT operator / active set: the CG window alone (per-iteration time) and each
    kernel of the iteration alone.  Bytes: the kernels walk the ACTIVE 32-point lattice segments only, so the
    unit is the lattice point of an active segment.  Two-reduction CG: K1 13 words + 1 activity byte, K2 18, K3 9;
    single-reduction CG: K1s 13 words + 1 byte (reads r, 7 coefficients; writes w), K2s 27 (reads r, w, p, s, x; writes p, s, x, r)."""
    stream = torch.cuda.current_stream().cuda_stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    segs, segs_total, rows = solver.active_info()
    pts = segs * 32
    mode = N.check(lib.fs_visc3d_cg_mode_in_use(solver._e.h), "mode")
    persistent = mode in (N.CG_PERSISTENT, N.CG_PERSISTENT_SR)
    sr = mode in (N.CG_KERNELS_SR, N.CG_PERSISTENT_SR)
    if sr:
        names = ((1, "K1s visc3d_apply_dot2"), (2, "K2s cg_update_sr"))
        kbytes = {names[0][1]: pts * (13 * esz + 1), names[1][1]: pts * 27 * esz}
    else:
        names = ((1, "K1 visc3d_apply_dot"), (2, "K2 cg_update_xr"), (3, "K3 cg_update_d"))
        kbytes = {names[0][1]: pts * (13 * esz + 1), names[1][1]: pts * 18 * esz, names[2][1]: pts * 9 * esz}
    N.check(lib.fs_visc3d_cg_enqueue(solver._e.h, scale, args.mu, 64, stream), "warm")
    torch.cuda.synchronize()
    if win is None:
        win = 512 if persistent else 192
    ev0.record()
    N.check(lib.fs_visc3d_cg_enqueue(solver._e.h, scale, args.mu, win, stream), "window")
    ev1.record()
    torch.cuda.synchronize()
    iter_ms = ev0.elapsed_time(ev1) / win
    kern = {}
    reps = 30
    for which, name in names:
        N.check(lib.fs_visc3d_kernel_enqueue(solver._e.h, which, scale, args.mu, 3, stream), "warm")
        torch.cuda.synchronize()
        ev0.record()
        N.check(lib.fs_visc3d_kernel_enqueue(solver._e.h, which, scale, args.mu, reps, stream), "time")
        ev1.record()
        torch.cuda.synchronize()
        kern[name] = ev0.elapsed_time(ev1) / reps
    modes = {N.CG_KERNELS: "kernels", N.CG_PERSISTENT: "persistent", N.CG_KERNELS_SR: "kernels_sr", N.CG_PERSISTENT_SR: "persistent_sr"}
    return {"iter_ms": iter_ms, "kernel_ms": kern, "kernel_bytes": kbytes, "iter_bytes": sum(kbytes.values()), "working_set": pts * (22 * esz + 1),
            "persistent": persistent, "sr": sr, "mode": modes[mode],
            "persistent_name": ("visc3d_cg_sr_resident2_kernel (phases A: apply + both dots, B: fused update; point-private data, r and segment ids resident in "
                                "shared memory; one cooperative launch per up to 256 iterations)" if sr else
                                "visc3d_cg_persistent_kernel (K1+K2+K3 phases of one cooperative launch per 64 iterations)"),
            "active": {"mode": "fluid" if solver._e.active_mode_name == "fluid" else "nonzero", "segments": segs, "segments_total": segs_total,
                       "computed_rows": rows, "cg_working_set_MB": pts * (22 * esz + 1) / 1e6}}


def cpu_baseline(args, sc, solver, lib, scale):
    """oracle/c_port on the host cores, a few iterations of the same workload from the same CG state."""
    import numpy as np
    import torch
    from oracle import c_port
    from solver import _native as N
    try:
        c_port.use_all_cores()                   # (torchrun exports OMP_NUM_THREADS=1)
        solver.max_iter = 0
        try:
            solver.solve(sc["dt"], args.mu, sc["rho"], sc["vx"], sc["vy"], sc["vz"], sc["sphi"], None, None, sc["lvol"], tol=0.0)
        except ValueError:
            pass
        solver.max_iter = args.iters
        x = [a.double().contiguous().cpu().numpy() for a in (solver.x_x, solver.x_y, solver.x_z)]
        r = [a.double().contiguous().cpu().numpy() for a in (solver.r_x, solver.r_y, solver.r_z)]
        d = [a.copy() for a in r]
        q = [np.zeros_like(a) for a in r]
        vol = (sc["lvol"] / (solver.cell_vol * 0.125)).cpu().numpy()
        sphi = sc["sphi"].cpu().numpy()
        delta = float(sum(np.sum(a * a) for a in r))
        c_port.cg(sc["gres"], scale, args.mu, x, r, d, q, sphi, vol, 0.0, 1, delta)      # warm-up (page-in, threads)
        k = args.ref_iters
        t0 = time.perf_counter()
        it, _ = c_port.cg(sc["gres"], scale, args.mu, x, r, d, q, sphi, vol, 0.0, k, delta)
        dt = time.perf_counter() - t0
        return {"value": it / dt, "unit": UNIT, "cores": c_port.num_threads(), "kind": "port",
                "sample": f"{k} CG iterations of the same {args.size}^3 scene/state, fp64 C/OpenMP restatement (oracle/c_port), {dt:.1f} s"}
    except Exception as e:                                           # the baseline must never sink the GPU number
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=200, help="CG iterations per step (fixed window)")
    ap.add_argument("--ref-iters", type=int, default=4, help="CG iterations per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--ref-full-step", type=int, default=None, help="CPU arm: also time one complete 200-iteration step (default: on at N=1)")
    ap.add_argument("--hbm-leg", type=int, default=1, help="also measure the HBM-bound regimes of the same kernels (roofline.hbm_leg, hbm_variant)")
    ap.add_argument("--mu", type=float, default=100.0)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="solver storage/arithmetic type (reference: f64)")
    ap.add_argument("--active-set", default="nonzero", choices=["nonzero", "fluid"],
                    help="rows the CG kernels visit: nonzero (default) = rows with a non-zero coefficient; fluid = every row the reference computes")
    ap.add_argument("--cg-mode", default="auto", choices=["auto", "kernels", "persistent", "kernels_sr", "persistent_sr"],
                    help="three kernels per iteration from a CUDA graph, one persistent cooperative kernel, or auto by working-set size")
    ap.add_argument("--scene", default="buckling", choices=["buckling", "column"],
                    help="buckling = BASELINE config 4 (default); column = dense-fluid viscous column (config 5 geometry) for kernel studies")
    args = ap.parse_args()
    if args.ref_full_step is None:
        args.ref_full_step = 1 if args.gpus == 1 else 0
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
