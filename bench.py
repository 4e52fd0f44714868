#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: walk-on-stars walks/s AND simulation steps/s, next to the reference's CPU path.

Default workload = BASELINE.json configs[4]: the 3D zombie3d solve, 1e6 query points x 500 walks per step on the
smoke3d scene (closed cube, the scene of smoke3d / smoke_obs / vortex_collide), 82^3 divergence grid.  STRONG scaling:
the 1e6 points are split across the ranks (one process per GPU, sharding.shard_bounds), every rank solves its block and
the estimates are all-gathered over NCCL INSIDE the timed region, so every rank ends a step holding all N x (1 + dim)
floats, as the time-stepper needs them.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]      our CUDA path
  python bench.py --impl reference ...                                       the reference's CPU solver (oracle/_ref)

Prints ONE JSON line (rank 0):
  value            walks/s, inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e              same metric through the C ABI with HOST (pinned) buffers: H2D points, solve, D2H p and grad p
  e2e_python       same through the compiled drop-in module zombie_bindings: wost() (nested lists, the reference's
                   signature) and the additive wost_array() (numpy)                                   [N = 1]
  sim_steps_per_sec  one operator-split time step (advect fit, divergence grid, solve, projection fit) of the example
                   this workload belongs to, at K = 1000 and K = 10000 Adam iterations per fit
  roofline         HBM: algorithmic bytes per launch / kernel time, per GPU, against MEASURED_PEAKS.json
  roofline_issue   the binding bound: warp instructions issued per second against the measured issue peak
  cpu_baseline     the reference's own solver headers on the host cores, bounded sample                 [N = 1]
  also             the other headline workloads (karman 1e5 points = BASELINE configs[1], karman3d 1e6), short runs
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_WALK = {2: 8.04, 3: 8.06}  # algorithmic HBM bytes per walk, SURVEY.md section 8(d)
WORKLOADS = {  # name -> (case, total query points per step, example whose time step `sim_steps_per_sec` runs)
    "smoke3d_1000000pts_x500walks": ("smoke3d", 1000000, "smoke3d"),
    "karman3d_1000000pts_x500walks": ("karman3d", 1000000, "karman3d"),
    "karman_100000pts_x500walks": ("karman", 100000, "karman"),
    "taylorgreen_262144pts_x500walks": ("taylorgreen_active", 262144, "taylorgreen"),
    "smoke3d_65536pts_x500walks": ("smoke3d", 65536, "smoke3d"),
    "channel_circle_100000pts_x500walks": ("channel_circle", 100000, None),
    "box_sphere_100000pts_x500walks": ("box_sphere", 100000, None),
}
WL = None  # the package's workloads module (set in main)
DEFAULT_WORKLOAD = "smoke3d_1000000pts_x500walks"
ALSO = ["karman_100000pts_x500walks", "karman3d_1000000pts_x500walks"]


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6550.0}, "fallback (B200_PROFILING.md)"


def kernel_profile(dim):
    """Per-build ncu figures of the walk kernel (profiles/walk_kernel_profile.json, written from the committed
    captures by profiles/summarize.py): DRAM traffic per launch, warp instructions per walk, active lanes."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "walk_kernel_profile.json")))
        return prof.get("fastKernel<%d>" % dim)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---- the reference's CPU solver (oracle/_ref, else the plain-C restatement) ------------------------------------------
def cpu_reference_rate(cfg, src, pts, threads, seed=1):
    from oracle import refbind, oraclebind
    dim = cfg["dim"]
    if refbind.available(dim):
        sc, kind = refbind.RefScene(dim, cfg["scene"], src), "reference"
    else:
        sc, kind = oraclebind.OracleScene(dim, cfg["scene"], src), "port"
    t = time.perf_counter()
    p, g, st = sc.wost(cfg["solver"], cfg["output"], pts, seed=seed, nthreads=threads, want_stats=True)
    dt = time.perf_counter() - t
    walks = int((st[:, 11] > 0).sum())*WL.walks_per_point(cfg["solver"])
    sc.close()
    return walks/dt, walks, dt, kind


REF_BUILD = ("the reference's solver headers (walk_on_stars.h, FCPW) compiled by oracle/Makefile: scalar SBVH (no FCPW_USE_ENOKI), "
             "std::thread over points instead of TBB; per core it is faster than the stock Enoki/TBB build (SURVEY 6.2), "
             "so ratios against it are conservative")


def run_reference(args, cfg, src, lo, hi, n_total):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    probe = WL.random_points(lo, hi, 1024, seed=100)
    rate, _, _, kind = cpu_reference_rate(cfg, src, probe, threads)
    n = int(min(n_total, max(1024, rate*5.0/500)))  # bounded sample: ~5 s of CPU work per step
    times, walks = [], 0
    for s in range(args.warmup + args.steps):
        pts = WL.random_points(lo, hi, n, seed=200 + s)
        r, w, dt, kind = cpu_reference_rate(cfg, src, pts, threads, seed=s)
        if s >= args.warmup:
            times.append(dt); walks += w
    total = sum(times)
    val = walks/total
    line = {"impl": "reference", "metric": "wost_walks_per_sec", "value": val, "unit": "walks/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3*total/max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "case": args.case, "points_per_step_total": n_total, "points_per_step_sample": n,
                       "nWalks": cfg["solver"]["nWalks"], "source_grid": list(src.shape),
                       "mode": "reference CPU solver, %d host threads; %s" % (threads, REF_BUILD)},
            "cpu_baseline": {"value": val, "unit": "walks/s", "cores": threads, "kind": kind,
                             "sample": "%d of %d points x %d walks per step" % (n, n_total, cfg["solver"]["nWalks"])},
            "e2e": {"value": val, "unit": "walks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm ------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def walk_workload(X, case, n_total, steps, warmup, mode, timed_gather=True, want_e2e=True, sample_clocks=False):
    """Strong-scaled walk workload: returns a dict with device-timed and end-to-end numbers (rank-reduced)."""
    import torch
    import torch.distributed as dist
    pkg, capi, dev, rank, world = X.pkg, X.pkg.capi, X.dev, X.rank, X.world
    cfg = WL.load_case(case)
    dim = cfg["dim"]
    src = WL.source_grid(case)
    scene = pkg.Scene(cfg["scene"], src, device=X.local)
    lo, hi = scene.bbox()
    opts = pkg.zombie.solver_opts(cfg["solver"], cfg["output"], mode=mode, seed=1234)
    b0, b1 = pkg.sharding.shard_bounds(n_total, rank, world)
    n = b1 - b0
    m = -(-n_total//world)  # padded shard size: equal blocks for all_gather_into_tensor
    total_steps = warmup + steps
    # synthetic query sets, one per step, resident in HBM before the timed region; every rank draws the same global
    # set (same seed) and keeps its block, so the union over ranks does not depend on the number of ranks
    tlo, thi = torch.tensor(lo, device=dev), torch.tensor(hi, device=dev)
    pts = []
    for s in range(total_steps):
        gen = torch.Generator(device=dev); gen.manual_seed(1000 + s)
        full = torch.rand((n_total, dim), generator=gen, device=dev)*(thi - tlo) + tlo
        pts.append(full[b0:b1].contiguous())
        del full
    p_loc = torch.zeros(m, device=dev); g_loc = torch.zeros((m, dim), device=dev)
    p_all = torch.empty(m*world, device=dev) if world > 1 else p_loc
    g_all = torch.empty((m*world, dim), device=dev) if world > 1 else g_loc
    stream = torch.cuda.current_stream().cuda_stream
    stats = capi.SolveStats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, timed):
        X.flush.fill_(float(i))  # L2 flush between iterations (256 MiB write), outside the timed events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scene.handle.solve_device(opts, pts[i].data_ptr(), n, p_loc.data_ptr(), g_loc.data_ptr(), index_offset=b0, stream=stream,
                                  stats=stats if timed else None)
        if world > 1 and timed_gather:  # the final gather of the point estimates (north_star), NCCL over NVLink
            dist.all_gather_into_tensor(p_all, p_loc)
            dist.all_gather_into_tensor(g_all, g_loc)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for i in range(warmup):
        step(i, False)
    barrier()
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(X.local); sampler.start()
    ms = walks = kms = launches = wsteps = 0
    lane = (0, 0)
    for i in range(warmup, total_steps):
        ms += step(i, True)
        walks += stats.walks_started; kms += stats.kernel_ms; launches += stats.kernel_launches; wsteps += stats.walk_steps
        lane = (lane[0] + stats.lane_slices, lane[1] + stats.warp_trips)
    barrier()
    clocks = sampler.summary() if sampler else None
    out = {"case": case, "dim": dim, "cfg": cfg, "src_shape": list(src.shape), "n_total": n_total, "n_local": n, "lo": lo, "hi": hi,
           "walks_local": walks, "kernel_ms_local": kms, "clocks": clocks, "lane": lane, "scene": scene, "opts": opts,
           "finite": bool(torch.isfinite(p_all).all().item() and torch.isfinite(g_all).all().item())}

    e2e_s, e2e_walks = 0.0, 0
    if want_e2e:  # the C ABI with host (pinned) buffers: H2D of the points, solve, D2H of p and grad p
        hp = [p.cpu().pin_memory() for p in pts[warmup:]] or [pts[0].cpu().pin_memory()]
        hpo = torch.empty(n).pin_memory(); hgo = torch.empty((n, dim)).pin_memory()
        st2 = capi.SolveStats()
        scene.handle.solve_ptr(opts, hp[0].data_ptr(), n, hpo.data_ptr(), hgo.data_ptr(), index_offset=b0, stats=st2)
        barrier()
        for i in range(len(hp)):
            X.flush.fill_(float(i)); torch.cuda.synchronize()
            t0 = time.perf_counter()
            scene.handle.solve_ptr(opts, hp[i].data_ptr(), n, hpo.data_ptr(), hgo.data_ptr(), index_offset=b0, stats=st2)
            e2e_s += time.perf_counter() - t0
            e2e_walks += st2.walks_started
        barrier()
        out["host_points"] = hp[0]

    if world > 1:  # max time over ranks, total walks over ranks
        t = torch.tensor([ms, e2e_s, kms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([walks, e2e_walks, wsteps, launches], device=dev, dtype=torch.float64); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ms, e2e_s, kms_max = t.tolist(); walks, e2e_walks, wsteps, launches = c.tolist()
    out.update(ms=ms, walks=walks, e2e_s=e2e_s, e2e_walks=e2e_walks, wsteps=wsteps, launches=int(launches),
               value=walks/(ms*1e-3) if ms > 0 else 0.0, ms_per_step=ms/max(steps, 1),
               e2e=(e2e_walks/e2e_s if e2e_s > 0 else None))
    return out


def python_e2e(X, R, reps=2):
    """The Python-visible drop-in: `import zombie_bindings` as src/2d / src/3d do, Scene(config, grid), then wost()
    (nested lists out, the reference's signature) and the additive wost_array() (numpy out); numpy points in."""
    dim = R["dim"]
    moddir = os.path.join(os.path.dirname(X.pkg.__file__), "zombie%dd" % dim)
    sys.path.insert(0, moddir)
    try:
        import zombie_bindings as zb
    finally:
        sys.path.remove(moddir)
    cfg = R["cfg"]
    sc = zb.Scene(dict(cfg["scene"], device=X.local), WL.source_grid(R["case"]))
    pts = R["host_points"].numpy()
    zb.set_mode("fast"); zb.set_seed(77)
    zb.wost_array(sc, cfg["solver"], cfg["output"], pts[:1024])
    out = {}
    for name, fn in (("wost_array_numpy", zb.wost_array), ("wost_lists", zb.wost)):
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = fn(sc, cfg["solver"], cfg["output"], pts)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
            del res
        walks = zb.last_stats()["walks_started"]
        out[name] = {"value": walks/best, "unit": "walks/s", "ms_per_call": 1e3*best}
    out["points"] = int(len(pts))
    out["note"] = ("wost(): the reference's signature, returns 3 nested Python lists (N x (2 dim + 1) float objects built "
                   "inside the call); wost_array(): additive numpy entry point of the same module")
    return out


def sim_steps(X, example, iters_list=(1000, 10000), steps=(2, 1)):
    """Simulation steps/s of the example the workload belongs to, through the device-resident stepper."""
    import torch
    import bench_step
    from importlib import import_module
    st = import_module(X.pkg.__name__ + ".stepper")
    out = {}
    a = argparse.Namespace(case=example, iters=iters_list[0], watertight=False, no_graph=False, device=X.local, distributed=X.world > 1)
    s, cfg, init_fn, shape, what = bench_step.build(a, X.pkg, st)
    s.fit_initial(init_fn, 200, lr=1e-3)
    s.step(50)  # warm-up: captures the two fit graphs
    for iters, nst in zip(iters_list, steps):
        torch.cuda.synchronize()
        if X.world > 1:
            import torch.distributed as dist
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            s.step(iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if X.world > 1:
            t = torch.tensor([dt], device=X.dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        out["K%d" % iters] = {"value": nst/dt, "unit": "steps/s", "ms_per_step": 1e3*dt/nst, "steps_timed": nst,
                              "pressure_ms": s.last.get("pressure_ms"), "walks_per_step": int(s.last.get("walks", 0))}
    out["config"] = what
    out["parallelism"] = ("fits replicated on every rank (identical samples, weights re-broadcast after each fit), pressure samples sharded + all_gather; "
                          "data-parallel fits (SplitStepper(fit_parallel='data')) are slower while an iteration is latency-bound" if X.world > 1 else "single GPU")
    s.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--case", default=None, help="override the workload's scene (with --points)")
    ap.add_argument("--points", type=int, default=None, help="override the TOTAL query points per step")
    ap.add_argument("--mode", default="fast", choices=["fast", "deterministic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sim-steps", action="store_true")
    ap.add_argument("--no-python-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    import __graft_entry__ as ge
    global WL
    pkg = ge.load_package()
    WL = pkg.workloads
    case, n_total, example = WORKLOADS[args.workload]
    if args.case or args.points:
        case = args.case or case
        n_total = args.points or n_total
        example = {"taylorgreen_active": "taylorgreen", "karman": "karman", "smoke3d": "smoke3d", "karman3d": "karman3d"}.get(case)
        args.workload = "%s_%dpts_x500walks" % (case, n_total)
    args.case = case
    cfg = WL.load_case(case)
    dim = cfg["dim"]

    if args.impl == "reference":
        v, _ = pkg.zombie.load_obj(cfg["scene"]["boundary"], dim)  # host-side OBJ reader of the product
        return run_reference(args, cfg, WL.source_grid(case), v.min(axis=0), v.max(axis=0), n_total)

    import torch
    import torch.distributed as dist
    X = Ctx()
    X.pkg = pkg
    X.rank = int(os.environ.get("RANK", "0")); X.world = int(os.environ.get("WORLD_SIZE", "1")); X.local = int(os.environ.get("LOCAL_RANK", "0"))
    if pkg.capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(X.local)
    X.dev = torch.device("cuda", X.local)
    if X.world > 1:
        dist.init_process_group("nccl", device_id=X.dev)
    X.flush = torch.empty(256*1024*1024//4, device=X.dev)  # > 126 MB L2
    mode = pkg.capi.MODE_FAST if args.mode == "fast" else pkg.capi.MODE_DETERMINISTIC
    world = X.world

    R = walk_workload(X, case, n_total, args.steps, args.warmup, mode, sample_clocks=True)
    also = {}
    if not args.no_also:
        for name in ALSO:
            if name == args.workload:
                continue
            c2, n2, _ = WORKLOADS[name]
            r2 = walk_workload(X, c2, n2, 5, 3, mode, want_e2e=False)
            also[name] = {"value": r2["value"], "unit": "walks/s", "ms_per_step": r2["ms_per_step"], "steps": 5, "n_gpus": world}
            r2["scene"].handle.close()
    sim = None
    if not args.no_sim_steps and example is not None:
        try:
            sim = sim_steps(X, example)
        except Exception as e:  # the walk numbers stand on their own; say why the other half is missing
            sim = {"error": "%s: %s" % (type(e).__name__, e)}
    pe2e = None
    if X.rank == 0 and world == 1 and not args.no_python_e2e:
        try:
            pe2e = python_e2e(X, R)
        except Exception as e:
            pe2e = {"error": "%s: %s" % (type(e).__name__, e)}
    issue_peak = pkg.capi.measure_peaks(X.local) if X.rank == 0 else None
    dispatch = pkg.capi.measure_issue_peak(X.local) if X.rank == 0 else None

    if X.rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk, which = peaks()
    steps = max(args.steps, 1)
    # per-GPU roofline of the dominant kernel (rank 0's launches): algorithmic bytes per launch / kernel time
    walks_per_launch = R["walks_local"]/steps
    k_ms = R["kernel_ms_local"]/steps
    ach = walks_per_launch*B_WALK[dim]/(k_ms*1e-3)/1e9
    prof = kernel_profile(dim)
    prof_ok = prof is not None and prof.get("workload") == args.workload
    line = {"metric": "wost_walks_per_sec", "value": R["value"], "unit": "walks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": R["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "case": case, "points_per_step_total": n_total, "points_per_step_per_gpu": R["n_local"],
                       "nWalks": cfg["solver"]["nWalks"], "mode": args.mode, "source_grid": R["src_shape"],
                       "l2": "flushed (256 MiB write) between timed steps",
                       "timed_region": "walk kernel on this rank's block" + (" + NCCL all_gather of p and grad p (%d bytes per rank out)" % (R["n_total"]*(1 + dim)*4) if world > 1 else ""),
                       "walk_steps_per_walk": R["wsteps"]/max(R["walks"], 1),
                       "lane_occupancy_of_slice_loop": (R["lane"][0]/(32.0*R["lane"][1]) if R["lane"][1] else None),
                       "parallelism": "points sharded (strong scaling), scene replicated, dp%d" % world, "also": also},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach/pk["hbm_gbs"],
                         "traffic": prof.get("dram_bytes_per_launch") if prof_ok else None, "per": "GPU (rank 0's launches)", "peak_source": which,
                         "kernel": "fastKernel<%d>" % dim, "algorithmic_bytes_per_launch": walks_per_launch*B_WALK[dim], "kernel_ms": k_ms,
                         "note": "HBM is not the binding bound of the walk kernel (source texels are L2 hits); see roofline_issue"},
            "e2e": {"value": R["e2e"], "unit": "walks/s", "h2d_bytes_per_step": R["n_local"]*dim*4*world, "d2h_bytes_per_step": R["n_local"]*(1 + dim)*4*world,
                    "through": "nmc_wost_solve (C ABI, pinned host buffers), every rank its block"},
            "e2e_python": pe2e, "sim_steps_per_sec": sim,
            "gpu_launches": R["launches"], "clocks": R["clocks"], "finite": R["finite"]}
    if issue_peak:
        rate = walks_per_launch/(k_ms*1e-3)  # walks/s of this GPU inside the kernel
        ri = {"bound": "issue", "peak": dispatch[0], "unit": "warp-inst/s",
              "peak_source": "measured live (nmc_measure_issue_peak): 4 dispatch slots per SM x SMs x the SM clock under load (clock64 vs globaltimer)",
              "sm_clock_hz": dispatch[1], "fma_pipe_peak": issue_peak[0], "mufu_peak": issue_peak[1], "fp64_fma_peak": issue_peak[2], "per": "GPU"}
        if prof is not None and prof.get("warp_inst_per_walk"):
            ri.update(achieved=rate*prof["warp_inst_per_walk"], frac=rate*prof["warp_inst_per_walk"]/dispatch[0],
                      warp_inst_per_walk=prof["warp_inst_per_walk"], active_lanes_per_inst=prof.get("active_lanes_per_inst"),
                      lane_issue_frac=rate*prof["warp_inst_per_walk"]*prof.get("active_lanes_per_inst", 32.0)/32.0/dispatch[0],
                      profile=prof.get("capture"), profile_workload=prof.get("workload"))
        line["roofline_issue"] = ri
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        probe = WL.random_points(R["lo"], R["hi"], 1024, seed=100)
        rate, _, _, _ = cpu_reference_rate(cfg, WL.source_grid(case), probe, threads)
        m = int(min(n_total, max(1024, rate*15.0/500)))  # ~15 s of CPU work
        rate, w, dt, kind = cpu_reference_rate(cfg, WL.source_grid(case), WL.random_points(R["lo"], R["hi"], m, seed=101), threads)
        line["cpu_baseline"] = {"value": rate, "unit": "walks/s", "cores": threads, "kind": kind,
                                "sample": "%d of %d points x %d walks, %.1f s" % (m, n_total, cfg["solver"]["nWalks"], dt), "build": REF_BUILD}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
