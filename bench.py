#!/usr/bin/env python
"""bench.py -- walks/s of the Monte Carlo pressure-projection solve (BASELINE.json metric).

One "step" = one wost() pass over one batch of synthetic query points on the karman scene
(BASELINE.json configs[1]: 2D channel + cylinder, walk-on-stars Neumann boundary, 1e5 query points/step,
shipped wost.json: nWalks 500, lambda 350, Russian roulette 0.99).  Weak scaling: every rank (one process
per GPU) solves its own 1e5-point block; the only exchange is the final all_gather of the estimates.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                            the reference's CPU solver (oracle/_ref)

Prints ONE JSON line (rank 0).  `value` = walks/s with inputs resident in HBM, CUDA events on the launch
stream, max over ranks; `e2e` = the same metric through the C ABI with HOST (pinned) buffers, copies
inside the timed region.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402  (fixtures + synthetic inputs shared with the tests)

B_WALK = {2: 8.04, 3: 8.06}  # algorithmic HBM bytes per walk, SURVEY.md section 8(d)
# dram__bytes_read.sum + dram__bytes_write.sum of fastKernel<2> for the default workload, one launch, from the
# ncu --set full capture summarised in profiles/r01_fastKernel2d_capture4.txt: 2.50 MB read (source grid + points,
# once each) + 15.02 MB written.  The kernel's own stores are 1.2 MB; the rest of the writes are dirty lines of the
# 256 MiB L2-flush buffer being evicted while the kernel runs.  The 8 B/walk of texel gathers (402 MB algorithmic)
# are served by L2, so the DRAM traffic is far below the algorithmic figure.
TRAFFIC_BYTES_DEFAULT_WORKLOAD = 17522432


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


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
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_rate(cfg, src, pts, threads, seed=1):
    """walks/s of the reference's own solver headers (oracle/_ref) -- or, if absent, the oracle port."""
    from oracle import refbind, oraclebind
    dim = cfg["dim"]
    if refbind.available(dim):
        sc, kind = refbind.RefScene(dim, cfg["scene"], src), "reference"
    else:
        sc, kind = oraclebind.OracleScene(dim, cfg["scene"], src), "port"
    t = time.perf_counter()
    p, g, st = sc.wost(cfg["solver"], cfg["output"], pts, seed=seed, nthreads=threads, want_stats=True)
    dt = time.perf_counter() - t
    nw = cfg["solver"].get("nWalks", 128)
    walks = int((st[:, 11] > 0).sum())*2*max(1, nw//2)
    sc.close()
    return walks/dt, walks, dt, kind


def run_reference(args, cfg, src, lo, hi):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    dim = cfg["dim"]
    # bounded sample of the workload: size it so that one step is ~5 s of CPU work
    probe = util.random_points(lo, hi, 1024, seed=100)
    rate, _, _, kind = cpu_reference_rate(cfg, src, probe, threads)
    n = int(min(args.points, max(1024, rate*5.0/500)))
    times, walks = [], 0
    for s in range(args.warmup + args.steps):
        pts = util.random_points(lo, hi, n, seed=200 + s)
        r, w, dt, kind = cpu_reference_rate(cfg, src, pts, threads, seed=s)
        if s >= args.warmup:
            times.append(dt); walks += w
    total = sum(times)
    val = walks/total
    line = {"impl": "reference", "metric": "wost_walks_per_sec", "value": val, "unit": "walks/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3*total/max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "case": args.case, "points_per_step_full": args.points, "points_per_step_sample": n,
                       "nWalks": cfg["solver"]["nWalks"], "mode": "reference CPU solver, %d host threads" % threads},
            "cpu_baseline": {"value": val, "unit": "walks/s", "cores": threads, "kind": kind,
                             "sample": "%d of %d points x %d walks per step" % (n, args.points, cfg["solver"]["nWalks"])},
            "e2e": {"value": val, "unit": "walks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--case", default="karman", choices=list(util.CASES))
    ap.add_argument("--points", type=int, default=100000, help="query points per step per GPU")
    ap.add_argument("--mode", default="fast", choices=["fast", "deterministic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.workload = "%s_%dpts_x500walks" % (args.case, args.points)

    cfg = util.load_case(args.case)
    dim = cfg["dim"]
    if dim == 2:  # divergence grid of the shape the time-stepper produces for this scene (401 x 1002)
        h, w = (401, 1002) if args.case == "karman" else (1002, 1002)
        y, x = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
        src = (np.sin(6.1*x)*np.sin(4.3*y + 0.3)).astype(np.float32)
    else:
        g = np.linspace(0, 1, 82)
        x, y, z = np.meshgrid(g, g, g, indexing="ij")
        src = (np.sin(6.1*x)*np.sin(4.3*y + 0.3)*np.cos(3*z)).astype(np.float32)

    import __graft_entry__ as ge
    v, _ = ge.load_package().zombie.load_obj(cfg["scene"]["boundary"], dim)  # host-side OBJ reader of the product
    lo, hi = v.min(axis=0), v.max(axis=0)

    if args.impl == "reference":
        return run_reference(args, cfg, src, lo, hi)

    import torch
    import torch.distributed as dist
    pkg = ge.load_package()
    capi = pkg.capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene = pkg.Scene(cfg["scene"], src, device=local)
    mode = capi.MODE_FAST if args.mode == "fast" else capi.MODE_DETERMINISTIC
    opts = pkg.zombie.solver_opts(cfg["solver"], cfg["output"], mode=mode, seed=1234)
    n = args.points
    total_steps = args.warmup + args.steps
    # synthetic query sets, one per step, resident in HBM before the timed region; global index offsets per rank
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    tlo, thi = torch.tensor(lo, device=dev), torch.tensor(hi, device=dev)
    pts = [(torch.rand((n, dim), generator=gen, device=dev)*(thi - tlo) + tlo).contiguous() for _ in range(total_steps)]
    p_out = torch.empty(n, device=dev); g_out = torch.empty((n, dim), device=dev)
    flush = torch.empty(256*1024*1024//4, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    stats = capi.SolveStats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, timed):
        flush.fill_(float(i))  # L2 flush between iterations, outside the timed events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scene.handle.solve_device(opts, pts[i].data_ptr(), n, p_out.data_ptr(), g_out.data_ptr(), index_offset=rank*n, stream=stream,
                                  stats=stats if timed else None)
        e1.record()
        torch.cuda.synchronize()
        step.lane = (stats.lane_slices, stats.warp_trips)
        return e0.elapsed_time(e1), stats.walks_started, stats.kernel_ms, stats.kernel_launches, stats.walk_steps

    for i in range(args.warmup):
        step(i, False)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ms, walks, kms, launches, wsteps = 0.0, 0, 0.0, 0, 0
    for i in range(args.warmup, total_steps):
        t, w, k, l, s = step(i, True)
        ms += t; walks += w; kms += k; launches += l; wsteps += s
    barrier()
    clocks = sampler.summary()

    # end-to-end through the C ABI with host (pinned) buffers: H2D of the points, solve, D2H of p and grad p
    hp = [p.cpu().pin_memory() for p in pts[args.warmup:]] or [pts[0].cpu().pin_memory()]
    hpo = torch.empty(n).pin_memory(); hgo = torch.empty((n, dim)).pin_memory()
    st2 = capi.SolveStats()
    scene.handle.solve_ptr(opts, hp[0].data_ptr(), n, hpo.data_ptr(), hgo.data_ptr(), index_offset=rank*n, stats=st2)
    barrier()
    e2e_s, e2e_walks = 0.0, 0
    for i in range(len(hp)):
        flush.fill_(float(i)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        scene.handle.solve_ptr(opts, hp[i].data_ptr(), n, hpo.data_ptr(), hgo.data_ptr(), index_offset=rank*n, stats=st2)
        e2e_s += time.perf_counter() - t0
        e2e_walks += st2.walks_started
    barrier()

    if world > 1:  # max time over ranks, total walks over ranks; one gather of the estimates as the API would do
        t = torch.tensor([ms, e2e_s, kms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([walks, e2e_walks, wsteps], device=dev, dtype=torch.float64); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ms, e2e_s, kms = t.tolist(); walks, e2e_walks, wsteps = c.tolist()
        pkg.sharding.gather_estimates(p_out, g_out, n*world, dim)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = walks/(ms*1e-3)
    pk, which = peaks()
    ach = (walks/max(args.steps, 1))*B_WALK[dim]/((kms/max(args.steps, 1))*1e-3)/1e9  # GB/s, algorithmic bytes per launch / kernel time
    line = {"metric": "wost_walks_per_sec", "value": value, "unit": "walks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms/max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "case": args.case, "points_per_step_per_gpu": n, "nWalks": cfg["solver"]["nWalks"],
                       "mode": args.mode, "source_grid": list(src.shape), "l2": "flushed (256 MiB write) between timed steps",
                       "walk_steps_per_walk": wsteps/max(walks, 1),
                       "lane_occupancy_of_slice_loop": (step.lane[0]/(32.0*step.lane[1]) if getattr(step, "lane", (0, 0))[1] else None), "parallelism": "points sharded, scene replicated, dp%d" % world},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach/pk["hbm_gbs"],
                         "traffic": TRAFFIC_BYTES_DEFAULT_WORKLOAD if (args.case == "karman" and n == 100000 and args.mode == "fast") else None,
                         "peak_source": which,
                         "note": "walk kernel is instruction-issue bound (transcendentals + BVH traversal), HBM fraction is small by construction; see DESIGN.md"},
            "e2e": {"value": e2e_walks/e2e_s, "unit": "walks/s", "h2d_bytes_per_step": n*dim*4, "d2h_bytes_per_step": n*(1 + dim)*4},
            "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        probe = util.random_points(lo, hi, 1024, seed=100)
        rate, _, _, _ = cpu_reference_rate(cfg, src, probe, threads)
        m = int(min(n, max(1024, rate*15.0/500)))  # ~15 s of CPU work
        rate, w, dt, kind = cpu_reference_rate(cfg, src, util.random_points(lo, hi, m, seed=101), threads)
        line["cpu_baseline"] = {"value": rate, "unit": "walks/s", "cores": threads, "kind": kind,
                                "sample": "%d of %d points x %d walks, %.1f s" % (m, n, cfg["solver"]["nWalks"], dt)}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
