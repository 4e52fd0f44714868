#!/usr/bin/env python
"""bench_step.py -- simulation steps per second (BASELINE.json metric, second half): one operator-split step of
the Taylor-Green configuration (examples/taylorgreen/run.sh: SIREN 6x64, batch 64^2, 512^2 pressure samples,
1002^2 divergence grid, nWalks 500) with K Adam iterations per fit, through the device-resident stepper, next to
the reference's arrangement of the same work: stock PyTorch ops for the fits (one loss.item() per iteration),
divergence grid to the host, the reference's CPU walk-on-stars (its rate comes from `bench.py --impl reference` on the same scene), gradients back.
`--watertight` keeps the scene as shipped (the reference then classifies every point outside and returns zeros);
default is the solver-active variant."""
import argparse
import json
import math
import os
import sys
import time
from importlib import import_module

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
util = ge.load_package().workloads  # scene fixtures + the karman obstacle


def _leave():
    """destroy_process_group() blocks while CUDA graphs holding captured NCCL collectives are alive: flush and exit."""
    torch.cuda.synchronize()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


def build(args, pkg, st):
    """(stepper, initial-velocity function, network shape, description) for the chosen example configuration."""
    dd = dict(device=getattr(args, "device", 0), distributed=getattr(args, "distributed", False), fit_parallel=getattr(args, "fit_parallel", "replicated"))
    if args.case == "taylorgreen":
        cfg = util.load_case("taylorgreen_shipped" if args.watertight else "taylorgreen_active")
        size = util.scene_size_from_obj(cfg["scene"]["boundary"])  # main.py:36-45: the samples live in the OBJ's box, not in [0, 2 pi]
        s = st.SplitStepper(cfg, scene_size=size, max_n_iters=args.iters, early_stop=False, use_cuda_graph=not args.no_graph, seed=1, **dd)
        tg = util.taylor_green_initial(size)
        return s, cfg, tg, (6, 64), "taylorgreen step (SIREN 6x64, batch 64^2, dt 1e-3), 512^2 pressure samples x 500 walks, 1002^2 divergence grid"
    if args.case in ("smoke3d", "karman3d"):
        # examples/smoke3d/run.sh (--src smoke): SIREN 5x64; examples/karman3d/run.sh: SIREN 2x128, karman_vel 0.5;
        # both: batch 128^2, dt 0.05, bdry_eps 1e-2, reset_wts 1, wost_resolution 256, 82^3 divergence grid
        cfg = util.load_case(args.case)
        smoke = args.case == "smoke3d"
        s = st.SplitStepper(cfg, scene_size=(-1.0, 1.0)*3, hidden_features=64 if smoke else 128, num_hidden_layers=5 if smoke else 2, dt=0.05,
                            lr=1e-5, sample_resolution=128, wost_resolution=256, grid_resolution=80, bdry_eps=1e-2, max_n_iters=args.iters,
                            early_stop=False, use_cuda_graph=not args.no_graph, boundary="smoke" if smoke else "karman3d",
                            obstacle=None if smoke else ((0.0, -0.8), 0.1), karman_vel=0.5, reset_wts=True, seed=1, **dd)
        if smoke:
            init = lambda x: torch.zeros_like(x)  # noqa: E731  (smoke_velocity, 3d sources.py:20-45: zero outside the inlet ball, which the envelope overrides)
        else:
            def init(x):  # karman_velocity, 3d sources.py:98-107: (0, 0, karman_vel) times the cylinder weight
                d = torch.sqrt(x[:, 0]**2 + (x[:, 2] + 0.8)**2) - 0.1
                w = torch.clamp(d, 0, 1e-2)/1e-2
                return torch.stack([torch.zeros_like(w), torch.zeros_like(w), 0.5*w], dim=-1)
        return s, cfg, init, ((5, 64) if smoke else (2, 128)), "%s step (SIREN %s 3->3, batch 128^2, dt 0.05, reset_wts), 256^2 pressure samples x 500 walks, 82^3 divergence grid" % (
            "smoke3d" if smoke else "karman3d", "5x64" if smoke else "2x128")
    if args.case == "smoke_obs":
        # examples/smoke_obs/run.sh: SIREN 5x64 3->3, batch 128^2, dt 0.05, bdry_eps 1e-2, reset_wts 1, wost_resolution 256, 82^3 grid
        cfg = util.load_case("smoke3d")
        s = st.SplitStepper(cfg, scene_size=(-1.0, 1.0)*3, hidden_features=64, num_hidden_layers=5, dt=0.05, lr=1e-5, sample_resolution=128,
                            wost_resolution=256, grid_resolution=80, bdry_eps=1e-2, max_n_iters=args.iters, early_stop=False,
                            use_cuda_graph=not args.no_graph, boundary="smoke_obs", obstacle=((0.0, 0.0, -0.3), 0.1), reset_wts=True, seed=1, **dd)
        rest = lambda x: torch.zeros_like(x)  # noqa: E731  (the smoke examples start from rest; the inlet ball drives the flow)
        return s, cfg, rest, (5, 64), "smoke_obs step (SIREN 5x64 3->3, batch 128^2, dt 0.05, reset_wts), 256^2 pressure samples x 500 walks, 82^3 divergence grid"
    # examples/karman/run.sh: SIREN 2x128, batch 128^2, dt 0.05, bdry_eps 3e-2, karman_vel 0.5, reset_wts 1, wost_resolution 512
    cfg = util.load_case("karman")
    centre, radius, size = util.karman_obstacle(cfg["output"]["boundaryDistanceMask"])
    s = st.SplitStepper(cfg, scene_size=size, hidden_features=128, num_hidden_layers=2, dt=0.05, lr=1e-5, sample_resolution=128,
                        wost_resolution=512, grid_resolution=1000, bdry_eps=3e-2, max_n_iters=args.iters, early_stop=False,
                        use_cuda_graph=not args.no_graph, boundary="karman", obstacle=(centre, radius), karman_vel=0.5, reset_wts=True, seed=1, **dd)
    return s, cfg, s.karman_initial_velocity, (2, 128), "karman step (SIREN 2x128, batch 128^2, dt 0.05, reset_wts), <= 512^2 pressure samples outside the cylinder x 500 walks, 401x1002 divergence grid"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="taylorgreen", choices=["taylorgreen", "karman", "smoke_obs", "smoke3d", "karman3d"])
    ap.add_argument("--iters", type=int, default=1000, help="Adam iterations per fit (reference: 10000)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--watertight", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--fit-parallel", dest="fit_parallel", default="replicated", choices=["replicated", "data"], help="multi-GPU fits: every rank the whole fit (default) or data-parallel with a gradient all_reduce per iteration")
    ap.add_argument("--gpus", type=int, default=1, help="under torchrun: data-parallel fits + sharded pressure solve")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    args.device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    args.distributed = world > 1
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(args.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.device))
    pkg = ge.load_package()
    st = import_module(pkg.__name__ + ".stepper")
    s, cfg, init_fn, (n_hidden, hidden), what = build(args, pkg, st)
    s.fit_initial(init_fn, 300, lr=1e-3)
    s.step(50)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    parts = {"advect_ms": 0.0, "pressure_ms": 0.0, "project_ms": 0.0}
    for _ in range(args.steps):
        a = time.perf_counter(); s._sync_prev(); s.advect_velocity(args.iters); torch.cuda.synchronize()
        b = time.perf_counter(); s._sync_prev()
        s.project_velocity(args.iters); torch.cuda.synchronize()   # divergence grid + wost + projection fit
        d = time.perf_counter()
        parts["advect_ms"] += 1e3*(b - a); parts["pressure_ms"] += s.last["pressure_ms"]; parts["project_ms"] += 1e3*(d - b) - s.last["pressure_ms"]
    total = time.perf_counter() - t0
    if world > 1:  # max over ranks; only rank 0 reports and times the reference arrangement
        import torch.distributed as dist
        t = torch.tensor([total] + [parts[k] for k in sorted(parts)], device=s.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t[0].item())
        for i, k in enumerate(sorted(parts)):
            parts[k] = float(t[1 + i].item())
        dist.barrier()
        if rank != 0:
            _leave()
    ours = args.steps/total

    # reference arrangement of the same step: stock torch ops + torch Adam + one loss.item() per iteration (base.py:142)
    S = pkg.load_siren()
    net = S.FusedSiren(s.dim, s.dim, n_hidden, hidden, nonlinearity="sine").cuda(); prev = S.FusedSiren(s.dim, s.dim, n_hidden, hidden, nonlinearity="sine").cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)
    nb = s.sample_resolution**2

    def ref_iter():
        x = s.sample_random(nb)
        with torch.no_grad():
            pu = s.apply_envelope_reference(x, prev.forward_reference(x))
            back = torch.clamp(x - pu*s.dt, min=s._lo, max=s._hi)
            adv = s.apply_envelope_reference(back, prev.forward_reference(back))
        loss = torch.mean((s.apply_envelope_reference(x, net.forward_reference(x)) - adv)**2)
        opt.zero_grad(); loss.backward(retain_graph=True); opt.step()
        return loss.item()
    for _ in range(20):
        ref_iter()
    torch.cuda.synchronize(); t = time.perf_counter()
    n_ref = min(args.iters, 300)
    for _ in range(n_ref):
        ref_iter()
    torch.cuda.synchronize(); ref_iter_ms = 1e3*(time.perf_counter() - t)/n_ref
    # CPU solver rate: the reference arm of bench.py (the one place that runs oracle/_ref) on this configuration's scene
    import subprocess
    n_press = int(s.last["pressure_samples"].shape[0])
    ref_case = {"taylorgreen": "taylorgreen_shipped" if args.watertight else "taylorgreen_active", "karman": "karman", "smoke_obs": "smoke3d", "smoke3d": "smoke3d", "karman3d": "karman3d"}[args.case]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--case", ref_case, "--points", str(args.cpu_sample),
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900)
    arm = json.loads(r.stdout.strip().splitlines()[-1])
    threads = arm["cpu_baseline"]["cores"]
    cpu_wost_s = float(s.last["walks"])/arm["value"] if arm["value"] > 0 else float("nan")
    ref_step_s = 2*args.iters*ref_iter_ms*1e-3 + cpu_wost_s
    print(json.dumps({"metric": "sim_steps_per_sec", "value": ours, "unit": "steps/s",
                      "config": {"workload": "%s, %d Adam iterations per fit" % (what, args.iters), "case": args.case,
                                 "scene": ("as shipped (isWatertight)" if args.watertight else "solver active (isWatertight:false)") if args.case == "taylorgreen" else "as shipped",
                                 "cuda_graph": not args.no_graph, "n_gpus": world,
                                 "parallelism": "replicated networks, data-parallel fits (1 gradient all_reduce per iteration), pressure samples sharded + all_gather" if world > 1 else "single GPU"},
                      "ms_per_step": 1e3*total/args.steps, "breakdown_ms_per_step": {k: v/args.steps for k, v in parts.items()},
                      "walks_per_step": int(s.last["walks"]), "pressure_samples": n_press, "wost_kernel_ms": s.last["wost_ms"],
                      "reference_arrangement": {"fit_iteration_ms_stock_torch": ref_iter_ms, "cpu_walks_per_sec": arm["value"], "cpu_wost_s": cpu_wost_s,
                                                "cores": threads, "step_s": ref_step_s, "steps_per_sec": 1.0/ref_step_s}}), flush=True)
    if world > 1:
        _leave()


if __name__ == "__main__":
    main()
