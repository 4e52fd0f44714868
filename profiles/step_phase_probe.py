#!/usr/bin/env python
"""Per-kernel time of the advect and project fit iterations of one configuration (torch.profiler, eager launches).
usage: step_phase_probe.py [taylorgreen|karman] [iterations]"""
import os
import sys
from importlib import import_module
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench_step  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "karman"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pkg = ge.load_package()
st = import_module(pkg.__name__ + ".stepper")
s, cfg, init_fn, _, what = bench_step.build(SimpleNamespace(case=case, iters=iters, watertight=False, no_graph=True), pkg, st)
s.fit_initial(init_fn, 50, lr=1e-3)
s.step(20)
for phase in ("advect", "project"):
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        s._sync_prev()
        (s.advect_velocity if phase == "advect" else s.project_velocity)(iters)
        torch.cuda.synchronize()
    print("==== %s, %s, %d iterations" % (case, phase, iters))
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
    for e in rows:
        print("%9.1f us total  %6d calls  %8.1f us/call  %s" % (e.device_time_total, e.count, e.device_time_total/max(e.count, 1), e.key[:110]))
