#!/usr/bin/env python
"""Timeline of the last graph-replayed fit iterations of a phase (torch.profiler / CUPTI): start, duration, gap to the previous
kernel end on the same stream, stream, name.  usage: iteration_timeline.py [case] [advect|project] [iterations]"""
import os, sys, json, tempfile
from importlib import import_module
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench_step  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "taylorgreen"
phase = sys.argv[2] if len(sys.argv) > 2 else "advect"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
pkg = ge.load_package()
st = import_module(pkg.__name__ + ".stepper")
s, cfg, init_fn, _, what = bench_step.build(SimpleNamespace(case=case, iters=iters, watertight=False, no_graph=False), pkg, st)
s.fit_initial(init_fn, 50, lr=1e-3)
s.step(20)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    s._sync_prev()
    (s.advect_velocity if phase == "advect" else s.project_velocity)(iters)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# last iterations: find the Adam kernels
adam = [i for i, e in enumerate(ev) if "adamKernelDev" in e["name"] or "adamFetchKernel" in e["name"]]   # the Adam launch closes an iteration
lo = adam[-4] + 1 if len(adam) >= 4 else 0
t0 = ev[lo]["ts"]
last_end = {}
print("%s %s: last 3 iterations (us)" % (case, phase))
for e in ev[lo:]:
    sid = e["args"].get("stream", 0)
    gap = e["ts"] - last_end.get(sid, e["ts"])
    last_end[sid] = e["ts"] + e["dur"]
    print("%8.1f  dur %6.1f  gap %5.1f  stream %3s  %s" % (e["ts"] - t0, e["dur"], gap, sid, e["name"][:70]))
if len(adam) >= 4:
    print("iteration period: %.1f us" % ((ev[adam[-1]]["ts"] - ev[adam[-4]]["ts"])/3))
