"""Diagnostic: per-point statistics of the default mode and of the reference (oracle/_ref) on one workload, dumped to
gpurun_out/ for offline analysis.  usage: python profiles/tools/dump_fast_vs_ref.py CASE NPOINTS [solver overrides as k=v]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import refbind  # noqa: E402

case, n = sys.argv[1], int(sys.argv[2])
pkg = ge.load_package()
wl = pkg.workloads
cfg = wl.load_case(case)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    cfg["solver"][k] = json.loads(v)
dim = cfg["dim"]
src = wl.source_grid(case)
ref = refbind.RefScene(dim, cfg["scene"], src)
lo, hi = ref.bbox()
pts = wl.random_points(lo, hi, n, seed=31)
rp, rg, rst = ref.wost(cfg["solver"], cfg["output"], pts, seed=11, nthreads=os.cpu_count() or 4, want_stats=True)
rp2, rg2, rst2 = ref.wost(cfg["solver"], cfg["output"], pts, seed=12, nthreads=os.cpu_count() or 4, want_stats=True)
sc = pkg.Scene(cfg["scene"], src, device=0)
out = {"pts": pts, "ref": rst, "ref2": rst2}
for seed in (1, 2):
    p, g, s, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=seed, want_stats=True)
    out["fast%d" % seed] = s
p, g, s, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_DETERMINISTIC, seed=11, want_stats=True)
out["det"] = s
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = "_".join([case, str(n)] + [a.replace("=", "-") for a in sys.argv[3:]])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % tag), **out)
act = rst[:, 11] > 0
for k in ("ref2", "fast1", "fast2", "det"):
    a = out[k]
    b = act & (a[:, 11] > 0)
    print(k, "mean len %.5f vs ref %.5f | completed %.5f vs %.5f" % (a[b, 10].mean(), rst[b, 10].mean(), a[b, 9].mean(), rst[b, 9].mean()))
