"""SURVEY.md section 8(d) row M-BVH: the walk kernel on synthetic obstacles refined to 1e3 .. 1e5+ primitives (the meshes of
tests/golden/make_synthetic_scenes.py at higher resolutions, written to a scratch directory), default mode, for the three
big-mesh paths of csrc/wost_fast.cu.  The path is chosen by the environment (NMC_BIG_MESH = packet | tree | flat2, read once
per process), so run it once per path:

    for m in packet tree flat2; do NMC_BIG_MESH=$m python profiles/tools/mbvh_bench.py >> gpurun_out/mbvh.jsonl; done

usage: mbvh_bench.py [points] [sizes2d] [levels3d]     e.g.  mbvh_bench.py 20000 1024,16384,131072 3,5,6
One JSON line per mesh: primitives, tree depth, walks/s (CUDA events around the kernel, best of 3 after one warm-up)."""
import importlib.util
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
W = importlib.import_module(pkg.__name__ + ".workloads")
spec = importlib.util.spec_from_file_location("mss", os.path.join(ROOT, "tests", "golden", "make_synthetic_scenes.py"))
mss = importlib.util.module_from_spec(spec); spec.loader.exec_module(mss)

n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sizes2d = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else [1024, 16384, 131072]
levels3d = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 and sys.argv[3] else [3, 5, 6]
mode = os.environ.get("NMC_BIG_MESH", "default")
tmp = tempfile.mkdtemp(prefix="mbvh_")


def run(name, base, dim, verts, prims, tag):
    path = os.path.join(tmp, name + ".obj")
    mss.write_obj(path, "synthetic M-BVH mesh", verts, prims, tag)
    cfg = W.load_case(base)
    cfg["scene"]["boundary"] = path
    src = W.source_grid(base)
    t0 = time.time()
    sc = pkg.Scene(cfg["scene"], src, device=0)
    build_s = time.time() - t0
    lo, hi = sc.bbox()
    pts = W.random_points(lo, hi, n_pts, seed=1)
    best, st = None, None
    for rep in range(4):
        p, g, _, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=7 + rep)
        if rep > 0:
            best = st.kernel_ms if best is None else min(best, st.kernel_ms)
    print(json.dumps({"mesh": name, "dim": dim, "primitives": int(len(prims)), "big_mesh_path": mode, "points": n_pts,
                      "walks_started": int(st.walks_started), "kernel_ms": best, "walks_per_sec": st.walks_started/(best*1e-3),
                      "walk_steps_per_walk": st.walk_steps/max(st.walks_started, 1), "scene_build_s": round(build_s, 2),
                      "finite": bool(np.isfinite(p).all() and np.isfinite(g).all())}), flush=True)
    sc.close() if hasattr(sc, "close") else None


for n in sizes2d:
    v, e = mss.channel_circle(n_circle=n)
    run("channel_circle_%d" % n, "channel_circle", 2, v, e, "l")
for lv in levels3d:
    v, f = mss.box_sphere(level=lv)
    run("box_sphere_l%d" % lv, "box_sphere", 3, v, f, "f")
