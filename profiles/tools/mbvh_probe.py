"""One default-mode solve on a refined synthetic obstacle, for ncu: mbvh_probe.py <2|3> <size or level> <points>."""
import importlib.util
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import importlib  # noqa: E402

pkg = ge.load_package()
W = importlib.import_module(pkg.__name__ + ".workloads")
spec = importlib.util.spec_from_file_location("mss", os.path.join(ROOT, "tests", "golden", "make_synthetic_scenes.py"))
mss = importlib.util.module_from_spec(spec); spec.loader.exec_module(mss)
dim, size, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
base = "channel_circle" if dim == 2 else "box_sphere"
v, p = mss.channel_circle(n_circle=size) if dim == 2 else mss.box_sphere(level=size)
path = os.path.join(tempfile.mkdtemp(), "m.obj")
mss.write_obj(path, "synthetic M-BVH mesh", v, p, "l" if dim == 2 else "f")
cfg = W.load_case(base); cfg["scene"]["boundary"] = path
sc = pkg.Scene(cfg["scene"], W.source_grid(base), device=0)
lo, hi = sc.bbox()
pts = W.random_points(lo, hi, n, seed=1)
for rep in range(2):
    _, _, _, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=7 + rep)
print(len(p), st.walks_started, st.kernel_ms)
