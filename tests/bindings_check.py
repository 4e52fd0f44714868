"""Exercises the compiled drop-in module `zombie_bindings` of one dimension in a fresh interpreter (the 2D and 3D
modules share their name, exactly like the reference's, so they cannot live in one process).
usage: bindings_check.py <2|3> <cpu|gpu>"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import util  # noqa: E402

dim, where = int(sys.argv[1]), sys.argv[2]
sys.path.insert(0, os.path.join(ROOT, "neural-monte-carlo-fluid-simulation_b200", "zombie%dd" % dim))  # what INTEGRATION.md tells a user to do
import zombie_bindings as zb  # noqa: E402

assert zb.__doc__ == "pybind11 WoSt"                       # demo.cpp:394
assert hasattr(zb, "wost") and hasattr(zb, "Scene")
assert hasattr(zb, "bvc") == (dim == 2)                     # zombie3d has no bvc (demo.cpp:119-125)
cfg = util.load_case("karman" if dim == 2 else "smoke3d")
src = util.source_grid(dim)


def raises(exc, fn):
    try:
        fn()
    except exc:
        return True
    except Exception as e:  # noqa: BLE001
        raise AssertionError("expected %s, got %r" % (exc.__name__, e))
    raise AssertionError("expected %s, nothing raised" % exc.__name__)


# error behaviour: raise where the reference abort()s / exit()s
bad = dict(cfg["scene"]); del bad["boundary"]
assert raises(KeyError, lambda: zb.Scene(bad, src))
assert raises(RuntimeError, lambda: zb.Scene(dict(cfg["scene"], boundary="/nonexistent.obj"), src))
assert raises(TypeError, lambda: zb.Scene(cfg["scene"], np.zeros((3,)*(dim + 1), np.float32)))
if dim == 2:  # Scene(config) (scene.h:22-52): the source grid comes from the image file config["sourceValue"]
    assert raises(KeyError, lambda: zb.Scene(cfg["scene"]))                                        # no "sourceValue"
    assert raises(RuntimeError, lambda: zb.Scene(dict(cfg["scene"], sourceValue="/nonexistent.pfm")))
    assert raises(RuntimeError, lambda: zb.Scene(dict(cfg["scene"], sourceValue="/tmp/source.png")))  # PFM only
else:
    assert raises(TypeError, lambda: zb.Scene(cfg["scene"]))                                       # zombie3d has no such overload
if where == "cpu":
    # no GPU here: creating a scene must fail loudly (no CPU fallback)
    assert raises(RuntimeError, lambda: zb.Scene(cfg["scene"], src))
    print("BINDINGS_OK cpu dim=%d" % dim)
    sys.exit(0)

scene = zb.Scene(cfg["scene"], src.tolist())                # nested lists, as src/*/models/model_split.py passes them
if dim == 2:
    # Scene(config): the same grid read back from a PFM file gives the same estimates as Scene(config, grid);
    # its defaults differ (isWatertight / flipOrientation true), so pass the two-argument form's explicitly
    import tempfile
    tmp = tempfile.mkdtemp()
    pfm = os.path.join(tmp, "source.pfm")
    with open(pfm, "wb") as f:
        f.write(b"Pf\n%d %d\n-1\n" % (src.shape[1], src.shape[0])); f.write(np.ascontiguousarray(src, "<f4").tobytes())
    scene1 = zb.Scene(dict(cfg["scene"], sourceValue=pfm, isWatertight=cfg["scene"].get("isWatertight", False),
                           flipOrientation=cfg["scene"].get("flipOrientation", False)))
    q = util.random_points(np.array([-1.0, -0.5], np.float32), np.array([1.8, 0.5], np.float32), 64, seed=4)
    zb.set_mode("deterministic"); zb.set_seed(3)
    pa1, _ = zb.wost_array(scene1, cfg["solver"], cfg["output"], q)
    pa2, _ = zb.wost_array(scene, cfg["solver"], cfg["output"], q)
    assert np.allclose(pa1, pa2, rtol=1e-5, atol=1e-9) and np.abs(pa2).max() > 0     # the grey conversion rounds 0.299 + 0.587 + 0.114
    # bvc(scene, solverConfig, outputConfig) -> None, writes the solution image and its colour-mapped copy (demo.cpp:265-363)
    bsolver = dict(cfg["solver"], boundaryCacheSize=512, domainCacheSize=512, nWalksForCachedSolutionEstimates=16)
    bout = dict(cfg["output"], gridRes=32, solutionFile=os.path.join(tmp, "out", "bvc.pfm"), colormap="turbo", colormapMinVal=-1e-3, colormapMaxVal=1e-3)
    assert zb.bvc(scene, bsolver, bout) is None
    assert os.path.getsize(os.path.join(tmp, "out", "bvc.pfm")) == len(b"PF\n32 32\n-1\n") + 32*32*3*4
    assert os.path.exists(os.path.join(tmp, "out", "bvc_color.pfm"))
    assert zb.bvc(scene, bsolver, dict(bout, solutionFile=os.path.join(tmp, "bvc.png"))) is None
    with open(os.path.join(tmp, "bvc_color.png"), "rb") as f:
        assert f.read(8) == b"\x89PNG\r\n\x1a\n"
    grid = zb.bvc_grid(scene, bsolver, bout)
    assert grid.shape == (32, 32) and np.isfinite(grid).all() and np.abs(grid).max() > 0
    zb.set_seed(11)
no_grid = dict(cfg["output"]); del no_grid["gridRes"]
pts = util.random_points(*[np.asarray(v, np.float32) for v in ((-0.9,)*dim, (0.9,)*dim)], 300, seed=3) if dim == 3 else \
    util.random_points(np.array([-1.0, -0.5], np.float32), np.array([1.8, 0.5], np.float32), 300, seed=3)
assert raises(KeyError, lambda: zb.wost(scene, cfg["solver"], no_grid, pts.tolist()))   # gridRes is required (demo.cpp:132)
assert raises(TypeError, lambda: zb.wost(scene, cfg["solver"], cfg["output"], np.zeros((5, dim + 1), np.float32)))
zb.set_mode("deterministic"); zb.set_seed(11)
out_pts, p, g = zb.wost(scene, cfg["solver"], cfg["output"], pts.tolist())
assert isinstance(out_pts, list) and isinstance(p, list) and isinstance(g, list) and isinstance(g[0], list)
assert len(p) == 300 and len(g) == 300 and len(g[0]) == dim and len(out_pts[0]) == dim
assert np.allclose(np.array(out_pts, np.float32), pts)
pa, ga = zb.wost_array(scene, cfg["solver"], cfg["output"], pts)
assert np.array_equal(np.array(p, np.float32), pa) and np.array_equal(np.array(g, np.float32), ga)
st = zb.last_stats()
assert st["walks_started"] > 0 and st["kernel_launches"] >= 1
# the same numbers as the Python mirror of the module (tests / bench use that one)
pkg = util.package()
sc = pkg.Scene(cfg["scene"], src, device=0)
pm, gm, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_DETERMINISTIC, seed=11)
assert np.array_equal(pa, pm) and np.array_equal(ga, gm)
zb.set_mode("fast"); zb.set_seed(None)
_, p2, g2 = zb.wost(scene, cfg["solver"], cfg["output"], pts)
_, p3, g3 = zb.wost(scene, cfg["solver"], cfg["output"], pts)
p2, p3 = np.array(p2), np.array(p3)
assert np.isfinite(p2).all() and not np.array_equal(p2, p3)            # unseeded: a fresh seed per call, like the reference's clock
zb.set_seed(5)                                                         # seeded default mode: the same estimate up to Monte Carlo noise
_, p4, _ = zb.wost(scene, cfg["solver"], cfg["output"], pts)
p4 = np.array(p4)
assert np.abs(p4 - pa).mean() < 0.05*np.abs(pa).mean() + 1e-7, (np.abs(p4 - pa).mean(), np.abs(pa).mean())
assert raises(ValueError, lambda: zb.set_mode("bogus"))
print("BINDINGS_OK gpu dim=%d" % dim, json.dumps({k: int(v) if isinstance(v, (int, np.integer)) else float(v) for k, v in st.items()}))
