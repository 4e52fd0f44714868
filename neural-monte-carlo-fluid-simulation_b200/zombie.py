"""Python mirror of the reference's `zombie_bindings` module surface on top of the C ABI.

Reference interface (same names, argument meaning and result types):
  bindings/zombie/demo/demo.cpp:393-401    m.def("wost"), m.def("bvc"), Scene(config), Scene(config, sourceValue)
  bindings/zombie3d/demo/demo.cpp:119-125  m.def("wost"), Scene(config, sourceValue)
  callers: src/2d/models/model_split.py:185-228, src/3d/models/model_split.py:190-230

The compiled drop-ins (zombie2d/zombie_bindings*.so, zombie3d/zombie_bindings*.so, csrc/zombie_bindings.cpp)
are what src/2d and src/3d import; this module offers the same calls to Python code that does not want
two same-named extension modules in one process (tests, bench.py), plus additive array entry points.
Errors raise (KeyError / RuntimeError) where the reference abort()s (demo/config.h:6-11).
"""
import os

import numpy as np

from . import capi

_DEFAULTS = {"mode": capi.MODE_FAST, "seed": None, "device": None}


def set_defaults(mode=None, seed=None, device=None):
    """Process-wide defaults for the arguments the reference API has no slot for."""
    if mode is not None:
        _DEFAULTS["mode"] = mode
    if seed is not None:
        _DEFAULTS["seed"] = seed
    if device is not None:
        _DEFAULTS["device"] = device


def _required(cfg, key):
    if key not in cfg:
        raise KeyError("Missing required setting: %s" % key)  # reference: abort(), demo/config.h:6-11
    return cfg[key]


def load_obj(path, dim, flip_orientation=False):
    """2D: `v x y`, `l i j` (demo/scene.h:104-145); 3D: `v x y z`, `f a[/t[/n]] b c`
    (fcpw/utilities/scene_loader.inl:99-150). Returns (verts[V,dim] f32, prims[P,dim] i32)."""
    if not os.path.exists(path):
        raise RuntimeError("Error opening file: %s" % path)
    verts, prims = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(x) for x in t[1:1 + dim]])
            elif t[0] == "l" and dim == 2:
                i, j = int(t[1]) - 1, int(t[2]) - 1
                prims.append([j, i] if flip_orientation else [i, j])
            elif t[0] == "f" and dim == 3:
                for tok in t[1:]:
                    i = int(tok.split("/")[0])
                    prims.append(i - 1 if i > 0 else len(verts) + i)
    v = np.asarray(verts, np.float32).reshape(-1, dim)
    p = np.asarray(prims, np.int32).reshape(-1)       # 2D rows are pairs already, 3D indices arrive one by one
    p = p[: (len(p) // dim) * dim].reshape(-1, dim)   # a trailing incomplete face is dropped
    return v, p


def normalize_domain(v):
    """scene.h:132-142 with the reference's float arithmetic: the centre of mass is accumulated vertex by vertex
    (numpy's pairwise sum would round differently), the radius is the largest sqrt(x*x + y*y), then a true division."""
    f = np.float32
    cx, cy = f(0), f(0)
    for x, y in v:
        cx = f(cx + x); cy = f(cy + y)
    n = f(len(v))
    cm = np.array([f(cx/n), f(cy/n)], f)
    w = (v - cm).astype(f)
    radius = f(0)
    for x, y in w:
        radius = max(radius, f(np.sqrt(f(f(x*x) + f(y*y)))))
    return (w/radius).astype(f)


def solver_opts(solver, output, mode=None, seed=None):
    """demo.cpp:121-137 (both dimensions): defaults and the misspelt `setps...` keys are API."""
    o = capi.SolverOpts()
    o.nWalks = int(solver.get("nWalks", 128))
    o.maxWalkLength = int(solver.get("maxWalkLength", 1024))
    o.stepsBeforeApplyingTikhonov = int(solver.get("setpsBeforeApplyingTikhonov", o.maxWalkLength))
    o.stepsBeforeUsingMaximalSpheres = int(solver.get("setpsBeforeUsingMaximalSpheres", o.maxWalkLength))
    o.epsilonShell = float(solver.get("epsilonShell", 1e-3))
    o.minStarRadius = float(solver.get("minStarRadius", 1e-3))
    o.silhouettePrecision = float(solver.get("silhouettePrecision", 1e-3))
    o.russianRouletteThreshold = float(solver.get("russianRouletteThreshold", 0.0))
    o.useGradientControlVariates = int(not solver.get("disableGradientControlVariates", False))
    o.useGradientAntitheticVariates = int(not solver.get("disableGradientAntitheticVariates", False))
    o.useCosineSamplingForDerivatives = int(bool(solver.get("useCosineSamplingForDirectionalDerivatives", False)))
    o.ignoreDirichlet = int(bool(solver.get("ignoreDirichlet", False)))
    o.ignoreNeumann = int(bool(solver.get("ignoreNeumann", False)))
    o.ignoreSource = int(bool(solver.get("ignoreSource", False)))
    _required(output, "gridRes")  # required although unused, demo.cpp:132
    o.boundaryDistanceMask = float(output.get("boundaryDistanceMask", 0.0))
    o.mode = _DEFAULTS["mode"] if mode is None else mode
    if seed is None:
        seed = _DEFAULTS["seed"]
    if seed is None:  # the reference seeds from the wall clock (walk_on_stars.h:498,639)
        seed = int.from_bytes(os.urandom(8), "little")
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return o


class Scene:
    """Scene(config, sourceValue): boundary OBJ + source grid (demo/scene.h:54-77, scene_3d.h:22-40).

    2D sourceValue[i][j]: i <-> y (rows), j <-> x; 3D sourceValue[i][j][k] <-> (x, y, z)."""

    def __init__(self, config, sourceValue=None, device=None):
        if sourceValue is None:
            # 1-argument constructor of the 2D module reads an image file (scene.h:22-52); src/ never uses it
            raise NotImplementedError("Scene(config) with an image-file source is not provided; pass sourceValue")
        src = np.ascontiguousarray(sourceValue, dtype=np.float32)
        if src.ndim not in (2, 3):
            raise TypeError("sourceValue must be a 2-D (zombie) or 3-D (zombie3d) array")
        self.dim = src.ndim
        boundary = _required(config, "boundary")
        self.isWatertight = bool(config.get("isWatertight", False))
        self.isDoubleSided = bool(config.get("isDoubleSided", False))
        flip = bool(config.get("flipOrientation", False)) if self.dim == 2 else False  # 3D ignores both flags
        v, p = load_obj(boundary, self.dim, flip)
        if self.dim == 2 and config.get("normalizeDomain", False):
            v = normalize_domain(v)
        dev = _DEFAULTS["device"] if device is None else device
        if dev is None:
            dev = int(os.environ.get("LOCAL_RANK", "0")) if capi.device_count() > 1 else 0
        self.handle = capi.SceneHandle(self.dim, v, p, src, float(config.get("absorptionCoeff", 0.0)),
                                       self.isWatertight, self.isDoubleSided, dev)

    def bbox(self):
        return self.handle.bbox()


def wost_array(scene, solverConfig, outputConfig, sample_points, mode=None, seed=None, index_offset=0, want_stats=False):
    """Additive fast entry point: numpy in, numpy out. Returns (p[N], grad[N,dim], stats12|None, SolveStats)."""
    opts = solver_opts(solverConfig, outputConfig, mode, seed)
    return scene.handle.solve(opts, sample_points, index_offset, want_stats)


def wost(scene, solverConfig, outputConfig, sample_points):
    """wost(scene, solverConfig, outputConfig, sample_points) -> (sample_points, solution, gradient) as nested
    Python lists, exactly what the pybind STL casters return (demo.cpp:119,204; zombie3d demo.cpp:15,116)."""
    pts = np.ascontiguousarray(sample_points, dtype=np.float32).reshape(-1, scene.dim)
    p, g, _, _ = wost_array(scene, solverConfig, outputConfig, pts)
    return pts.tolist(), p.tolist(), g.tolist()


def bvc(scene, solverConfig, outputConfig):
    """Boundary value caching (2D module only, demo.cpp:265-363); never called from src/ (SURVEY.md section 8f)."""
    raise NotImplementedError("bvc is outside the pressure-projection hot path and is not provided yet")
