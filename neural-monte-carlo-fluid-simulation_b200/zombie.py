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
        one_arg = sourceValue is None
        if one_arg:
            # Scene(config) of the 2D module (scene.h:22-52): the source grid is the image file config["sourceValue"];
            # isWatertight and flipOrientation default to True in this form.  PFM files only.
            sourceValue = read_pfm_grey(_required(config, "sourceValue"))
        src = np.ascontiguousarray(sourceValue, dtype=np.float32)
        if src.ndim not in (2, 3):
            raise TypeError("sourceValue must be a 2-D (zombie) or 3-D (zombie3d) array")
        self.dim = src.ndim
        boundary = _required(config, "boundary")
        self.isWatertight = bool(config.get("isWatertight", one_arg))
        self.isDoubleSided = bool(config.get("isDoubleSided", False))
        flip = bool(config.get("flipOrientation", one_arg)) if self.dim == 2 else False  # 3D ignores both flags
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


def read_pfm_grey(path):
    """Image<1>::readPFM (demo/image.h:104-148): rows in file order (no flip on read); colour files are reduced to
    0.299 r + 0.587 g + 0.114 b (evaluated in double, as the reference's literals are), grey files go through the same sum."""
    if not str(path).lower().endswith(".pfm"):
        raise RuntimeError("Scene(config): only PFM source images are supported (%s)" % path)
    if not os.path.exists(path):
        raise RuntimeError("Error opening file: %s" % path)
    with open(path, "rb") as f:
        kind = f.readline().strip()
        if kind not in (b"PF", b"Pf"):
            raise RuntimeError("Invalid PFM file detected while reading %s" % path)
        w, h = [int(t) for t in f.readline().split()]
        scale = float(f.readline())
        nc = 3 if kind == b"PF" else 1
        data = np.frombuffer(f.read(4*w*h*nc), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, nc).astype(np.float64)
    r, g, b = data[..., 0], data[..., nc // 2 if nc == 3 else 0], data[..., nc - 1]
    return (0.299*r + 0.587*g + 0.114*b).astype(np.float32)


def bvc_opts(solver, output, epsilon_shell):
    """The solver options bvc() reads beyond solver_opts (demo.cpp:274-292)."""
    b = capi.BvcOpts()
    b.boundaryCacheSize = int(solver.get("boundaryCacheSize", 1024))
    b.domainCacheSize = int(solver.get("domainCacheSize", 1024))
    b.nWalksForCachedSolutionEstimates = int(solver.get("nWalksForCachedSolutionEstimates", 128))
    b.nWalksForCachedGradientEstimates = int(solver.get("nWalksForCachedGradientEstimates", 640))
    b.gridRes = int(_required(output, "gridRes"))
    b.normalOffsetForCachedDirichletSamples = float(solver.get("normalOffsetForCachedDirichletSamples", 5.0*epsilon_shell))
    b.radiusClampForKernels = float(solver.get("radiusClampForKernels", 1e-3))
    b.regularizationForKernels = float(solver.get("regularizationForKernels", 0.0))
    return b


def bvc_grid(scene, solverConfig, outputConfig, seed=None, want_cache=False):
    """Additive: the masked evaluation grid of bvc() as an array [i][j] (point (i, j) of createEvaluationGrid, grid.h:352-368)
    instead of image files; with want_cache also the boundary cache points (x, y, nx, ny, solution, pdf)."""
    if scene.dim != 2:
        raise RuntimeError("bvc exists in the 2D module only (bindings/zombie/demo/demo.cpp:393-401)")
    opts = solver_opts(solverConfig, outputConfig, capi.MODE_DETERMINISTIC, seed)
    grid, cache, n_domain = scene.handle.bvc_solve(opts, bvc_opts(solverConfig, outputConfig, opts.epsilonShell))
    return (grid, cache, n_domain) if want_cache else grid


def _write_pfm3(path, img):
    with open(path, "wb") as f:  # Image<3>::writePFM (image.h:173-198): rows flipped
        f.write(b"PF\n%d %d\n-1\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img[::-1], dtype="<f4").tobytes())


def _write_png3(path, img):
    import struct
    import zlib
    h, w, _ = img.shape  # Image<3>::writePNG (image.h:200-215): clamp(int(v * 255), 0, 255)
    raw = np.clip((img*np.float32(255.0)).astype(np.int64), 0, 255).astype(np.uint8)
    rows = b"".join(b"\x00" + raw[i].tobytes() for i in range(h))

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b""))


_TURBO = None


def _turbo(v):
    """The knot table of csrc/image_io.h applyColormap (every 8th entry of the public Turbo look-up table)."""
    global _TURBO
    if _TURBO is None:
        _TURBO = np.array([[0.190, 0.072, 0.232], [0.225, 0.164, 0.451], [0.251, 0.252, 0.634], [0.268, 0.338, 0.780], [0.276, 0.421, 0.891],
                           [0.275, 0.501, 0.966], [0.259, 0.580, 0.999], [0.214, 0.659, 0.980], [0.158, 0.736, 0.923], [0.112, 0.806, 0.845],
                           [0.093, 0.866, 0.762], [0.120, 0.912, 0.687], [0.197, 0.949, 0.595], [0.305, 0.977, 0.490], [0.428, 0.994, 0.386],
                           [0.547, 0.999, 0.296], [0.644, 0.990, 0.234], [0.726, 0.965, 0.206], [0.805, 0.925, 0.205], [0.875, 0.873, 0.216],
                           [0.933, 0.812, 0.227], [0.973, 0.747, 0.225], [0.993, 0.674, 0.203], [0.996, 0.587, 0.169], [0.984, 0.493, 0.128],
                           [0.958, 0.400, 0.088], [0.921, 0.315, 0.055], [0.874, 0.245, 0.033], [0.816, 0.185, 0.018], [0.746, 0.131, 0.009],
                           [0.664, 0.084, 0.004], [0.571, 0.045, 0.005], [0.480, 0.016, 0.011]], np.float32)
    idx = (np.clip(v, 0.0, 1.0)*255).astype(np.int64)
    knots = np.array(list(range(0, 256, 8)) + [255])
    return np.stack([np.interp(idx, knots, _TURBO[:, c]) for c in range(3)], axis=-1).astype(np.float32)


def bvc(scene, solverConfig, outputConfig):
    """bvc(scene, solverConfig, outputConfig) -> None (2D module only, demo.cpp:265-363): boundary value caching on the
    device (csrc/bvc.cu), then the files of saveEvaluationGrid / writeSolution (demo/grid.h:9-33, 370-415):
    outputConfig["solutionFile"] (default "solution.pfm") and <stem>_color<ext> unless saveColormapped is false."""
    grid = bvc_grid(scene, solverConfig, outputConfig)
    img = np.repeat(grid.T[:, :, None], 3, axis=2)  # solution->get(j, i): image row <-> y index, column <-> x index
    path = outputConfig.get("solutionFile", "solution.pfm")
    if os.path.dirname(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
    write = _write_pfm3 if path.lower().endswith(".pfm") else _write_png3
    write(path, img)
    if outputConfig.get("saveColormapped", True):
        lo, hi = float(outputConfig.get("colormapMinVal", 0.0)), float(outputConfig.get("colormapMaxVal", 1.0))
        val = np.clip((grid.T - np.float32(lo))/np.float32(hi - lo), 0.0, 1.0)
        col = _turbo(val) if outputConfig.get("colormap", "") == "turbo" else np.repeat(val[:, :, None], 3, axis=2)
        stem, ext = os.path.splitext(path)
        write(stem + "_color" + ext, col.astype(np.float32))
    return None


def estimate_solution(scene, solverConfig, outputConfig, points, n_walks, normals=None, types=None, aligned=None, seed=None, index_offset=0):
    """Additive: EstimationQuantity::Solution (walk_on_stars.h:354-461) at caller-given points, optionally starting ON the
    reflecting boundary (types[i] == 2 with normals[i]); deterministic replay.  Returns (solution[N], stats[N, 4])."""
    opts = solver_opts(solverConfig, outputConfig, capi.MODE_DETERMINISTIC, seed)
    return scene.handle.estimate_solution(opts, points, n_walks, normals, types, aligned, index_offset)
