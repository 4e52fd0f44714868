"""ctypes view of oracle/libnmc_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE.

Same method names as oracle/refbind.py so tests can run one body against both.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32
_LIB = os.path.join(_HERE, "libnmc_oracle.so")


class SolverOpts(C.Structure):
    _fields_ = [("nWalks", C.c_int), ("maxWalkLength", C.c_int),
                ("stepsBeforeApplyingTikhonov", C.c_int), ("stepsBeforeUsingMaximalSpheres", C.c_int),
                ("epsilonShell", C.c_float), ("minStarRadius", C.c_float),
                ("silhouettePrecision", C.c_float), ("russianRouletteThreshold", C.c_float),
                ("useGradientControlVariates", C.c_int), ("useGradientAntitheticVariates", C.c_int),
                ("useCosineSamplingForDerivatives", C.c_int), ("ignoreDirichlet", C.c_int),
                ("ignoreNeumann", C.c_int), ("ignoreSource", C.c_int),
                ("boundaryDistanceMask", C.c_float)]


def solver_opts(solver, output):
    """Same defaults and (misspelt) key names as bindings/zombie/demo/demo.cpp:121-142."""
    o = SolverOpts()
    o.nWalks = int(solver.get("nWalks", 128))
    o.maxWalkLength = int(solver.get("maxWalkLength", 1024))
    o.stepsBeforeApplyingTikhonov = int(solver.get("setpsBeforeApplyingTikhonov", o.maxWalkLength))
    o.stepsBeforeUsingMaximalSpheres = int(solver.get("setpsBeforeUsingMaximalSpheres", o.maxWalkLength))
    o.epsilonShell = solver.get("epsilonShell", 1e-3)
    o.minStarRadius = solver.get("minStarRadius", 1e-3)
    o.silhouettePrecision = solver.get("silhouettePrecision", 1e-3)
    o.russianRouletteThreshold = solver.get("russianRouletteThreshold", 0.0)
    o.useGradientControlVariates = int(not solver.get("disableGradientControlVariates", False))
    o.useGradientAntitheticVariates = int(not solver.get("disableGradientAntitheticVariates", False))
    o.useCosineSamplingForDerivatives = int(solver.get("useCosineSamplingForDirectionalDerivatives", False))
    o.ignoreDirichlet = int(solver.get("ignoreDirichlet", False))
    o.ignoreNeumann = int(solver.get("ignoreNeumann", False))
    o.ignoreSource = int(solver.get("ignoreSource", False))
    if "gridRes" not in output:
        raise KeyError("Missing required setting: gridRes")  # demo.cpp:132 aborts
    o.boundaryDistanceMask = output.get("boundaryDistanceMask", 0.0)
    return o


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def available():
    return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(_LIB)
        L.nmo_scene_create.restype = C.c_void_p
        L.nmo_scene_create.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_int,
                                       C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_int, C.c_int]
        L.nmo_scene_destroy.argtypes = [C.c_void_p]
        L.nmo_scene_bbox.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.nmo_scene_num_nodes.argtypes = [C.c_void_p]
        L.nmo_scene_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.nmo_wost.restype = C.c_int
        L.nmo_wost.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.POINTER(C.c_float), C.c_int,
                               C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float),
                               C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.nmo_point_seed.restype = C.c_uint64
        L.nmo_point_seed.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def load_obj(path, dim, flip_orientation=False):
    """2D: `v x y`, `l i j` (demo/scene.h:104-145). 3D: `v x y z`, `f a[/..] b c`
    (fcpw/utilities/scene_loader.inl:99-150; polygons are taken as listed, three indices each)."""
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
                idx = [int(tok.split("/")[0]) for tok in t[1:]]
                idx = [(i - 1) if i > 0 else (len(verts) + i) for i in idx]
                prims.extend(idx)
    v = np.asarray(verts, _f32).reshape(-1, dim)
    p = np.asarray(prims, np.int32).reshape(-1, dim)
    return v, p


class OracleScene:
    def __init__(self, dim, config, source):
        self.dim = dim
        self.L = lib()
        # constructor defaults of the 2-argument Scene (scene.h:54-63, scene_3d.h:22-31)
        flip = bool(config.get("flipOrientation", False)) if dim == 2 else False
        v, p = load_obj(config["boundary"], dim, flip)
        if dim == 2 and config.get("normalizeDomain", False):  # scene.h:132-142
            # the reference accumulates the centre of mass vertex by vertex in float and divides by the largest norm
            cx, cy = _f32(0), _f32(0)
            for x, y in v:
                cx = _f32(cx + x); cy = _f32(cy + y)
            cm = np.array([_f32(cx / _f32(len(v))), _f32(cy / _f32(len(v)))], _f32)
            v = (v - cm).astype(_f32)
            radius = _f32(0)
            for x, y in v:
                radius = max(radius, _f32(np.sqrt(_f32(_f32(x * x) + _f32(y * y)))))
            v = (v / radius).astype(_f32)
        self.verts, self.prims = np.ascontiguousarray(v), np.ascontiguousarray(p)
        src = np.ascontiguousarray(source, dtype=_f32)
        assert src.ndim == dim
        shp = list(src.shape) + [1] * (3 - dim)
        self._src = src
        self.h = self.L.nmo_scene_create(dim, _fp(self.verts), len(self.verts), _ip(self.prims), len(self.prims),
                                         _fp(src), shp[0], shp[1], shp[2],
                                         C.c_float(config.get("absorptionCoeff", 0.0)),
                                         int(bool(config.get("isWatertight", False))),
                                         int(bool(config.get("isDoubleSided", False))))

    def close(self):
        if self.h:
            self.L.nmo_scene_destroy(self.h)
            self.h = None

    def bbox(self):
        out = np.zeros(2 * self.dim, _f32)
        self.L.nmo_scene_bbox(self.h, _fp(out))
        return out[: self.dim].copy(), out[self.dim:].copy()

    def nodes(self):
        n = self.L.nmo_scene_num_nodes(self.h)
        out = np.zeros((n, 16), _f32)
        self.L.nmo_scene_nodes(self.h, _fp(out))
        return out

    def wost(self, solver, output, pts, seed=0, index_offset=0, nthreads=1, want_stats=False):
        pts = np.ascontiguousarray(pts, dtype=_f32).reshape(-1, self.dim)
        n = pts.shape[0]
        p = np.zeros(n, _f32)
        g = np.zeros((n, self.dim), _f32)
        st = np.zeros((n, 12), _f32) if want_stats else None
        o = solver_opts(solver, output)
        rc = self.L.nmo_wost(self.h, C.byref(o), _fp(pts), n, C.c_uint64(seed), C.c_uint64(index_offset), nthreads,
                             _fp(p), _fp(g), _fp(st) if want_stats else None)
        if rc != 0:
            raise RuntimeError("nmo_wost failed: %d" % rc)
        return p, g, st

    def _pts(self, pts):
        return np.ascontiguousarray(pts, dtype=_f32).reshape(-1, self.dim)

    def dist_neumann(self, pts, signed=False):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.nmo_dist_neumann(C.c_void_p(self.h), _fp(pts), len(pts), int(signed), _fp(out))
        return out

    def dist_dirichlet(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.nmo_dist_dirichlet(C.c_void_p(self.h), _fp(pts), len(pts), _fp(out))
        return out

    def inside_domain(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), np.int32)
        self.L.nmo_inside_domain(C.c_void_p(self.h), _fp(pts), len(pts), _ip(out))
        return out

    def outside_bbox(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), np.int32)
        self.L.nmo_outside_bbox(C.c_void_p(self.h), _fp(pts), len(pts), _ip(out))
        return out

    def star_radius(self, pts, min_r, max_r, prec=1e-3, flip=False):
        pts = self._pts(pts)
        mr = np.ascontiguousarray(np.broadcast_to(np.asarray(max_r, _f32), (len(pts),)))
        out = np.zeros(len(pts), _f32)
        self.L.nmo_star_radius(C.c_void_p(self.h), _fp(pts), len(pts), C.c_float(min_r), _fp(mr),
                               C.c_float(prec), int(flip), _fp(out))
        return out

    def intersect_neumann(self, org, nrm, dirs, tmax, onb):
        org, nrm, dirs = self._pts(org), self._pts(nrm), self._pts(dirs)
        n = len(org)
        tm = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, _f32), (n,)))
        ob = np.ascontiguousarray(np.broadcast_to(np.asarray(onb, np.int32), (n,)))
        out = np.zeros((n, 2 + 2 * self.dim), _f32)
        self.L.nmo_intersect_neumann(C.c_void_p(self.h), _fp(org), _fp(nrm), _fp(dirs), _fp(tm), _ip(ob), n, _fp(out))
        return out

    def blocked(self, xi, xj, ni, nj, offi, offj):
        xi, xj, ni, nj = self._pts(xi), self._pts(xj), self._pts(ni), self._pts(nj)
        n = len(xi)
        oi = np.ascontiguousarray(np.broadcast_to(np.asarray(offi, np.int32), (n,)))
        oj = np.ascontiguousarray(np.broadcast_to(np.asarray(offj, np.int32), (n,)))
        out = np.zeros(n, np.int32)
        self.L.nmo_blocked(C.c_void_p(self.h), _fp(xi), _fp(xj), _fp(ni), _fp(nj), _ip(oi), _ip(oj), n, _ip(out))
        return out

    def source(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.nmo_source(C.c_void_p(self.h), _fp(pts), len(pts), _fp(out))
        return out


# ---- scene-free probes (dim kept for signature parity with refbind) -----------------------------
def pcg32_uint(dim, initstate, initseq, n):
    s = (C.c_uint64 * 2)()
    L = lib()
    L.nmo_pcg32_seed(s, C.c_uint64(initstate), C.c_uint64(initseq))
    L.nmo_pcg32_uint.restype = C.c_uint32
    return np.array([L.nmo_pcg32_uint(s) for _ in range(n)], np.uint32)


def pcg32_float(dim, initstate, initseq, n):
    s = (C.c_uint64 * 2)()
    L = lib()
    L.nmo_pcg32_seed(s, C.c_uint64(initstate), C.c_uint64(initseq))
    L.nmo_pcg32_float.restype = C.c_float
    return np.array([L.nmo_pcg32_float(s) for _ in range(n)], _f32)


def pcg32_bounded(dim, initstate, initseq, bounds):
    s = (C.c_uint64 * 2)()
    L = lib()
    L.nmo_pcg32_seed(s, C.c_uint64(initstate), C.c_uint64(initseq))
    L.nmo_pcg32_bounded.restype = C.c_uint32
    return np.array([L.nmo_pcg32_bounded(s, C.c_uint32(int(b))) for b in bounds], np.uint32)


def point_seed(dim, seed, index):
    return int(lib().nmo_point_seed(C.c_uint64(seed), C.c_uint64(index)))


def stratified(dim, initstate, n_samples):
    out = np.zeros((dim - 1) * n_samples, _f32)
    st = np.zeros(2, np.uint64)
    lib().nmo_stratified(dim, C.c_uint64(initstate), n_samples, _fp(out), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out, st


def sphere_dir(dim, u):
    u = np.ascontiguousarray(u, _f32).reshape(-1, dim - 1)
    out = np.zeros((len(u), dim), _f32)
    lib().nmo_sphere_dir(dim, _fp(u), len(u), _fp(out))
    return out


def bessel(dim, kind, x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    lib().nmo_bessel(kind, x.ctypes.data_as(C.POINTER(C.c_double)), x.size, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def greens_ball(dim, lam, R, r):
    R = np.ascontiguousarray(R, _f32)
    r = np.ascontiguousarray(r, _f32)
    out = np.zeros((len(R), 10), _f32)
    lib().nmo_greens_ball(dim, C.c_float(lam), _fp(R), _fp(r), len(R), _fp(out))
    return out


def sample_volume(dim, lam, R, seeds):
    R = np.ascontiguousarray(R, _f32)
    seeds = np.ascontiguousarray(seeds, np.uint64)
    r = np.zeros(len(R), _f32)
    pdf = np.zeros(len(R), _f32)
    draws = np.zeros(len(R), np.int32)
    lib().nmo_sample_volume(dim, C.c_float(lam), _fp(R), seeds.ctypes.data_as(C.POINTER(C.c_uint64)), len(R),
                            _fp(r), _fp(pdf), _ip(draws))
    return r, pdf, draws


def offset_point(dim, p, n):
    p = np.ascontiguousarray(p, _f32).reshape(-1, dim)
    n = np.ascontiguousarray(n, _f32).reshape(-1, dim)
    out = np.zeros_like(p)
    lib().nmo_offset_point(dim, _fp(p), _fp(n), len(p), _fp(out))
    return out
