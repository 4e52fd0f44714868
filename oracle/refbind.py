"""ctypes view of oracle/_ref/libnmc_ref{2d,3d}.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The shared objects are the reference's own solver headers compiled from /root/reference by
oracle/Makefile around oracle/ref_harness.cpp (deterministic seeding rule documented there).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def available(dim, simd=False):
    return os.path.exists(_path(dim, simd))


def _path(dim, simd):
    return os.path.join(_HERE, "_ref", "libnmc_ref%dd%s.so" % (dim, "_simd" if simd else ""))


_libs = {}


def lib(dim, simd=False):
    key = (dim, simd)
    if key not in _libs:
        L = C.CDLL(_path(dim, simd))
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_bbox.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.ref_wost.restype = C.c_int
        L.ref_wost.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_float), C.c_int,
                               C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float),
                               C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_point_seed.restype = C.c_uint64
        L.ref_point_seed.argtypes = [C.c_uint64, C.c_uint64]
        _libs[key] = L
    return _libs[key]


class RefScene:
    """Reference `Scene(config, sourceValue)` (bindings/zombie/demo/scene.h:54-77,
    bindings/zombie3d/demo/scene_3d.h:22-40). `config["boundary"]` is an OBJ path."""

    def __init__(self, dim, config, source, simd=False):
        self.dim = dim
        self.L = lib(dim, simd)
        src = np.ascontiguousarray(source, dtype=_f32)
        assert src.ndim == dim
        shp = list(src.shape) + [1] * (3 - dim)
        self.h = self.L.ref_scene_create(json.dumps(config).encode(), _fp(src), shp[0], shp[1], shp[2])
        self._src = src

    def close(self):
        if self.h:
            self.L.ref_scene_destroy(self.h)
            self.h = None

    def bbox(self):
        out = np.zeros(2 * self.dim, _f32)
        self.L.ref_scene_bbox(self.h, _fp(out))
        return out[: self.dim].copy(), out[self.dim:].copy()

    def wost(self, solver, output, pts, seed=0, index_offset=0, nthreads=1, want_stats=False):
        """Deterministic run of the reference estimator. Returns (p[N], grad[N,dim], stats[N,12]|None)."""
        pts = np.ascontiguousarray(pts, dtype=_f32).reshape(-1, self.dim)
        n = pts.shape[0]
        p = np.zeros(n, _f32)
        g = np.zeros((n, self.dim), _f32)
        st = np.zeros((n, 12), _f32) if want_stats else None
        self.L.ref_wost(self.h, json.dumps(solver).encode(), json.dumps(output).encode(), _fp(pts), n,
                        C.c_uint64(seed), C.c_uint64(index_offset), nthreads, _fp(p), _fp(g),
                        _fp(st) if want_stats else None)
        return p, g, st

    def estimate_solution(self, solver, pts, n_walks, normals=None, types=None, aligned=None, seed=0, index_offset=0, nthreads=1):
        """EstimationQuantity::Solution (walk_on_stars.h:354-461) at the given points; types: 0 InDomain, 2 OnNeumannBoundary.
        Returns (solution[N], stats[N, 4] = variance, number of estimates, mean walk length, first sphere radius)."""
        pts = self._pts(pts)
        n = len(pts)
        nr = np.ascontiguousarray(normals, dtype=_f32).reshape(n, self.dim) if normals is not None else None
        ty = np.ascontiguousarray(types, dtype=np.int32) if types is not None else None
        al = np.ascontiguousarray(aligned, dtype=np.int32) if aligned is not None else None
        sol = np.zeros(n, _f32); st = np.zeros((n, 4), _f32)
        self.L.ref_estimate_solution(C.c_void_p(self.h), json.dumps(solver).encode(), _fp(pts), _fp(nr) if nr is not None else None,
                                     _ip(ty) if ty is not None else None, _ip(al) if al is not None else None, n, int(n_walks),
                                     C.c_uint64(seed), C.c_uint64(index_offset), nthreads, _fp(sol), _fp(st))
        return sol, st

    def bvc(self, solver, output, seed=0, nthreads=1, cache_cap=1 << 16):
        """runBoundaryValueCaching (demo.cpp:265-363) up to the masked evaluation grid (2D only).  Returns
        (grid[gridRes, gridRes] indexed [i][j] as createEvaluationGrid, cache[n, 6] = x, y, nx, ny, solution, pdf, n_domain)."""
        res = int(output["gridRes"])
        grid = np.zeros((res, res), _f32)
        cache = np.zeros((cache_cap, 6), _f32)
        nb = C.c_int(0); nd = C.c_int(0)
        self.L.ref_bvc(C.c_void_p(self.h), json.dumps(solver).encode(), json.dumps(output).encode(), C.c_uint64(seed), nthreads,
                       _fp(grid), _fp(cache), cache_cap, C.byref(nb), C.byref(nd))
        return grid, cache[: min(nb.value, cache_cap)].copy(), nd.value

    # ---- probes -----------------------------------------------------------------------------
    def _pts(self, pts):
        return np.ascontiguousarray(pts, dtype=_f32).reshape(-1, self.dim)

    def dist_neumann(self, pts, signed=False):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.ref_dist_neumann(C.c_void_p(self.h), _fp(pts), len(pts), int(signed), _fp(out))
        return out

    def dist_dirichlet(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.ref_dist_dirichlet(C.c_void_p(self.h), _fp(pts), len(pts), _fp(out))
        return out

    def inside_domain(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), np.int32)
        self.L.ref_inside_domain(C.c_void_p(self.h), _fp(pts), len(pts), _ip(out))
        return out

    def outside_bbox(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), np.int32)
        self.L.ref_outside_bbox(C.c_void_p(self.h), _fp(pts), len(pts), _ip(out))
        return out

    def star_radius(self, pts, min_r, max_r, prec=1e-3, flip=False):
        pts = self._pts(pts)
        mr = np.ascontiguousarray(np.broadcast_to(np.asarray(max_r, _f32), (len(pts),)))
        out = np.zeros(len(pts), _f32)
        self.L.ref_star_radius(C.c_void_p(self.h), _fp(pts), len(pts), C.c_float(min_r), _fp(mr),
                               C.c_float(prec), int(flip), _fp(out))
        return out

    def intersect_neumann(self, org, nrm, dirs, tmax, onb):
        org, nrm, dirs = self._pts(org), self._pts(nrm), self._pts(dirs)
        n = len(org)
        tm = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, _f32), (n,)))
        ob = np.ascontiguousarray(np.broadcast_to(np.asarray(onb, np.int32), (n,)))
        out = np.zeros((n, 2 + 2 * self.dim), _f32)
        self.L.ref_intersect_neumann(C.c_void_p(self.h), _fp(org), _fp(nrm), _fp(dirs), _fp(tm), _ip(ob), n, _fp(out))
        return out

    def blocked(self, xi, xj, ni, nj, offi, offj):
        xi, xj, ni, nj = self._pts(xi), self._pts(xj), self._pts(ni), self._pts(nj)
        n = len(xi)
        oi = np.ascontiguousarray(np.broadcast_to(np.asarray(offi, np.int32), (n,)))
        oj = np.ascontiguousarray(np.broadcast_to(np.asarray(offj, np.int32), (n,)))
        out = np.zeros(n, np.int32)
        self.L.ref_blocked(C.c_void_p(self.h), _fp(xi), _fp(xj), _fp(ni), _fp(nj), _ip(oi), _ip(oj), n, _ip(out))
        return out

    def sample_neumann(self, pts, radius, rnd):
        pts, rnd = self._pts(pts), self._pts(rnd)
        n = len(pts)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, _f32), (n,)))
        out = np.zeros((n, 2 + 2 * self.dim), _f32)
        self.L.ref_sample_neumann(C.c_void_p(self.h), _fp(pts), _fp(rad), _fp(rnd), n, _fp(out))
        return out

    def source(self, pts):
        pts = self._pts(pts)
        out = np.zeros(len(pts), _f32)
        self.L.ref_source(C.c_void_p(self.h), _fp(pts), len(pts), _fp(out))
        return out


# ---- scene-free probes -----------------------------------------------------------------------
def pcg32_uint(dim, initstate, initseq, n):
    out = np.zeros(n, np.uint32)
    lib(dim).ref_pcg32_uint(C.c_uint64(initstate), C.c_uint64(initseq), n, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def pcg32_float(dim, initstate, initseq, n):
    out = np.zeros(n, _f32)
    lib(dim).ref_pcg32_float(C.c_uint64(initstate), C.c_uint64(initseq), n, _fp(out))
    return out


def pcg32_bounded(dim, initstate, initseq, bounds):
    b = np.ascontiguousarray(bounds, np.uint32)
    out = np.zeros(len(b), np.uint32)
    lib(dim).ref_pcg32_bounded(C.c_uint64(initstate), C.c_uint64(initseq), b.ctypes.data_as(C.POINTER(C.c_uint32)),
                               len(b), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def point_seed(dim, seed, index):
    return int(lib(dim).ref_point_seed(C.c_uint64(seed), C.c_uint64(index)))


def stratified(dim, initstate, n_samples):
    out = np.zeros((dim - 1) * n_samples, _f32)
    st = np.zeros(2, np.uint64)
    lib(dim).ref_stratified(C.c_uint64(initstate), n_samples, _fp(out), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out, st


def sphere_dir(dim, u):
    u = np.ascontiguousarray(u, _f32).reshape(-1, dim - 1)
    out = np.zeros((len(u), dim), _f32)
    lib(dim).ref_sphere_dir(_fp(u), len(u), _fp(out))
    return out


def bessel(dim, kind, x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    lib(dim).ref_bessel(kind, x.ctypes.data_as(C.POINTER(C.c_double)), x.size, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def greens_ball(dim, lam, R, r):
    R = np.ascontiguousarray(R, _f32)
    r = np.ascontiguousarray(r, _f32)
    out = np.zeros((len(R), 10), _f32)
    lib(dim).ref_greens_ball(C.c_float(lam), _fp(R), _fp(r), len(R), _fp(out))
    return out


def sample_volume(dim, lam, R, seeds):
    R = np.ascontiguousarray(R, _f32)
    seeds = np.ascontiguousarray(seeds, np.uint64)
    r = np.zeros(len(R), _f32)
    pdf = np.zeros(len(R), _f32)
    draws = np.zeros(len(R), np.int32)
    lib(dim).ref_sample_volume(C.c_float(lam), _fp(R), seeds.ctypes.data_as(C.POINTER(C.c_uint64)), len(R),
                               _fp(r), _fp(pdf), _ip(draws))
    return r, pdf, draws


def offset_point(dim, p, n):
    p = np.ascontiguousarray(p, _f32).reshape(-1, dim)
    n = np.ascontiguousarray(n, _f32).reshape(-1, dim)
    out = np.zeros_like(p)
    lib(dim).ref_offset_point(_fp(p), _fp(n), len(p), _fp(out))
    return out
