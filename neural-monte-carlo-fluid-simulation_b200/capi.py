"""ctypes view of libnmcfs.so (include/nmcfs.h).  No compute happens in Python; a missing library or a
missing CUDA device raises -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NMC_LIBNMCFS: another build of the same library (A/B measurements of kernel variants, profiles/tools/ab_walk.py)
LIB_PATH = os.environ.get("NMC_LIBNMCFS") or os.path.join(_HERE, "libnmcfs.so")

MODE_FAST = 0
MODE_DETERMINISTIC = 1

PROBE_DIST_NEUMANN, PROBE_SIGNED_DIST_NEUMANN, PROBE_DIST_DIRICHLET, PROBE_INSIDE_DOMAIN = 0, 1, 2, 3
PROBE_STAR_RADIUS, PROBE_RAY, PROBE_SOURCE, PROBE_GREENS, PROBE_SAMPLE_VOLUME = 4, 5, 6, 7, 8
PROBE_GREENS_FAST, PROBE_SAMPLE_RADIUS_FAST = 9, 10
PROBE_STAR_RADIUS_PACKET, PROBE_RAY_PACKET, PROBE_CLOSEST_PACKET = 11, 12, 13


class SceneOpts(C.Structure):
    _fields_ = [("absorptionCoeff", C.c_float), ("isWatertight", C.c_int), ("isDoubleSided", C.c_int)]


class SolverOpts(C.Structure):
    _fields_ = [("nWalks", C.c_int), ("maxWalkLength", C.c_int),
                ("stepsBeforeApplyingTikhonov", C.c_int), ("stepsBeforeUsingMaximalSpheres", C.c_int),
                ("epsilonShell", C.c_float), ("minStarRadius", C.c_float),
                ("silhouettePrecision", C.c_float), ("russianRouletteThreshold", C.c_float),
                ("useGradientControlVariates", C.c_int), ("useGradientAntitheticVariates", C.c_int),
                ("useCosineSamplingForDerivatives", C.c_int), ("ignoreDirichlet", C.c_int),
                ("ignoreNeumann", C.c_int), ("ignoreSource", C.c_int),
                ("boundaryDistanceMask", C.c_float), ("mode", C.c_int), ("seed", C.c_uint64)]


class BvcOpts(C.Structure):
    """nmc_bvc_opts: the solver options of bvc() beyond SolverOpts (demo.cpp:274-292)."""
    _fields_ = [("boundaryCacheSize", C.c_int), ("domainCacheSize", C.c_int), ("nWalksForCachedSolutionEstimates", C.c_int),
                ("nWalksForCachedGradientEstimates", C.c_int), ("gridRes", C.c_int),
                ("normalOffsetForCachedDirichletSamples", C.c_float), ("radiusClampForKernels", C.c_float),
                ("regularizationForKernels", C.c_float)]


class SolveStats(C.Structure):
    _fields_ = [("walks_started", C.c_uint64), ("walks_completed", C.c_uint64), ("walk_steps", C.c_uint64),
                ("active_points", C.c_uint64), ("kernel_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int), ("warp_trips", C.c_uint64), ("lane_slices", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/nmcfs.h declares (tests check that the library exports them all)
EXPORTS = ["nmc_last_error", "nmc_device_count", "nmc_scene_create", "nmc_scene_destroy", "nmc_scene_set_source",
           "nmc_scene_dim", "nmc_scene_bbox", "nmc_scene_num_nodes", "nmc_scene_nodes", "nmc_wost_solve",
           "nmc_wost_solve_device", "nmc_wost_solve_stats", "nmc_point_seed", "nmc_probe", "nmc_scene_set_source_async",
           "nmc_measure_peaks", "nmc_measure_issue_peak", "nmc_bessel_table", "nmc_estimate_solution", "nmc_bvc_solve", "nmc_bvc_splat"]
SIREN_EXPORTS = ["nmc_siren_last_error", "nmc_siren_forward", "nmc_siren_backward", "nmc_siren_forward_tc", "nmc_siren_weight_grads", "nmc_siren_backward_tc", "nmc_siren_weight_grads_tc", "nmc_siren_backward_fused_tc", "nmc_fit_sample_uniform", "nmc_fit_gather", "nmc_fit_fetch", "nmc_mse_grad_fit", "nmc_adam_update_device", "nmc_adam_update_fetch", "nmc_adam_step", "nmc_adam_step_device", "nmc_mse_grad"]

_lib = None
_fp = C.POINTER(C.c_float)


def lib():
    """Loads libnmcfs.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libnmcfs.so is not built: run `make -C %s` (there is no Python/CPU fallback)" % _HERE)
        L = C.CDLL(LIB_PATH)
        L.nmc_last_error.restype = C.c_char_p
        L.nmc_scene_create.restype = C.c_void_p
        L.nmc_scene_create.argtypes = [C.c_int, _fp, C.c_int, C.POINTER(C.c_int), C.c_int, _fp, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(SceneOpts), C.c_int]
        L.nmc_scene_destroy.argtypes = [C.c_void_p]
        L.nmc_scene_set_source.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.nmc_scene_set_source_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.nmc_measure_peaks.argtypes = [C.c_int, _fp]
        L.nmc_measure_issue_peak.argtypes = [C.c_int, _fp]
        L.nmc_scene_dim.argtypes = [C.c_void_p]
        L.nmc_scene_bbox.argtypes = [C.c_void_p, _fp]
        L.nmc_scene_num_nodes.argtypes = [C.c_void_p]
        L.nmc_scene_nodes.argtypes = [C.c_void_p, _fp]
        L.nmc_wost_solve.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_void_p, C.c_int64, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.POINTER(SolveStats)]
        L.nmc_wost_solve_stats.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_void_p, C.c_int64, C.c_uint64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SolveStats)]
        L.nmc_wost_solve_device.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_void_p, C.c_int64, C.c_uint64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SolveStats)]
        L.nmc_point_seed.restype = C.c_uint64
        L.nmc_point_seed.argtypes = [C.c_uint64, C.c_uint64]
        L.nmc_probe.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.nmc_estimate_solution.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.nmc_bvc_solve.argtypes = [C.c_void_p, C.POINTER(SolverOpts), C.POINTER(BvcOpts), C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nmc_bvc_splat.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                    C.c_float, C.c_void_p]
        _lib = L
    return _lib


def last_error():
    return lib().nmc_last_error().decode()


def check(rc):
    if rc != 0:
        raise RuntimeError("libnmcfs: %s (code %d)" % (last_error(), rc))


def device_count():
    return int(lib().nmc_device_count())


def measure_peaks(device=0):
    """Pipe-throughput peaks of the device in warp-instructions/s: (fp32 FMA, MUFU, fp64 FMA)."""
    out = (C.c_float*3)()
    check(lib().nmc_measure_peaks(int(device), out))
    return float(out[0]), float(out[1]), float(out[2])


def measure_issue_peak(device=0):
    """(dispatch limit in warp-instructions/s = 4 x SMs x SM clock under load, that clock in Hz)."""
    out = (C.c_float*2)()
    check(lib().nmc_measure_issue_peak(int(device), out))
    return float(out[0]), float(out[1])


def bessel_table():
    """(coef[n, 4, 4], t0, per_octave) of the default mode's Bessel lookup table (host-only call)."""
    L = lib()
    L.nmc_bessel_table.argtypes = [_fp, C.c_int, _fp, C.POINTER(C.c_int)]
    t0, po = C.c_float(), C.c_int()
    n = L.nmc_bessel_table(None, 0, C.byref(t0), C.byref(po))
    out = np.zeros((n, 4, 4), np.float32)
    L.nmc_bessel_table(out.ctypes.data_as(_fp), out.size, None, None)
    return out, float(t0.value), int(po.value)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SceneHandle:
    """Owns one nmc_scene (boundary structure + source grid resident on `device`)."""

    def __init__(self, dim, verts, prims, source, absorption=0.0, watertight=False, double_sided=False, device=0):
        L = lib()
        self.dim = int(dim)
        v = _f32(verts).reshape(-1, self.dim)
        p = np.ascontiguousarray(prims, dtype=np.int32).reshape(-1, self.dim)
        src = _f32(source)
        if src.ndim != self.dim:
            raise ValueError("sourceValue must have %d dimensions" % self.dim)
        shp = list(src.shape) + [1] * (3 - self.dim)
        so = SceneOpts(float(absorption), int(bool(watertight)), int(bool(double_sided)))
        self.device = int(device)
        self._h = L.nmc_scene_create(self.dim, v.ctypes.data_as(_fp), len(v), p.ctypes.data_as(C.POINTER(C.c_int)), len(p),
                                     src.ctypes.data_as(_fp), shp[0], shp[1], shp[2], C.byref(so), self.device)
        if not self._h:
            raise RuntimeError("libnmcfs: " + last_error())

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().nmc_scene_destroy(self._h)
            except TypeError:  # interpreter shutdown: module globals are already gone, the driver reclaims the memory
                pass
            self._h = None

    __del__ = close

    def set_source(self, source):
        src = _f32(source)
        shp = list(src.shape) + [1] * (3 - self.dim)
        check(lib().nmc_scene_set_source(self._h, _ptr(src), shp[0], shp[1], shp[2], 0))

    def set_source_device(self, ptr, shape, stream=None):
        """Device-resident grid.  With `stream` (a cudaStream_t value) the copy is enqueued on that stream, i.e. ordered
        after the grid's producer and before a solve_device on the same stream."""
        shp = list(shape) + [1] * (3 - self.dim)
        if stream is None:
            check(lib().nmc_scene_set_source(self._h, C.c_void_p(ptr), shp[0], shp[1], shp[2], 1))
        else:
            check(lib().nmc_scene_set_source_async(self._h, C.c_void_p(ptr), shp[0], shp[1], shp[2], 1, C.c_void_p(stream)))

    def bbox(self):
        out = np.zeros(2 * self.dim, np.float32)
        check(lib().nmc_scene_bbox(self._h, out.ctypes.data_as(_fp)))
        return out[: self.dim].copy(), out[self.dim:].copy()

    def nodes(self):
        n = lib().nmc_scene_num_nodes(self._h)
        out = np.zeros((n, 16), np.float32)
        if n:
            check(lib().nmc_scene_nodes(self._h, out.ctypes.data_as(_fp)))
        return out

    def solve(self, opts, pts, index_offset=0, want_stats12=False):
        """Host buffers in/out (numpy). Returns p[N], grad[N, dim], stats12|None, SolveStats."""
        pts = _f32(pts).reshape(-1, self.dim)
        n = len(pts)
        p = np.zeros(n, np.float32)
        g = np.zeros((n, self.dim), np.float32)
        st12 = np.zeros((n, 12), np.float32) if want_stats12 else None
        st = SolveStats()
        check(lib().nmc_wost_solve_stats(self._h, C.byref(opts), _ptr(pts), n, C.c_uint64(index_offset), _ptr(p), _ptr(g),
                                         _ptr(st12), C.byref(st)))
        return p, g, st12, st

    def solve_ptr(self, opts, pts_ptr, n, p_ptr, g_ptr, index_offset=0, stats=None):
        """Raw HOST pointers (e.g. pinned torch tensors)."""
        check(lib().nmc_wost_solve(self._h, C.byref(opts), C.c_void_p(pts_ptr), n, C.c_uint64(index_offset),
                                   C.c_void_p(p_ptr), C.c_void_p(g_ptr), C.byref(stats) if stats is not None else None))

    def solve_device(self, opts, pts_ptr, n, p_ptr, g_ptr, index_offset=0, stream=0, stats=None):
        """Raw DEVICE pointers on this scene's device, enqueued on `stream` (a cudaStream_t value)."""
        check(lib().nmc_wost_solve_device(self._h, C.byref(opts), C.c_void_p(pts_ptr), n, C.c_uint64(index_offset),
                                          C.c_void_p(p_ptr), C.c_void_p(g_ptr), C.c_void_p(stream),
                                          C.byref(stats) if stats is not None else None))

    def estimate_solution(self, opts, pts, n_walks, normals=None, types=None, aligned=None, index_offset=0):
        """EstimationQuantity::Solution at the given points (deterministic replay); types: 0 in the domain, 2 on the
        reflecting boundary.  Returns (solution[N], stats[N, 4] = variance, averaged walks, mean walk length, first radius)."""
        pts = _f32(pts).reshape(-1, self.dim)
        n = len(pts)
        nr = None if normals is None else _f32(normals).reshape(n, self.dim)
        ty = None if types is None else np.ascontiguousarray(types, dtype=np.int32)
        al = None if aligned is None else np.ascontiguousarray(aligned, dtype=np.int32)
        sol = np.zeros(n, np.float32); st = np.zeros((n, 4), np.float32)
        check(lib().nmc_estimate_solution(self._h, C.byref(opts), _ptr(pts), _ptr(nr), _ptr(ty), _ptr(al), n, int(n_walks),
                                          C.c_uint64(index_offset), _ptr(sol), _ptr(st)))
        return sol, st

    def bvc_solve(self, opts, bvc_opts, cache_cap=1 << 16):
        """Boundary value caching on the device.  Returns (grid[res, res] indexed [i][j] as createEvaluationGrid,
        cache[n, 6] = x, y, nx, ny, estimated solution, pdf, number of domain cache points)."""
        res = int(bvc_opts.gridRes)
        grid = np.zeros((res, res), np.float32)
        cache = np.zeros((cache_cap, 6), np.float32)
        nb, nd = C.c_int(0), C.c_int(0)
        check(lib().nmc_bvc_solve(self._h, C.byref(opts), C.byref(bvc_opts), _ptr(grid), _ptr(cache), cache_cap, C.byref(nb), C.byref(nd)))
        return grid, cache[: min(nb.value, cache_cap)].copy(), nd.value

    def probe(self, kind, n, pts=None, aux0=None, aux1=None, aux2=None, aux3=None, params=None):
        width = {PROBE_RAY: 2 + 2 * self.dim, PROBE_RAY_PACKET: 2 + 2 * self.dim, PROBE_CLOSEST_PACKET: 2, PROBE_GREENS: 10,
                 PROBE_GREENS_FAST: 10, PROBE_SAMPLE_VOLUME: 3, PROBE_SAMPLE_RADIUS_FAST: 2}.get(kind, 1)
        arrs = [None if a is None else _f32(a) for a in (pts, aux0, aux1, aux2, aux3)]
        par = _f32(list(params or []) + [0.0] * (4 - len(params or [])))
        out = np.zeros((n, width), np.float32)
        check(lib().nmc_probe(self._h, kind, n, *[_ptr(a) for a in arrs], _ptr(par), _ptr(out)))
        return out if width > 1 else out[:, 0]
