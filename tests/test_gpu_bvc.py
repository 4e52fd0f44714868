"""Boundary value caching (bindings/zombie/demo/demo.cpp:265-363) and the solution-only
estimator it runs at its cache points (row f2; walk_on_stars.h:354-461) against the reference: golden vectors generated
from the reference's own headers (tests/golden/make_bvc_vectors.py), and the live reference build where it is present."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
VEC = os.path.join(util.GOLDEN, "vectors", "bvc.npz")
CASES = ["karman", "taylorgreen_active", "karman_doublesided"]


def _cfg(name):
    if name == "karman_doublesided":
        cfg = util.load_case("karman"); cfg["scene"]["isDoubleSided"] = True
        return cfg
    return util.load_case(name)


@pytest.fixture(scope="module")
def pkg():
    p = util.package()
    assert p.capi.device_count() > 0
    return p


@pytest.fixture(scope="module")
def vec():
    return np.load(VEC)


@pytest.mark.parametrize("name", CASES)
def test_solution_estimator_replays_the_reference(pkg, vec, name):
    """EstimationQuantity::Solution at 128 points in the domain and 128 points ON the reflecting boundary (walks start on
    the boundary with its normal; every other one on the back side in the double-sided scene): the deterministic replay
    reproduces the reference's estimates to the 1e-5 bar, its averaged-walk counts and its first sphere radii."""
    cfg = _cfg(name)
    sc = pkg.Scene(cfg["scene"], util.source_grid(2), device=0)
    k = name + "/"
    sol, st = pkg.zombie.estimate_solution(sc, cfg["solver"], cfg["output"], vec[k + "pts"], 48, normals=vec[k + "normals"],
                                           types=vec[k + "types"], aligned=vec[k + "aligned"], seed=17)
    ref, rst = vec[k + "solution"], vec[k + "stats"]
    assert np.array_equal(st[:, 3], rst[:, 3])                                  # first sphere radius: bit-exact geometry
    assert (st[:, 1] == rst[:, 1]).mean() >= 0.99                               # the same walks were averaged in
    ok = util.close_mask(sol, ref)
    assert ok.mean() >= 0.99, (name, ok.mean())
    bad = ~ok                                                                   # a flipped accept/reject decision moves a point by O(1/48)
    se = np.sqrt(rst[:, 0]/np.maximum(rst[:, 1], 1)) + 1e-12
    assert (np.abs(sol[bad] - ref[bad]) <= 3*se[bad]).all()
    on_b = vec[k + "types"] == 2
    assert ok[on_b].mean() >= 0.99
    if name != "taylorgreen_active":  # (its boundary is the bounding box: walks started on it leave the box and are discarded, in the reference too)
        assert np.abs(ref[on_b]).max() > 0                                      # the boundary starts carry signal


def test_solution_estimator_rejects_dirichlet_samples(pkg):
    cfg = _cfg("karman")
    sc = pkg.Scene(cfg["scene"], util.source_grid(2), device=0)
    with pytest.raises(RuntimeError, match="sample type"):
        pkg.zombie.estimate_solution(sc, cfg["solver"], cfg["output"], np.zeros((2, 2), np.float32), 4, types=[0, 1])


def _live(name, cfg):
    from oracle import refbind
    if not refbind.available(2):
        pytest.skip("oracle/_ref is not built")
    return refbind.RefScene(2, cfg["scene"], util.source_grid(2))


@pytest.mark.parametrize("name", ["karman", "karman_doublesided"])
def test_solution_estimator_against_live_reference_at_more_points(pkg, name):
    cfg = _cfg(name)
    ref = _live(name, cfg)
    sc = pkg.Scene(cfg["scene"], util.source_grid(2), device=0)
    lo, hi = sc.bbox()
    pts = util.random_points(lo, hi, 2048, seed=77)
    rsol, rst = ref.estimate_solution(cfg["solver"], pts, 32, seed=5, nthreads=8)
    sol, st = pkg.zombie.estimate_solution(sc, cfg["solver"], cfg["output"], pts, 32, seed=5)
    ref.close()
    assert np.array_equal(st[:, 3], rst[:, 3])
    assert util.close_mask(sol, rsol).mean() >= 0.99


@pytest.mark.parametrize("absorption", [0.0, 4.0])
def test_splat_kernel_matches_double_precision_formulas(pkg, absorption):
    """Splatter::splat (splatter.h:203-290) with the free-space Green's functions of distributions.h:85-219:
    mean over boundary samples of (G dudn - P u)/pdf plus mean over source samples of G f/pdf, in numpy double."""
    from scipy.special import k0, k1
    rng = np.random.default_rng(3)
    ne, nb, nd = 500, 300, 200
    x = rng.random((ne, 2))*2 - 1
    cache = np.zeros((nb + nd, 8), np.float32)
    cache[:, :2] = rng.random((nb + nd, 2))*2 - 1
    ang = rng.random(nb)*2*np.pi
    cache[:nb, 2], cache[:nb, 3] = np.cos(ang), np.sin(ang)
    cache[:, 4] = rng.standard_normal(nb + nd)
    cache[:nb, 5] = rng.standard_normal(nb)*0.3
    cache[:nb, 6], cache[nb:, 6] = 0.125, 0.25
    cache[:nb:3, 7] = 2.0                                                      # every third boundary sample is normal-aligned
    cache[nb:, 7] = 1.0
    dd = np.full(ne, 1.0, np.float32); dd[::50] = 0.0                          # some evaluation points below the cut-off
    out = np.full(ne, -7.0, np.float32)
    import ctypes as C
    L = pkg.capi.lib()
    xe = np.ascontiguousarray(x, np.float32)
    pkg.capi.check(L.nmc_bvc_splat(2, C.c_float(absorption), xe.ctypes.data_as(C.c_void_p), dd.ctypes.data_as(C.c_void_p), ne,
                                   cache.ctypes.data_as(C.c_void_p), nb + nd, C.c_float(1e-3), C.c_float(0.0), C.c_float(0.5),
                                   out.ctypes.data_as(C.c_void_p)))
    c = cache.astype(np.float64); xe = xe.astype(np.float64)
    d = xe[:, None, :] - c[None, :, :2]
    r = np.maximum(np.linalg.norm(d, axis=2), 1e-3)
    sign = np.where(c[:, 7] == 2.0, -1.0, 1.0)
    ndot = (d*(c[None, :, 2:4]*sign[None, :, None])).sum(2)
    if absorption > 0:
        mu = np.sqrt(absorption)
        G = k0(mu*r)/(2*np.pi); P = ndot*mu*k1(mu*r)/(2*np.pi*r)
    else:
        G = -np.log(r)/(2*np.pi); P = ndot/(2*np.pi*r*r)
    est_b = (G*c[None, :, 5] - P*c[None, :, 4])/c[None, :, 6]
    est_s = G*c[None, :, 4]/c[None, :, 6]
    grp0, grp1, grp2 = c[:, 7] == 0.0, c[:, 7] == 2.0, c[:, 7] == 1.0
    ref = est_b[:, grp0].mean(1) + est_b[:, grp1].mean(1) + est_s[:, grp2].mean(1)
    live = dd >= 0.5
    assert np.all(out[~live] == -7.0)                                          # below the cut-off: untouched (splatter.h:60)
    assert np.abs(out[live] - ref[live]).max() <= 2e-5*np.abs(ref[live]).max()


@pytest.mark.parametrize("name", CASES)
def test_bvc_grid_matches_the_reference_statistically(pkg, vec, name):
    """The whole pipeline (cache points, boundary-start estimates, splat, mask) against four independent runs of the
    reference (runBoundaryValueCaching): same masked cells, the same number of cache points, and a field that differs
    from the reference's mean by no more than the reference's own runs differ from one another."""
    cfg = _cfg(name)
    k = name + "/"
    ref = vec[k + "bvc_grids"]
    solver = dict(cfg["solver"], boundaryCacheSize=2048, domainCacheSize=2048, nWalksForCachedSolutionEstimates=64)
    out = dict(cfg["output"], gridRes=ref.shape[1])
    sc = pkg.Scene(cfg["scene"], util.source_grid(2), device=0)
    runs = []
    for rep in range(4):
        grid, cache, nd = pkg.zombie.bvc_grid(sc, solver, out, seed=200 + rep, want_cache=True)
        runs.append(grid)
    ours = np.stack(runs)
    assert abs(len(cache) - len(vec[k + "bvc_cache"])) <= 2 and abs(nd - int(vec[k + "bvc_n_domain"])) <= 0.02*max(nd, 1) + 2
    assert np.array_equal(ours[0] == 0.0, ref[0] == 0.0)                       # masking: bit-exact geometry
    live = ref[0] != 0.0
    rm, om = ref.mean(0)[live], ours.mean(0)[live]
    spread = ref.std(0, ddof=1)[live]                                          # per-cell run-to-run deviation of the reference
    scale = np.abs(rm).mean()
    # the two means of four runs differ by ~ sigma * sqrt(2 / 4) per cell; sigma from the pooled within-group variance of
    # the eight runs (6 degrees of freedom, so z follows Student's t_6: P(|t_6| > 4) = 0.7 %)
    pooled = np.sqrt(0.5*(ref.var(0, ddof=1)[live] + ours.var(0, ddof=1)[live]))
    z = (om - rm)/(pooled*np.sqrt(0.5) + 1e-3*scale)
    assert (np.abs(z) < 4).mean() >= 0.98, (name, (np.abs(z) < 4).mean())
    assert abs(z.mean()) < 0.25, z.mean()
    assert abs((om - rm).mean()) <= 0.25*spread.mean() + 1e-3*scale
    noise = 0.25*(pooled**2).mean()                                            # variance of a mean of four runs
    signal = max(rm.var() - noise, 0.0)
    assert np.corrcoef(om, rm)[0, 1] >= signal/(signal + noise) - 0.05          # as correlated as two noisy copies of one field can be
    ratio = ours.std(0, ddof=1)[live].mean()/max(spread.mean(), 1e-30)        # the same estimator: the same noise level
    assert 0.6 < ratio < 1.6, ratio


def test_bvc_writes_the_reference_files(pkg, tmp_path):
    """bvc() returns None and writes solutionFile plus <stem>_color<ext> (demo/grid.h:9-33); PFM rows are flipped on write
    (image.h:173-198), image row <-> y index, column <-> x index (grid.h:388-411)."""
    cfg = _cfg("karman")
    sc = pkg.Scene(cfg["scene"], util.source_grid(2), device=0)
    solver = dict(cfg["solver"], boundaryCacheSize=256, domainCacheSize=256, nWalksForCachedSolutionEstimates=8)
    path = str(tmp_path/"sub"/"solution.pfm")
    out = dict(cfg["output"], gridRes=24, solutionFile=path, colormap="turbo", colormapMinVal=-1e-3, colormapMaxVal=1e-3)
    pkg.zombie.set_defaults(seed=9)
    try:
        assert pkg.zombie.bvc(sc, solver, out) is None
        grid = pkg.zombie.bvc_grid(sc, solver, out)
    finally:
        pkg.zombie._DEFAULTS["seed"] = None
    with open(path, "rb") as f:
        assert f.readline() == b"PF\n" and f.readline() == b"24 24\n" and f.readline() == b"-1\n"
        img = np.frombuffer(f.read(), "<f4").reshape(24, 24, 3)
    assert np.array_equal(img[::-1, :, 0], grid.T) and np.array_equal(img[..., 0], img[..., 2])
    col = np.fromfile(str(tmp_path/"sub"/"solution_color.pfm"), "<f4", offset=len(b"PF\n24 24\n-1\n")).reshape(24, 24, 3)
    assert col.min() >= 0.0 and col.max() <= 1.0 and np.ptp(col) > 0.1
