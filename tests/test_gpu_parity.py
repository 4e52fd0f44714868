"""Parity tests proper: the CUDA path, called through the C ABI (libnmcfs.so), against the oracle and the
golden vectors generated from the reference.  Run with `-m gpu` on the B200 box."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
V = os.path.join(util.GOLDEN, "vectors")


@pytest.fixture(scope="module")
def pkg():
    p = util.package()
    assert p.capi.device_count() > 0, "no CUDA device visible"
    return p


def _scene(pkg, cfg, src=None):
    dim = cfg["dim"]
    return pkg.Scene(cfg["scene"], util.source_grid(dim) if src is None else src, device=0)


def _bits_equal(a, b):
    return np.ascontiguousarray(a, np.float32).view(np.uint32) == np.ascontiguousarray(b, np.float32).view(np.uint32)


@pytest.mark.parametrize("case", list(util.CASES))
def test_device_queries_against_golden(pkg, case):
    """Flattened-BVH queries on the device vs the reference's FCPW results: distances, signed distances,
    inside test, source lookup, star radius and closest-hit rays are IEEE float work compiled with
    -fmad=false, so they are expected to be bit-exact."""
    c = pkg.capi
    k = np.load(os.path.join(V, case + ".npz"))
    cfg = util.load_case(case)
    dim = cfg["dim"]
    sc = _scene(pkg, cfg)
    h = sc.handle
    lo, hi = h.bbox()
    assert _bits_equal(lo, k["bbox_lo"]).all() and _bits_equal(hi, k["bbox_hi"]).all()
    q = k["q"]; n = len(q)
    assert _bits_equal(h.probe(c.PROBE_DIST_NEUMANN, n, q), k["dist"]).all()
    assert _bits_equal(h.probe(c.PROBE_SIGNED_DIST_NEUMANN, n, q), k["sdist"]).all()
    assert _bits_equal(h.probe(c.PROBE_DIST_DIRICHLET, n, q), k["ddist"]).all()
    assert np.array_equal(h.probe(c.PROBE_INSIDE_DOMAIN, n, q) > 0, k["inside"] > 0)
    assert _bits_equal(h.probe(c.PROBE_SOURCE, n, q), k["source"]).all()
    # the cone test uses acos/asin/atan2 (double on the device, glibc float on the host): a culling decision
    # can differ only in borderline cases, so allow a tiny mismatch rate there
    for flip, key in ((0.0, "star0"), (1.0, "star1")):
        s = h.probe(c.PROBE_STAR_RADIUS, n, q, aux0=k["ddist"], params=[1e-3, 1e-3, flip])
        assert _bits_equal(s, k[key]).mean() >= 0.999
    ray = h.probe(c.PROBE_RAY, n, q, aux0=np.zeros_like(q), aux1=k["dirs"], aux2=k["tmax"], aux3=np.zeros(n, np.float32))
    assert np.array_equal(ray[:, 0], k["ray"][:, 0])
    hit = ray[:, 0] > 0
    assert _bits_equal(ray[hit], k["ray"][hit]).all()
    m = len(k["onb_p"])
    if m:
        ray = h.probe(c.PROBE_RAY, m, k["onb_p"], aux0=k["onb_n"], aux1=k["onb_d"], aux2=k["onb_t"], aux3=np.ones(m, np.float32))
        assert np.array_equal(ray[:, 0], k["onb_ray"][:, 0])
        hit = ray[:, 0] > 0
        assert _bits_equal(ray[hit], k["onb_ray"][hit]).all()


def test_device_special_functions_against_golden(pkg):
    """Bessel-based ball Green's functions and the rejection sampler of the deterministic mode.
    Double-precision polynomials + exp/log/sqrt on the device vs glibc on the host: equal after
    narrowing to float except for rare last-bit differences."""
    c = pkg.capi
    k = np.load(os.path.join(V, "special.npz"))
    R, r = k["R"], k["r"]
    seeds = np.ascontiguousarray(k["seeds"]).view(np.uint32).astype(np.uint32).view(np.float32)
    for case, dim in (("karman", 2), ("smoke3d", 3)):
        h = _scene(pkg, util.load_case(case)).handle
        for lam in (350.0, 0.0):
            g = h.probe(c.PROBE_GREENS, len(R), None, aux0=R, aux1=r, params=[lam])
            ref = k["greens_%d_%g" % (dim, lam)]
            fin = np.isfinite(ref) & np.isfinite(g)
            assert (np.isfinite(ref) == np.isfinite(g)).mean() > 0.999
            rel = np.abs(g[fin] - ref[fin])/np.maximum(np.abs(ref[fin]), 1e-30)
            assert (rel < 2e-6).mean() > 0.999, rel.max()
            sv = h.probe(c.PROBE_SAMPLE_VOLUME, len(R), None, aux0=R, aux1=seeds, params=[lam])
            same_draws = sv[:, 2] == k["sv_draws_%d_%g" % (dim, lam)]
            assert same_draws.mean() > 0.995  # an accept/reject decision may flip on a 1-ulp difference
            ref_r = k["sv_r_%d_%g" % (dim, lam)][same_draws]
            if dim == 3 and lam == 0.0:  # closed form with cbrt and cos (distributions.h:483-496): 1-ulp differences allowed
                assert np.allclose(sv[same_draws, 0], ref_r, rtol=1e-6, atol=0)
            else:  # r = nextFloat()*R: exact
                assert _bits_equal(sv[same_draws, 0], ref_r).mean() > 0.999


@pytest.mark.parametrize("case", list(util.CASES))
def test_deterministic_mode_against_reference_vectors(pkg, case, oracle_lib):
    """north_star criterion 1: deterministic mode, per-point estimates within 1e-5 relative of the
    reference on the same seeds and walk counts (golden vectors from oracle/_ref; the oracle is run live
    as a second witness)."""
    k = np.load(os.path.join(V, case + ".npz"))
    cfg = util.load_case(case)
    dim = cfg["dim"]
    sc = _scene(pkg, cfg)
    p, g, st12, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], k["pts"], mode=pkg.capi.MODE_DETERMINISTIC,
                                           seed=int(k["seed"]), want_stats=True)
    assert st.kernel_launches >= 1
    if case == "taylorgreen_shipped":
        assert not p.any() and not g.any() and st.walks_started == 0
        return
    assert st.walks_started == 500*int((k["stats"][:, 11] > 0).sum())
    # identical walk histories <=> identical number of averaged walks per point
    assert (st12[:, 9] == k["stats"][:, 9]).mean() >= 0.99
    okp = util.close_mask(p, k["p"]); okg = util.close_mask(g, k["g"])
    assert okp.mean() >= 0.99, "p: %.4f within tolerance" % okp.mean()
    assert okg.mean() >= 0.99, "grad: %.4f within tolerance" % okg.mean()
    # outliers (a flipped accept/reject decision in one of 500 walks) stay within 3 standard errors
    se = np.sqrt(np.maximum(k["stats"][:, 1], 0)/np.maximum(k["stats"][:, 9], 1))
    assert (np.abs(p - k["p"]) <= 3*se + 1e-12)[~okp].all()
    osc = oracle_lib.OracleScene(dim, cfg["scene"], util.source_grid(dim))
    op, og, _ = osc.wost(cfg["solver"], cfg["output"], k["pts"], seed=int(k["seed"]), nthreads=4)
    assert util.close_mask(p, op).mean() >= 0.99 and util.close_mask(g, og).mean() >= 0.99


@pytest.mark.parametrize("case", ["taylorgreen_active", "karman", "smoke3d", "karman3d", "channel_circle", "box_sphere"])
def test_fast_mode_matches_reference_statistically(pkg, case):
    """north_star criterion 2: default mode vs the reference's per-point means within 3 standard errors,
    variance ratios near 1 (the reference's means/variances come from the golden vectors' SampleStatistics)."""
    k = np.load(os.path.join(V, case + ".npz"))
    cfg = util.load_case(case)
    dim = cfg["dim"]
    sc = _scene(pkg, cfg)
    p, g, s, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], k["pts"], mode=pkg.capi.MODE_FAST, seed=12345, want_stats=True)
    ref = k["stats"]
    act = ref[:, 11] > 0
    assert ((s[:, 11] > 0) <= act).all() and st.walks_started > 0  # masked boundary points are not walked in fast mode
    both = act & (s[:, 11] > 0)
    nf, nr = np.maximum(s[both, 9], 1), np.maximum(ref[both, 9], 1)
    # completion rates agree (escaped / over-long walks are discarded on both sides)
    assert abs(nf.mean() - nr.mean()) < 0.03*500
    z = (s[both, 0] - ref[both, 0])/np.sqrt(s[both, 1]/nf + ref[both, 1]/nr + 1e-30)
    assert (np.abs(z) < 3).mean() >= 0.98, "solution z-scores: %.3f within 3 sigma" % (np.abs(z) < 3).mean()
    assert abs(z.mean()) < 0.35, "systematic bias in p: mean z = %.3f" % z.mean()
    for d in range(dim):
        zg = (s[both, 2 + d] - ref[both, 2 + d])/np.sqrt(s[both, 5 + d]/nf + ref[both, 5 + d]/nr + 1e-30)
        assert (np.abs(zg) < 3).mean() >= 0.98, "gradient z-scores dim %d: %.3f" % (d, (np.abs(zg) < 3).mean())
        assert abs(zg.mean()) < 0.35
    vr = np.median(s[both, 1]/np.maximum(ref[both, 1], 1e-30))
    assert 0.7 < vr < 1.4, "median solution variance ratio %.3f" % vr
    vg = np.median(s[both, 5]/np.maximum(ref[both, 5], 1e-30))
    assert 0.6 < vg < 1.6, "median gradient variance ratio %.3f" % vg


@pytest.mark.parametrize("variant", list(util.OPTION_VARIANTS))
@pytest.mark.parametrize("case", list(util.OPTION_CASES))
def test_option_variants_in_both_modes(pkg, case, variant):
    """SURVEY 8(f) rank 2, the part the bindings can express: Tikhonov switch-over after k harmonic steps, double-sided
    boundaries, control / antithetic variates off, Russian roulette off, maximal spheres, ignoreSource.
    Deterministic mode against the reference's vectors (same bar as the shipped configs), default mode statistically."""
    k = np.load(os.path.join(V, "options.npz"))
    key = case + "/" + variant
    cfg = util.load_variant(case, variant)
    dim = cfg["dim"]
    sc = _scene(pkg, cfg)
    pts, ref = k[key + "/pts"], k[key + "/stats"]
    nw = cfg["solver"]["nWalks"]
    p, g, st12, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_DETERMINISTIC, seed=9, want_stats=True)
    assert st.walks_started == (nw if variant == "no_anti" else 2*(nw//2))*int((ref[:, 11] > 0).sum())
    assert (st12[:, 9] == ref[:, 9]).mean() >= 0.98, key
    assert util.close_mask(p, k[key + "/p"]).mean() >= 0.98 and util.close_mask(g, k[key + "/g"]).mean() >= 0.98, key
    pf, gf, s, stf = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=4242, want_stats=True)
    assert np.isfinite(pf).all() and np.isfinite(gf).all()
    both = (ref[:, 11] > 0) & (s[:, 11] > 0)
    if variant == "no_rr":       # every walk runs out of steps and is discarded on both sides
        assert not ref[:, 9].any() and not s[:, 9].any() and not pf.any()
        return
    if variant == "ignore_source":
        assert not pf.any() and not gf.any() and not k[key + "/p"].any()
        return
    nf, nr = np.maximum(s[both, 9], 1), np.maximum(ref[both, 9], 1)
    assert abs(nf.mean() - nr.mean()) < 0.04*nw, key
    assert abs(s[both, 10].mean() - ref[both, 10].mean()) < 0.1*max(ref[both, 10].mean(), 0.2), key   # mean walk length
    z = (s[both, 0] - ref[both, 0])/np.sqrt(s[both, 1]/nf + ref[both, 1]/nr + 1e-30)
    assert (np.abs(z) < 3.2).mean() >= 0.95 and abs(z.mean()) < 0.5, (key, z.mean())
    for d in range(dim):
        zg = (s[both, 2 + d] - ref[both, 2 + d])/np.sqrt(s[both, 5 + d]/nf + ref[both, 5 + d]/nr + 1e-30)
        assert (np.abs(zg) < 3.2).mean() >= 0.95 and abs(zg.mean()) < 0.5, (key, d, zg.mean())
    vr = np.median(s[both, 1]/np.maximum(ref[both, 1], 1e-30))
    assert 0.6 < vr < 1.6, (key, vr)


@pytest.mark.parametrize("mode", ["fast", "det"])
def test_size_independent_properties_at_scale(pkg, mode):
    """Properties checked at a size the CPU oracle would need minutes for: linearity in the source (a power
    of two scales every estimate exactly), zero source -> zero, and invariance to how points are sharded
    (RNG keyed by the global point index)."""
    m = pkg.capi.MODE_FAST if mode == "fast" else pkg.capi.MODE_DETERMINISTIC
    cfg = util.load_case("karman")
    n = 65536 if mode == "fast" else 8192
    src = util.source_grid(2)
    sc = _scene(pkg, cfg, src)
    lo, hi = sc.bbox()
    pts = util.random_points(lo, hi, n, seed=2)
    p1, g1, _, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=m, seed=7)
    assert np.isfinite(p1).all() and np.isfinite(g1).all()
    assert st.walks_started == 500*st.active_points
    sc.handle.set_source(4.0*src)
    p4, g4, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=m, seed=7)
    if mode == "det":
        assert np.array_equal(p4, 4.0*p1) and np.array_equal(g4, 4.0*g1)
    else:  # shared-memory float atomics feed the control variate: summation order may differ between runs
        assert util.close_mask(p4, 4.0*p1, rtol=1e-4, atol_scale=1e-5).mean() > 0.999
        assert util.close_mask(g4, 4.0*g1, rtol=1e-3, atol_scale=1e-4).mean() > 0.99
    sc.handle.set_source(np.zeros_like(src))
    p0, g0, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[:4096], mode=m, seed=7)
    assert not p0.any() and not g0.any()
    sc.handle.set_source(src)
    cut = n//3
    pa, ga, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[:cut], mode=m, seed=7, index_offset=0)
    pb, gb, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[cut:], mode=m, seed=7, index_offset=cut)
    ps, gs = np.concatenate([pa, pb]), np.concatenate([ga, gb])
    if mode == "det":
        assert np.array_equal(ps, p1) and np.array_equal(gs, g1)
    else:
        assert util.close_mask(ps, p1, rtol=1e-4, atol_scale=1e-5).mean() > 0.999


@pytest.mark.parametrize("n_circle", [24, 128, 400])
def test_mesh_size_regimes_of_the_default_mode(pkg, oracle_lib, tmp_path, n_circle):
    """The default mode picks its data path by mesh size: flat tables in shared memory (<= 128 primitives), beyond that
    warp-packet traversals of the tree through L1/L2 (csrc/nmc_packet.cuh).
    Each regime is checked against the oracle: bit-exact star radii and statistically equal estimates."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mss", os.path.join(util.GOLDEN, "make_synthetic_scenes.py"))
    mss = importlib.util.module_from_spec(spec); spec.loader.exec_module(mss)
    v, e = mss.channel_circle(n_circle=n_circle, wall_split=2 if n_circle > 24 else 1)
    obj = str(tmp_path/"mesh.obj")
    mss.write_obj(obj, "mesh size regime test", v, e, "l")
    cfg = util.load_case("karman")
    cfg["scene"]["boundary"] = obj
    src = util.source_grid(2)
    sc = pkg.Scene(cfg["scene"], src, device=0)
    osc = oracle_lib.OracleScene(2, cfg["scene"], src)
    lo, hi = sc.bbox()
    q = util.random_points(lo, hi, 3000, seed=9)
    dd = osc.dist_dirichlet(q)
    star = sc.handle.probe(pkg.capi.PROBE_STAR_RADIUS, len(q), q, aux0=dd, params=[1e-3, 1e-3, 0.0])
    assert _bits_equal(star, osc.star_radius(q, 1e-3, dd, 1e-3, False)).mean() >= 0.999
    pts = util.random_points(lo, hi, 96, seed=4)
    _, _, ref = osc.wost(cfg["solver"], cfg["output"], pts, seed=5, nthreads=4, want_stats=True)
    p, g, s, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=77, want_stats=True)
    both = (ref[:, 11] > 0) & (s[:, 11] > 0)
    assert both.sum() > 48
    nf, nr = np.maximum(s[both, 9], 1), np.maximum(ref[both, 9], 1)
    assert abs(nf.mean() - nr.mean()) < 0.03*500
    z = (s[both, 0] - ref[both, 0])/np.sqrt(s[both, 1]/nf + ref[both, 1]/nr + 1e-30)
    assert (np.abs(z) < 3).mean() >= 0.97 and abs(z.mean()) < 0.4
    for d in range(2):
        zg = (s[both, 2 + d] - ref[both, 2 + d])/np.sqrt(s[both, 5 + d]/nf + ref[both, 5 + d]/nr + 1e-30)
        assert (np.abs(zg) < 3).mean() >= 0.97 and abs(zg.mean()) < 0.4


@pytest.mark.parametrize("case", ["channel_circle", "box_sphere", "karman", "karman3d"])
def test_warp_packet_tree_queries_on_the_device(pkg, oracle_lib, case):
    """The default mode's big-mesh path (csrc/nmc_packet.cuh: one tree traversal per warp, per-lane radii) through the packet
    probes: every lane must get what the deterministic kernel's private traversal gets, on coherent packets (32 queries in one
    small ball, the walk kernel's situation) and incoherent ones (32 queries anywhere), with idle lanes at the end.  The fast
    file is compiled with -use_fast_math (approximate divisions), hence 1e-5 relative instead of bits."""
    cfg = util.load_case(case)
    dim = cfg["dim"]
    sc = _scene(pkg, cfg)
    lo, hi = sc.bbox()
    ext = float((hi - lo).max())
    rng = np.random.default_rng(31)
    far = util.random_points(lo, hi, 4096, seed=5)
    centres = util.random_points(lo, hi, 128, seed=6)
    near = (np.repeat(centres, 32, 0) + (rng.random((4096, dim), dtype=np.float32) - 0.5)*0.06*ext).astype(np.float32)
    # packets ON the boundary (where a walk continues after a reflection) with every other lane moved off it: on the fine
    # circle a lane that worked on a leaf its own cone test had dropped would accept vertices inside the precision band
    v, pr = oracle_lib.load_obj(cfg["scene"]["boundary"], dim, False)
    first = rng.integers(0, max(1, len(pr) - 32), 128)
    idx = np.minimum((first[:, None] + np.arange(32)[None, :]).reshape(-1), len(pr) - 1)
    mixed = v[pr[idx]].mean(1).astype(np.float32)[:, :dim]
    mixed[1::2] += ((rng.random((len(mixed)//2, dim), dtype=np.float32) - 0.5)*0.03*ext).astype(np.float32)
    q = np.ascontiguousarray(np.concatenate([far, mixed, near])[:-13])
    n = len(q)
    max_r = (rng.random(n, dtype=np.float32)*ext).astype(np.float32); max_r[::7] = np.float32(3.0e38); max_r[4096:8192] = np.float32(3.0e38)
    for flip in (0.0, 1.0):
        want = sc.handle.probe(pkg.capi.PROBE_STAR_RADIUS, n, q, aux0=max_r, params=[1e-3, 1e-3, flip])
        got = sc.handle.probe(pkg.capi.PROBE_STAR_RADIUS_PACKET, n, q, aux0=max_r, params=[1e-3, 1e-3, flip])
        rel = np.abs(got - want)/np.maximum(np.abs(want), 1e-6)
        assert (rel < 1e-5).mean() >= 0.9995, (case, flip, (rel < 1e-5).mean(), rel.max())
    u = rng.random((n, 2), dtype=np.float32)
    if dim == 2:
        a = 2*np.pi*u[:, 0]; d = np.stack([np.cos(a), np.sin(a)], 1).astype(np.float32)
    else:
        z = 1 - 2*u[:, 0]; r = np.sqrt(np.maximum(0, 1 - z*z)); a = 2*np.pi*u[:, 1]
        d = np.stack([r*np.cos(a), r*np.sin(a), z], 1).astype(np.float32)
    tmax = (rng.random(n, dtype=np.float32)*ext).astype(np.float32)
    nrm = np.zeros_like(q); onb = np.zeros(n, np.float32)
    want = sc.handle.probe(pkg.capi.PROBE_RAY, n, q, aux0=nrm, aux1=d, aux2=tmax, aux3=onb)
    got = sc.handle.probe(pkg.capi.PROBE_RAY_PACKET, n, q, aux0=nrm, aux1=d, aux2=tmax, aux3=onb)
    off = np.ones(n, bool); off[4096:8192] = False   # a ray that starts ON a primitive hits it at t = 0 +- rounding: not a traversal property
    assert want[off, 0].sum() > 100
    assert (got[off, 0] == want[off, 0]).mean() >= 0.9995, case
    both = (got[:, 0] > 0) & (want[:, 0] > 0) & off
    assert (np.abs(got[both, 1] - want[both, 1]) <= 2e-5*np.abs(want[both, 1]) + 1e-6).all(), case
    assert (np.abs(got[both, 2:2 + dim] - want[both, 2:2 + dim]) <= 1e-4*ext).all(), case
    assert (np.abs(got[both, 2 + dim:] - want[both, 2 + dim:]).max(1) < 1e-5).mean() >= 0.999, case   # equidistant hits may pick the neighbour
    cl = sc.handle.probe(pkg.capi.PROBE_CLOSEST_PACKET, n, q)
    dn = sc.handle.probe(pkg.capi.PROBE_DIST_NEUMANN, n, q)
    sd = sc.handle.probe(pkg.capi.PROBE_SIGNED_DIST_NEUMANN, n, q)
    assert (np.abs(cl[:, 0] - dn) <= 1e-5*np.abs(dn) + 1e-7).mean() >= 0.9999, case
    assert (np.sign(cl[off, 1]) == np.sign(sd[off])).mean() >= 0.999, case      # ON a primitive the side is decided by rounding


def test_edge_cases(pkg):
    cfg = util.load_case("karman")
    sc = _scene(pkg, cfg)
    lo, hi = sc.bbox()
    for m in (pkg.capi.MODE_FAST, pkg.capi.MODE_DETERMINISTIC):
        # empty input
        p, g, _, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], np.zeros((0, 2), np.float32), mode=m, seed=1)
        assert p.shape == (0,) and g.shape == (0, 2)
        # one point, odd and tiny walk counts (nWalks/2 pairs, at least one)
        mid = ((lo + hi)/2).reshape(1, 2).astype(np.float32) + np.float32(0.3)*(hi - lo)*np.array([[0.5, 0.2]], np.float32)
        for nw in (1, 3, 501):
            p, g, _, st = pkg.zombie.wost_array(sc, dict(cfg["solver"], nWalks=nw), cfg["output"], mid, mode=m, seed=1)
            assert st.walks_started == 2*max(1, nw//2) and np.isfinite(p).all()
        # points inside the obstacle / outside the channel are classified outside (watertight): gradient masked
        far = np.array([[lo[0] - 1.0, lo[1] - 1.0], [hi[0] + 5.0, hi[1] + 2.0]], np.float32)
        p, g, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], far, mode=m, seed=1)
        assert np.isfinite(p).all() and not g.any()
    # the public list-based API returns nested lists like the pybind module
    pts, sol, grad = pkg.wost(sc, cfg["solver"], cfg["output"], [[0.1, 0.2], [0.3, 0.1]])
    assert isinstance(sol, list) and isinstance(grad[0], list) and len(grad[0]) == 2 and pts[1][0] == pytest.approx(0.3)
    with pytest.raises(RuntimeError, match="unknown mode"):
        pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], mid, mode=7)


def test_full_size_3d_properties_one_million_points(pkg):
    """BASELINE.json configs[4] size (3D, ~1e6 query points/step, 5e8 walks): finite results, exact walk
    accounting, linearity in the source and shard invariance checked on slices."""
    cfg = util.load_case("smoke3d")
    src = util.source_grid(3)
    sc = _scene(pkg, cfg, src)
    lo, hi = sc.bbox()
    n = 1000000
    pts = util.random_points(lo, hi, n, seed=8)
    p, g, _, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=pkg.capi.MODE_FAST, seed=21)
    assert np.isfinite(p).all() and np.isfinite(g).all()
    assert st.walks_started == 500*st.active_points and st.active_points > 0.99*n
    assert st.walks_completed >= 0.999*st.walks_started  # closed cube: (almost) no walk escapes (reference: 100 % on 1024 points)
    sc.handle.set_source(2.0*src)
    p2, g2, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[:50000], mode=pkg.capi.MODE_FAST, seed=21)
    assert util.close_mask(p2, 2.0*p[:50000], rtol=1e-4, atol_scale=1e-5).mean() > 0.999
    sc.handle.set_source(src)
    pb, gb, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[900000:], mode=pkg.capi.MODE_FAST, seed=21, index_offset=900000)
    assert util.close_mask(pb, p[900000:], rtol=1e-4, atol_scale=1e-5).mean() > 0.999
    # a smooth source gives a smooth field: the Monte Carlo estimate must correlate with itself across seeds
    p3, _, _, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts[:50000], mode=pkg.capi.MODE_FAST, seed=22)
    assert np.corrcoef(p[:50000], p3)[0, 1] > 0.9


@pytest.mark.parametrize("dim", [2, 3])
def test_compiled_drop_in_module_on_the_gpu(dim):
    """`import zombie_bindings` exactly as src/2d / src/3d do (module directory on sys.path): Scene(config, nested
    lists), wost(...) -> (points, p, grad p) as nested lists, identical to the C ABI results; additive wost_array,
    set_mode, set_seed, last_stats."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "bindings_check.py"), str(dim), "gpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BINDINGS_OK gpu dim=%d" % dim in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_random_meshes_against_the_oracle(pkg, oracle_lib, tmp_path):
    """Beyond the fixtures: seeded random polygons / perturbed icospheres (closed and open, both orientations,
    double-sided; the oracle is pinned bit-for-bit to the reference on the same meshes by the CPU suite).
    Device queries bit-exact, deterministic estimator at the 1e-5 bar, default mode statistically."""
    c = pkg.capi
    for name, dim, cfg in util.random_meshes(tmp_path):
        src = util.source_grid(dim)
        sc = pkg.Scene(cfg["scene"], src, device=0)
        osc = oracle_lib.OracleScene(dim, cfg["scene"], src)
        h = sc.handle
        lo, hi = osc.bbox()
        q = util.random_points(lo, hi, 2000, seed=5, margin=0.1); n = len(q)
        for kind, sg in ((c.PROBE_DIST_NEUMANN, False), (c.PROBE_SIGNED_DIST_NEUMANN, True)):
            got, want = h.probe(kind, n, q), osc.dist_neumann(q, sg)
            eq = _bits_equal(got, want)
            rel = np.abs(got - want)/np.maximum(np.abs(want), 1e-30)
            print("%s %s distance: %d of %d differ in bits, max relative difference %.2e" % (name, "signed" if sg else "unsigned", int((~eq).sum()), n, rel.max()))
            assert eq.mean() >= 0.995 and rel.max() <= 5e-7, (name, sg, int((~eq).sum()), rel.max())
        assert np.array_equal(h.probe(c.PROBE_INSIDE_DOMAIN, n, q) > 0, osc.inside_domain(q) > 0), name
        dd = osc.dist_dirichlet(q)
        for flip in (0.0, 1.0):
            s = h.probe(c.PROBE_STAR_RADIUS, n, q, aux0=dd, params=[1e-3, 1e-3, flip])
            assert _bits_equal(s, osc.star_radius(q, 1e-3, dd, 1e-3, bool(flip))).mean() >= 0.999, (name, flip)
        pts = util.random_points(lo, hi, 96, seed=11)
        op, og, ost = osc.wost(cfg["solver"], cfg["output"], pts, seed=3, nthreads=8, want_stats=True)
        p, g, st12, st = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=c.MODE_DETERMINISTIC, seed=3, want_stats=True)
        assert (st12[:, 9] == ost[:, 9]).mean() >= 0.96, name
        okp, okg = util.close_mask(p, op), util.close_mask(g, og)
        assert okp.mean() >= 0.96 and okg.mean() >= 0.96, (name, okp.mean(), okg.mean())   # 96 points: up to 3 flipped decisions
        se = np.sqrt(np.maximum(ost[:, 1], 0)/np.maximum(ost[:, 9], 1))
        assert (np.abs(p - op) <= 3*se + 1e-12)[~okp].all(), name                            # and those stay within 3 standard errors
        pf, gf, s, _ = pkg.zombie.wost_array(sc, cfg["solver"], cfg["output"], pts, mode=c.MODE_FAST, seed=99, want_stats=True)
        both = (ost[:, 11] > 0) & (s[:, 11] > 0)
        if both.sum() < 16:
            continue
        nf, nr = np.maximum(s[both, 9], 1), np.maximum(ost[both, 9], 1)
        assert abs(nf.mean() - nr.mean()) < 0.05*500, (name, nf.mean(), nr.mean())
        z = (s[both, 0] - ost[both, 0])/np.sqrt(s[both, 1]/nf + ost[both, 1]/nr + 1e-30)
        assert (np.abs(z) < 3.5).mean() >= 0.94 and abs(z.mean()) < 0.6, (name, (np.abs(z) < 3.5).mean(), z.mean())
