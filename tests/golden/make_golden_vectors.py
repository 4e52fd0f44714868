"""Generates tests/golden/vectors/*.npz from the REFERENCE ITSELF (oracle/_ref, i.e. the reference's
own headers compiled by oracle/Makefile around oracle/ref_harness.cpp).  Run in the build container:
    make -C oracle ref && python tests/golden/make_golden_vectors.py
The vectors pin (i) the pcg32 stream, (ii) Bessel / ball Green's function tables, (iii) geometric
queries on the four example meshes, (iv) per-point (p, grad p) of the deterministic estimator.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402
from oracle import refbind  # noqa: E402

OUT = os.path.join(HERE, "vectors")


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2024)
    # (i) RNG + samplers
    kat = {}
    for dim in (2, 3):
        kat["uint_%d" % dim] = refbind.pcg32_uint(dim, 42, 1, 64)
        kat["float_%d" % dim] = refbind.pcg32_float(dim, 0x9E3779B97F4A7C15, 1, 64)
        b = np.arange(1, 129, dtype=np.uint32)
        kat["bounds"] = b
        kat["bounded_%d" % dim] = refbind.pcg32_bounded(dim, 7, 1, b)
        s, st = refbind.stratified(dim, 99, 100)
        kat["strat_%d" % dim] = s
        kat["strat_state_%d" % dim] = st
        u = rng.random((256, dim - 1), dtype=np.float32)
        kat["sphere_u_%d" % dim] = u
        kat["sphere_%d" % dim] = refbind.sphere_dir(dim, u)
        kat["point_seed"] = np.array([refbind.point_seed(dim, 5, i) for i in range(16)], np.uint64)
    np.savez_compressed(os.path.join(OUT, "rng.npz"), **kat)
    # (ii) special functions
    x = np.concatenate([np.geomspace(1e-4, 200, 400), rng.random(200)*20])
    sf = {"x": x}
    for k in range(5):
        sf["bessel_%d" % k] = refbind.bessel(2, k, x)
    R = np.concatenate([rng.random(400)*1.2 + 1e-3, rng.random(50)*9]).astype(np.float32)
    r = (rng.random(len(R))*R).astype(np.float32)
    seeds = rng.integers(0, 2**63, len(R), dtype=np.uint64)
    sf["R"], sf["r"], sf["seeds"] = R, r, seeds
    for dim in (2, 3):
        for lam in (350.0, 0.0):
            sf["greens_%d_%g" % (dim, lam)] = refbind.greens_ball(dim, lam, R, r)
            a, b, c = refbind.sample_volume(dim, lam, R, seeds)
            sf["sv_r_%d_%g" % (dim, lam)], sf["sv_pdf_%d_%g" % (dim, lam)], sf["sv_draws_%d_%g" % (dim, lam)] = a, b, c
    np.savez_compressed(os.path.join(OUT, "special.npz"), **sf)
    # (iii) + (iv) per scene
    for case in util.CASES:
        cfg = util.load_case(case)
        dim = cfg["dim"]
        src = util.source_grid(dim)
        sc = refbind.RefScene(dim, cfg["scene"], src)
        lo, hi = sc.bbox()
        d = {"bbox_lo": lo, "bbox_hi": hi}
        q = util.random_points(lo, hi, 2000, seed=5, margin=0.1)
        d["q"] = q
        d["dist"] = sc.dist_neumann(q)
        d["sdist"] = sc.dist_neumann(q, True)
        d["ddist"] = sc.dist_dirichlet(q)
        d["inside"] = sc.inside_domain(q)
        d["source"] = sc.source(q)
        d["star0"] = sc.star_radius(q, 1e-3, d["ddist"], 1e-3, False)
        d["star1"] = sc.star_radius(q, 1e-3, d["ddist"], 1e-3, True)
        dirs = refbind.sphere_dir(dim, rng.random((len(q), dim - 1), dtype=np.float32))
        tmax = (rng.random(len(q), dtype=np.float32)*(hi - lo).max()).astype(np.float32)
        d["dirs"], d["tmax"] = dirs, tmax
        ray = sc.intersect_neumann(q, np.zeros_like(q), dirs, tmax, 0)
        d["ray"] = ray
        hit = ray[:, 0] > 0
        hp, hn = ray[hit][:, 2:2 + dim].copy(), ray[hit][:, 2 + dim:].copy()
        d2 = dirs[: len(hp)].copy()
        d2[(hn*d2).sum(1) > 0] *= -1
        d["onb_p"], d["onb_n"], d["onb_d"], d["onb_t"] = hp, hn, d2, tmax[: len(hp)]
        d["onb_ray"] = sc.intersect_neumann(hp, hn, d2, tmax[: len(hp)], 1)
        d["onb_star"] = sc.star_radius(hp, 1e-3, sc.dist_dirichlet(hp), 1e-3, False)
        pts = util.random_points(lo, hi, 192, seed=11)
        p, g, st = sc.wost(cfg["solver"], cfg["output"], pts, seed=3, index_offset=0, nthreads=8, want_stats=True)
        d["pts"], d["seed"], d["p"], d["g"], d["stats"] = pts, np.uint64(3), p, g, st
        np.savez_compressed(os.path.join(OUT, case + ".npz"), **d)
        sc.close()
        print(case, "ok: mean|p| %.3e  completed/started %.3f" % (np.abs(p).mean(), st[:, 9].mean()/500))
    # (v) option variants (deterministic estimator only)
    d = {}
    for case in util.OPTION_CASES:
        for name in util.OPTION_VARIANTS:
            cfg = util.load_variant(case, name)
            dim = cfg["dim"]
            sc = refbind.RefScene(dim, cfg["scene"], util.source_grid(dim))
            lo, hi = sc.bbox()
            pts = util.random_points(lo, hi, 64, seed=21)
            p, g, st = sc.wost(cfg["solver"], cfg["output"], pts, seed=9, index_offset=0, nthreads=8, want_stats=True)
            k = case + "/" + name
            d[k + "/pts"], d[k + "/p"], d[k + "/g"], d[k + "/stats"] = pts, p, g, st
            sc.close()
            print("variant", k, "completed/started %.3f  mean walk length %.2f" % (st[:, 9].mean()/cfg["solver"]["nWalks"], st[:, 10].mean()))
    np.savez_compressed(os.path.join(OUT, "options.npz"), **d)


if __name__ == "__main__":
    main()
