"""Golden vectors of the solution-only estimator and of boundary value caching, generated from the reference's own
headers (oracle/_ref, built by oracle/Makefile).  Run in the build container (needs /root/reference):
    python tests/golden/make_bvc_vectors.py      -> tests/golden/vectors/bvc.npz
Contents per case: sample points in the domain and ON the boundary (with normals), the reference's
EstimationQuantity::Solution estimates (walk_on_stars.h:354-461) for a fixed seed, and one boundary-value-caching
evaluation grid (demo.cpp:265-363) with its cache points."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import util  # noqa: E402
from oracle import refbind  # noqa: E402

OUT = os.path.join(HERE, "vectors", "bvc.npz")
BVC_SOLVER = {"boundaryCacheSize": 2048, "domainCacheSize": 2048, "nWalksForCachedSolutionEstimates": 64}
BVC_GRID = 48


def boundary_points(pkg, cfg, n, seed):
    """n points on the scene's segments (uniform in the segment index and along it) with the reference's sample normal
    (s_y, -s_x)/|s| (sampleLineSegmentUniformly, sampling.h:213-224)."""
    v, p = pkg.zombie.load_obj(cfg["scene"]["boundary"], 2, bool(cfg["scene"].get("flipOrientation", False)))
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(p), n)
    u = rng.random(n, dtype=np.float32)
    a, b = v[p[idx, 0]], v[p[idx, 1]]
    s = b - a
    pts = (a + u[:, None]*s).astype(np.float32)
    nr = np.stack([s[:, 1], -s[:, 0]], axis=1)
    nr = (nr/np.linalg.norm(nr, axis=1, keepdims=True)).astype(np.float32)
    return pts, nr


def cases():
    yield "karman", util.load_case("karman")
    yield "taylorgreen_active", util.load_case("taylorgreen_active")
    ds = util.load_case("karman"); ds["scene"]["isDoubleSided"] = True
    yield "karman_doublesided", ds


def main():
    pkg = util.package()
    d = {}
    for name, cfg in cases():
        src = util.source_grid(2)
        sc = refbind.RefScene(2, cfg["scene"], src)
        lo, hi = sc.bbox()
        dom = util.random_points(lo, hi, 128, seed=31)
        bpts, bnr = boundary_points(pkg, cfg, 128, seed=32)
        pts = np.concatenate([dom, bpts]); nr = np.concatenate([np.zeros_like(dom), bnr])
        ty = np.concatenate([np.zeros(len(dom), np.int32), np.full(len(bpts), 2, np.int32)])
        al = np.zeros(len(pts), np.int32)
        if cfg["scene"].get("isDoubleSided"):
            al[len(dom)::2] = 1  # every other boundary point estimates the normal-aligned side
        sol, st = sc.estimate_solution(cfg["solver"], pts, 48, normals=nr, types=ty, aligned=al, seed=17, nthreads=8)
        k = name + "/"
        d[k + "pts"], d[k + "normals"], d[k + "types"], d[k + "aligned"], d[k + "solution"], d[k + "stats"] = pts, nr, ty, al, sol, st
        solver = dict(cfg["solver"], **BVC_SOLVER)
        out = dict(cfg["output"], gridRes=BVC_GRID)
        grids = []
        for rep in range(4):  # four independent runs: their spread is the yardstick of the statistical comparison
            grid, cache, nd = sc.bvc(solver, out, seed=100 + rep, nthreads=8)
            grids.append(grid)
        d[k + "bvc_grids"], d[k + "bvc_cache"], d[k + "bvc_n_domain"] = np.stack(grids), cache, np.int32(nd)
        sc.close()
        print(name, "solution mean|u| %.3e, averaged walks %.1f of 48; bvc: %d boundary + %d domain cache points, grid mean|u| %.3e, run-to-run rms %.3e" % (
            np.abs(sol).mean(), st[:, 1].mean(), len(cache), nd, np.abs(grids[0]).mean(), np.std(np.stack(grids), axis=0).mean()))
    np.savez_compressed(OUT, **d)


if __name__ == "__main__":
    main()
