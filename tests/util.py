"""Shared helpers for the tests: fixtures, synthetic inputs, loaders for the oracle and the product."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(GOLDEN, "scenes")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# (fixture, scene overrides): the four distinct example geometries; Taylor-Green both as shipped
# (isWatertight:true -> every point is classified outside -> zeros) and with the solver active.
CASES = {
    "taylorgreen_active": ("taylorgreen", {"isWatertight": False}),
    "taylorgreen_shipped": ("taylorgreen", {}),
    "karman": ("karman", {}),
    "smoke3d": ("smoke3d", {}),
    "karman3d": ("karman3d", {}),
    # SURVEY 8(d) M-BVH: > 128 primitives, so the default mode walks the tree (tests/golden/make_synthetic_scenes.py)
    "channel_circle": ("channel_circle", {}),
    "box_sphere": ("box_sphere", {}),
}


def load_case(case):
    name, over = CASES[case]
    cfg = json.load(open(os.path.join(SCENES, name + ".json")))
    cfg["scene"]["boundary"] = os.path.join(SCENES, name + ".obj")
    cfg["scene"].update(over)
    return cfg


def source_grid(dim, scale=1.0):
    """Smooth synthetic source on a small grid (non-square on purpose: row/column mix-ups show up)."""
    if dim == 2:
        h, w = 101, 103
        y, x = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
        return (scale*np.sin(6.1*x)*np.sin(4.3*y + 0.3)).astype(np.float32)
    n0, n1, n2 = 22, 23, 24
    x, y, z = np.meshgrid(np.linspace(0, 1, n0), np.linspace(0, 1, n1), np.linspace(0, 1, n2), indexing="ij")
    return (scale*np.sin(6.1*x)*np.sin(4.3*y + 0.3)*np.cos(3*z)).astype(np.float32)


def random_points(lo, hi, n, seed=0, margin=0.0):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    ext = hi - lo
    return (rng.random((n, len(lo)), dtype=np.float32)*ext*(1 + 2*margin) + lo - margin*ext).astype(np.float32)


def package():
    return importlib.import_module("neural-monte-carlo-fluid-simulation_b200")


def close_mask(a, b, rtol=1e-5, atol_scale=1e-7):
    """|a-b| <= rtol*max(|a|,|b|) + atol, atol = atol_scale * mean|channel|."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    atol = atol_scale*max(np.abs(b).mean(), 1e-30)
    return np.abs(a - b) <= rtol*np.maximum(np.abs(a), np.abs(b)) + atol


def karman_obstacle(mask=1e-3):
    return package().workloads.karman_obstacle(mask)


# Solver / scene options beyond the shipped configs that the bindings can express:
# name -> (scene overrides, solver overrides).  Golden vectors: tests/golden/vectors/options.npz.
OPTION_VARIANTS = {
    "tikhonov2": ({}, {"setpsBeforeApplyingTikhonov": 2}),                       # two harmonic steps, then screened (walk_on_stars.h:319-321)
    "harmonic_then_yukawa": ({}, {"setpsBeforeApplyingTikhonov": 1, "maxWalkLength": 64}),
    "doublesided": ({"isDoubleSided": True}, {}),                                # normal flipping on thin boundaries (:154-160)
    "no_cv": ({}, {"disableGradientControlVariates": True}),
    "no_anti": ({}, {"disableGradientAntitheticVariates": True, "nWalks": 120}),
    "no_rr": ({}, {"russianRouletteThreshold": 0.0, "maxWalkLength": 256}),      # nothing stops a walk: all exceed the length and are discarded
    "maxsph": ({}, {"setpsBeforeUsingMaximalSpheres": 1}),
    "ignore_source": ({}, {"ignoreSource": True}),
    "cosine": ({}, {"useCosineSamplingForDirectionalDerivatives": True}),    # first boundary sample from a cosine lobe around +/- e_x (:550-554)
}
OPTION_CASES = ("karman", "karman3d")


def load_variant(case, variant):
    cfg = load_case(case)
    so, sv = OPTION_VARIANTS[variant]
    cfg["scene"].update(so); cfg["solver"].update(sv)
    return cfg


def random_meshes(tmpdir):
    """Seeded random boundaries beyond the fixtures: star-shaped polygons (closed / open chains, both orientations,
    double-sided) and perturbed icospheres (both orientations, with holes).  Yields (name, dim, cfg)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mss", os.path.join(GOLDEN, "make_synthetic_scenes.py"))
    mss = importlib.util.module_from_spec(spec); spec.loader.exec_module(mss)
    rng = np.random.default_rng(7)
    out = []
    for k in range(4):
        n = int(rng.integers(5, 60))
        th = np.sort(rng.random(n))*2*np.pi
        rad = 0.5 + 0.25*rng.random(n)
        v = np.stack([rad*np.cos(th), rad*np.sin(th)], 1)
        e = np.stack([np.arange(n), (np.arange(n) + 1) % n], 1)
        if k % 2:
            e = e[:, ::-1]
        if k >= 2:
            e = e[:-2]
        out.append(("poly%d" % k, 2, v, e, {"isWatertight": k < 2, "isDoubleSided": k == 3}))
    for k in range(3):
        sv, sf = mss.icosphere(1 + (k % 2))
        sv = sv*(0.6 + 0.15*rng.random((len(sv), 1)))
        if k == 1:
            sf = sf[:, ::-1]
        if k == 2:
            sf = sf[:-7]
        out.append(("ico%d" % k, 3, sv, sf, {"isWatertight": k < 2, "isDoubleSided": k == 2}))
    for name, dim, v, e, over in out:
        obj = os.path.join(str(tmpdir), name + ".obj")
        mss.write_obj(obj, name, v, e, "l" if dim == 2 else "f")
        cfg = load_case("karman" if dim == 2 else "smoke3d")
        cfg["scene"]["boundary"] = obj
        cfg["scene"].update(over)
        yield name, dim, cfg
