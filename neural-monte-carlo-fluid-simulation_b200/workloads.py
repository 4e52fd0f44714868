"""Synthetic workloads of the shape each example configuration produces: scene fixture,
source grid with the resolution the time-stepper writes, seeded query points.  Shared by bench.py and the
bench-size parity tests so that what is benchmarked is what is checked against the reference.

The scene fixtures (OBJ + wost.json) live under tests/golden/scenes; they are inputs, not code.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")

# case -> (fixture name, scene overrides).  taylorgreen_active: isWatertight false, the solver-active variant of
# the Taylor-Green scene (as shipped every point is classified outside and wost() returns zeros)
CASES = {
    "taylorgreen_active": ("taylorgreen", {"isWatertight": False}),
    "taylorgreen_shipped": ("taylorgreen", {}),
    "karman": ("karman", {}),
    "smoke3d": ("smoke3d", {}),
    "karman3d": ("karman3d", {}),
    "channel_circle": ("channel_circle", {}),
    "box_sphere": ("box_sphere", {}),
}

# divergence-grid shapes of the time-stepper (sample_uniform_2D(1000, with_boundary=True): 1002 samples along the
# longer axis, int(1000*ratio) + 2 along the other; 3D: vis_resolution 80 + 2 per axis)
GRID_SHAPES = {"taylorgreen_active": (1002, 1002), "taylorgreen_shipped": (1002, 1002), "karman": (401, 1002),
               "channel_circle": (401, 1002), "smoke3d": (82, 82, 82), "karman3d": (82, 82, 82), "box_sphere": (82, 82, 82)}


def load_case(case):
    name, over = CASES[case]
    cfg = json.load(open(os.path.join(SCENES, name + ".json")))
    cfg["scene"]["boundary"] = os.path.join(SCENES, name + ".obj")
    cfg["scene"].update(over)
    return cfg


def source_grid(case):
    """Smooth synthetic divergence field on the grid the time-stepper would write for this scene."""
    shp = GRID_SHAPES[case]
    if len(shp) == 2:
        h, w = shp
        y, x = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
        return (np.sin(6.1*x)*np.sin(4.3*y + 0.3)).astype(np.float32)
    g = [np.linspace(0, 1, s) for s in shp]
    x, y, z = np.meshgrid(*g, indexing="ij")
    return (np.sin(6.1*x)*np.sin(4.3*y + 0.3)*np.cos(3*z)).astype(np.float32)


def random_points(lo, hi, n, seed=0):
    """Uniform points in the scene's bounding box (sample_random_2D, utils/model_utils.py:22-31)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    return (rng.random((n, len(lo)), dtype=np.float32)*(hi - lo) + lo).astype(np.float32)


def walks_per_point(solver):
    """One walk = one sampler.seed() + walk() call (walk_on_stars.h:579-581): 2 * max(1, nWalks / 2) per active point
    with antithetic pairs, nWalks without."""
    nw = int(solver.get("nWalks", 128))
    return nw if solver.get("disableGradientAntitheticVariates", False) else 2*max(1, nw//2)


def scene_size_from_obj(path, dim=2):
    """cfg.scene_size the way the reference's drivers derive it (src/2d/main.py:36-45, src/3d: the same with z): the
    bounding box of the OBJ's vertices, (x0, x1, y0, y1[, z0, z1]).  The training / pressure samples are drawn in THIS box:
    the Taylor-Green square spans [0.000447, 6.279553], not [0, 2 pi], and a sample outside the boundary mesh starts
    walks outside the domain."""
    v = []
    for line in open(path):
        t = line.split()
        if t and t[0] == "v":
            v.append([float(c) for c in t[1:1 + dim]])
    v = np.array(v)
    lo, hi = v.min(0), v.max(0)
    out = []
    for k in range(dim):
        out += [float(lo[k]), float(hi[k])]
    return tuple(out)


def taylor_green_initial(scene_size):
    """taylorgreen_velocity (src/2d/sources.py:19-31): the samples are rescaled from the scene box to (0, 2 pi)."""
    import math

    import torch

    def fn(x):
        a = (x[:, 0] - scene_size[0])/(scene_size[1] - scene_size[0])*(2*math.pi)
        b = (x[:, 1] - scene_size[2])/(scene_size[3] - scene_size[2])*(2*math.pi)
        return torch.stack([torch.sin(a)*torch.cos(b), -torch.cos(a)*torch.sin(b)], dim=-1)
    return fn


def karman_obstacle(mask=1e-3):
    """Centre and radius of the cylinder of the karman fixture the way src/2d/main.py:36-57,90-102 derives them:
    the vertices strictly inside the bounding box, mean centre, mean radius + output.boundaryDistanceMask.
    Returns (centre, radius, scene_size)."""
    v = []
    for line in open(os.path.join(SCENES, "karman.obj")):
        t = line.split()
        if t and t[0] == "v":
            v.append([float(t[1]), float(t[2])])
    v = np.array(v)
    lo, hi = v.min(0), v.max(0)
    inner = v[(v[:, 0] > lo[0]) & (v[:, 0] < hi[0]) & (v[:, 1] > lo[1]) & (v[:, 1] < hi[1])]
    inner = np.unique(inner, axis=0)
    c = inner.mean(0)
    return (float(c[0]), float(c[1])), float(np.linalg.norm(inner - c, axis=1).mean() + mask), (float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1]))
