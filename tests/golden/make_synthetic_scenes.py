"""Writes the two synthetic boundary meshes of SURVEY.md section 8(d) row M-BVH into tests/golden/scenes/:
meshes large enough (> 128 primitives) that the default mode walks the BVH / SNCH instead of scanning flat
tables.  No reference config has such a mesh; the reference's solver (oracle/_ref) handles them like any other.

  channel_circle.obj  2D: the karman channel (walls of karman.obj, each segment split in 4) with the cylinder
                      replaced by a 1024-gon of the same centre, radius and orientation           (1184 segments)
  box_sphere.obj      3D: the smoke3d cube plus an icosphere obstacle (3 subdivisions, 1280 triangles) oriented
                      so that the fluid is outside the sphere                                      (1292 triangles)
Deterministic (no RNG); coordinates are printed with 9 significant digits so every float parses back exactly.
    python tests/golden/make_synthetic_scenes.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = os.path.join(HERE, "scenes")


def read_obj(path, dim):
    v, e = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(c) for c in t[1:1 + dim]])
        elif t[0] in ("l", "f"):
            e.append([int(c.split("/")[0]) - 1 for c in t[1:1 + dim]])
    return np.array(v, np.float64), np.array(e, np.int64)


def write_obj(path, header, verts, prims, tag):
    with open(path, "w") as f:
        f.write("# %s\n" % header)
        for p in verts.astype(np.float32):
            f.write("v " + " ".join("%.9g" % c for c in p) + (" 0" if len(p) == 2 else "") + "\n")
        for e in prims:
            f.write(tag + " " + " ".join(str(int(i) + 1) for i in e) + "\n")


def channel_circle(n_circle=1024, wall_split=4):
    v, e = read_obj(os.path.join(SCENES, "karman.obj"), 2)
    # the cylinder is the only closed loop: the component containing vertex 0
    nxt = {int(a): int(b) for a, b in e}
    loop, i = [0], nxt[0]
    while i != 0:
        loop.append(i); i = nxt[i]
    c = v[loop].mean(0)
    r = np.linalg.norm(v[loop] - c, axis=1).mean()
    a = v[loop]
    area = 0.5*np.sum(a[:, 0]*np.roll(a[:, 1], -1) - np.roll(a[:, 0], -1)*a[:, 1])
    th = np.arange(n_circle)*(2*np.pi/n_circle)*(1.0 if area > 0 else -1.0)
    verts = [c + r*np.stack([np.cos(th), np.sin(th)], 1)]
    prims = [np.stack([np.arange(n_circle), (np.arange(n_circle) + 1) % n_circle], 1)]
    base = n_circle
    inloop = set(loop)
    for a_, b_ in e:
        if int(a_) in inloop:
            continue
        t = np.linspace(0, 1, wall_split + 1)[:, None]
        seg = v[a_]*(1 - t) + v[b_]*t          # unshared end points on purpose: open chains, like the walls' ends
        verts.append(seg)
        prims.append(np.stack([base + np.arange(wall_split), base + np.arange(wall_split) + 1], 1))
        base += wall_split + 1
    return np.concatenate(verts), np.concatenate(prims)


def icosphere(level):
    t = (1 + 5**0.5)/2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64)/np.linalg.norm(p) for p in v]
    for _ in range(level):
        mid, nf = {}, []

        def m(a, b):
            k = (min(a, b), max(a, b))
            if k not in mid:
                p = v[a] + v[b]
                v.append(p/np.linalg.norm(p)); mid[k] = len(v) - 1
            return mid[k]
        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), np.array(f, np.int64)


def box_sphere(level=3, centre=(0.2, -0.1, 0.1), radius=0.4):
    v, f = read_obj(os.path.join(SCENES, "smoke3d.obj"), 3)
    vol = np.sum(np.einsum("ij,ij->i", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])))/6   # > 0: normals point out of the cube
    sv, sf = icosphere(level)                                                            # outward-facing
    if vol > 0:
        sf = sf[:, ::-1]                                                                   # fluid outside the sphere
    sv = sv*radius + np.array(centre)
    return np.concatenate([v, sv]), np.concatenate([f, sf + len(v)])


def main():
    v2, e2 = channel_circle()
    write_obj(os.path.join(SCENES, "channel_circle.obj"), "synthetic M-BVH mesh, made by make_synthetic_scenes.py", v2, e2, "l")
    v3, f3 = box_sphere()
    write_obj(os.path.join(SCENES, "box_sphere.obj"), "synthetic M-BVH mesh, made by make_synthetic_scenes.py", v3, f3, "f")
    for src, dst, obj in (("karman", "channel_circle", "channel_circle.obj"), ("smoke3d", "box_sphere", "box_sphere.obj")):
        cfg = json.load(open(os.path.join(SCENES, src + ".json")))
        cfg["scene"]["boundary"] = obj
        json.dump(cfg, open(os.path.join(SCENES, dst + ".json"), "w"), indent=1)
    print("channel_circle: %d vertices, %d segments; box_sphere: %d vertices, %d triangles" % (len(v2), len(e2), len(v3), len(f3)))


if __name__ == "__main__":
    main()
