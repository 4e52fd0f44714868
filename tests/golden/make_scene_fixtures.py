"""Regenerates tests/golden/scenes/* from the reference's example inputs (data, not code).

Run in the build container (where /root/reference exists):
    python tests/golden/make_scene_fixtures.py
Each OBJ keeps only its `v`, `l` and `f` records, byte-for-byte as in the example file, because the
loaders (bindings/zombie/demo/scene.h:104-145 for 2D, fcpw scene_loader.inl:101-150 for 3D) look at
nothing else. The solver/scene/output dictionaries are the shipped examples/*/wost.json with the
`boundary` path left relative (tests rewrite it to an absolute path).
"""
import json
import os

REF = "/root/reference/examples"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "scenes")

SCENES = {
    # fixture name -> (example dir, obj file, dim)
    "taylorgreen": ("taylorgreen", "square.obj", 2),
    "karman": ("karman", "geometry_1cyl_long_open.obj", 2),
    "smoke3d": ("smoke3d", "cube.obj", 3),  # identical OBJ in smoke_obs and vortex_collide
    "karman3d": ("karman3d", "cube.obj", 3),
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, (ex, obj, dim) in SCENES.items():
        with open(os.path.join(REF, ex, obj)) as f:
            keep = [ln for ln in f if ln.split(" ", 1)[0] in ("v", "l", "f")]
        with open(os.path.join(OUT, name + ".obj"), "w") as f:
            f.write("# geometry of examples/%s/%s (v/l/f records only)\n" % (ex, obj))
            f.writelines(keep)
        cfg = json.load(open(os.path.join(REF, ex, "wost.json")))
        cfg["dim"] = dim
        cfg["scene"]["boundary"] = name + ".obj"
        json.dump(cfg, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))
