"""A/B of walk-kernel builds on the GPU box: for every libnmcfs_*.so under build/variants (or the names given), runs the
bench workloads in a fresh process (NMC_LIBNMCFS selects the library) and prints walks/s.
usage: python profiles/tools/ab_walk.py [name ...]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vdir = os.path.join(ROOT, "neural-monte-carlo-fluid-simulation_b200", "build", "variants")
names = sys.argv[1:] or sorted(os.path.basename(f)[len("libnmcfs_"):-3] for f in glob.glob(os.path.join(vdir, "libnmcfs_*.so")))
WL = ["smoke3d_1000000pts_x500walks", "karman_100000pts_x500walks", "karman3d_1000000pts_x500walks"]
for rep in range(2):
    for name in names:
        env = dict(os.environ, NMC_LIBNMCFS=os.path.join(vdir, "libnmcfs_%s.so" % name))
        row = []
        for wl in WL:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", "5", "--warmup", "3", "--no-sim-steps",
                                "--no-python-e2e", "--no-cpu-baseline", "--no-also"], capture_output=True, text=True, env=env)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                row.append("%s %.4g (occ %.3f)" % (wl.split("_")[0], d["value"], d["config"]["lane_occupancy_of_slice_loop"] or 0))
            except Exception:
                row.append("%s FAILED %s" % (wl, r.stderr[-300:]))
        print("%-14s rep %d | %s" % (name, rep, " | ".join(row)), flush=True)
