"""The Taylor-Green example to the end of its published run (examples/taylorgreen/run.sh: SIREN 6x64, lr 1e-5, dt 1e-3,
10000 Adam iterations per fit with early stop, 64^2 training samples, 512^2 pressure samples) through the device-resident
stepper, with the error metric of src/2d/tlgn_error.py (mean |u - u_TG|^2 on a 1000^2 grid) after every time step, next
to the curve the reference publishes (final_material/error_txt/error_ours.txt, kept as tests/golden/published_taylorgreen_error_ours.txt).
usage: taylor_green_run.py [steps] [shipped|active] [init_iters]
`shipped` keeps the scene as the reference ships it (isWatertight: true: every sample is classified outside, the
pressure gradient is zero -- SURVEY.md Appendix E); `active` runs the pressure projection."""
import json
import math
import os
import sys
import time
from importlib import import_module
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench_step  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
variant = sys.argv[2] if len(sys.argv) > 2 else "shipped"
init_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
pkg = ge.load_package()
st = import_module(pkg.__name__ + ".stepper")
F = pkg.load_fields()
torch.manual_seed(0)
s, cfg, init_fn, _, what = bench_step.build(SimpleNamespace(case="taylorgreen", iters=10000, watertight=variant == "shipped", no_graph=os.environ.get("NMC_NO_GRAPH", "0") == "1"), pkg, st)
s.early_stop = True
if os.environ.get("NMC_PREV_FP32", "0") == "1":   # experiment: the frozen network evaluated by the fp32 kernels at every batch size
    s.velocity_field_prev.tensor_cores = False
t0 = time.time()
s.fit_initial(init_fn, init_iters, lr=1e-5)  # add_source (base.py:330-335): 10000 iterations at the run's learning rate
size = s.size
published = np.loadtxt(os.path.join(ROOT, "tests", "golden", "published_taylorgreen_error_ours.txt"))
err = [float(F.taylor_green_error(s.velocity_field, size, 1000))]
print("initial fit: %d iterations, %.1f s, error %.4e (published curve starts at %.4e)" % (init_iters, time.time() - t0, err[0], published[0]), flush=True)
rows = []
for k in range(steps):
    t1 = time.time()
    info = s.step()
    torch.cuda.synchronize()
    e = float(F.taylor_green_error(s.velocity_field, size, 1000))
    err.append(e)
    pub = published[k + 1] if k + 1 < len(published) else float("nan")
    rows.append({"step": k + 1, "error": e, "published": pub, "advect_iters": info["advect_iters"], "project_iters": info["project_iters"],
                 "seconds": time.time() - t1})
    print("step %3d  error %.4e  published %.4e  ratio %.3f  (%d + %d iterations, %.2f s)" % (k + 1, e, pub, e/pub if pub == pub else float("nan"),
                                                                                    info["advect_iters"], info["project_iters"], time.time() - t1), flush=True)
n = min(len(err), len(published))
print(json.dumps({"variant": variant, "steps": steps, "mean_error": float(np.mean(err[:n])), "published_mean_error": float(np.mean(published[:n])),
                  "final_error": err[n - 1], "published_final_error": float(published[n - 1]),
                  "error_growth": err[n - 1] - err[0], "published_error_growth": float(published[n - 1] - published[0]), "config": what}))
s.close()
