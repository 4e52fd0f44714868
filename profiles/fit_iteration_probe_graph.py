"""A few graph-replayed advect and project fit iterations of the device-resident stepper (taylorgreen shape), as shipped: batches
and targets from the chunked generator, single-stream iteration.  `ncu --metrics gpu__time_duration.sum` lists the kernels
(profiles/r02_fit_iteration_launches_final.csv)."""
import math, os, sys
from importlib import import_module
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util, __graft_entry__ as ge
pkg = ge.load_package()
st = import_module(pkg.__name__ + ".stepper")
s = st.SplitStepper(util.load_case("taylorgreen_active"), scene_size=(0.0, 2*math.pi, 0.0, 2*math.pi), grid_resolution=200,
                    wost_resolution=64, max_n_iters=40, early_stop=False, use_cuda_graph=True, seed=1)
s._sync_prev()
s.advect_velocity(40)   # 3 warm-up + 17 single-iteration replays + one replay of the 20-iteration graph
s._sync_prev()
s.project_velocity(40)
torch.cuda.synchronize()
print("done")
