import math, os, sys
from importlib import import_module
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench_step, __graft_entry__ as ge
pkg = ge.load_package(); st = import_module(pkg.__name__ + ".stepper"); F = pkg.load_fields()
torch.manual_seed(0)
variant = sys.argv[1] if len(sys.argv) > 1 else "shipped"
s, cfg, init_fn, _, what = bench_step.build(SimpleNamespace(case="taylorgreen", iters=10000, watertight=variant == "shipped", no_graph=True), pkg, st)
s.early_stop = False
s.fit_initial(init_fn, 5000, lr=1e-4)
size = s.size
print("init err %.4e" % float(F.taylor_green_error(s.velocity_field, size, 1000)))
for K in (100, 1000, 10000):
    snap = {k: v.clone() for k, v in s.velocity_field.state_dict().items()}
    s._sync_prev()
    it, loss = s.advect_velocity(K); la = float(loss)
    e1 = float(F.taylor_green_error(s.velocity_field, size, 1000))
    s._sync_prev()
    it2, loss2 = s.project_velocity(K); lp = float(loss2)
    e2 = float(F.taylor_green_error(s.velocity_field, size, 1000))
    gp = s.last["grad_p"]
    print("K=%5d advect loss %.3e err %.4e | project loss %.3e err %.4e | grad_p: mean|.| %.3e max %.3e nonzero %.3f, div grid rms %.3e" % (
        K, la, e1, lp, e2, gp.abs().mean().item(), gp.abs().max().item(), (gp != 0).float().mean().item(), s.last["div"].pow(2).mean().sqrt().item()), flush=True)
    s.velocity_field.load_state_dict(snap)
