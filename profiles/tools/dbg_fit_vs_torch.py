"""Long-horizon check of one advection fit: DirectFit (fused kernels) vs torch autograd + torch.optim.Adam on stock ops,
same start, same hyper-parameters (lr 1e-5, 64^2 samples), loss trajectory and Taylor-Green error."""
import copy, math, os, sys
from importlib import import_module
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench_step, __graft_entry__ as ge
pkg = ge.load_package(); st = import_module(pkg.__name__ + ".stepper"); F = pkg.load_fields(); S = pkg.load_siren()
torch.manual_seed(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
s, cfg, init_fn, _, what = bench_step.build(SimpleNamespace(case="taylorgreen", iters=K, watertight=True, no_graph=True), pkg, st)
s.early_stop = False
s.fit_initial(init_fn, 5000, lr=1e-4)
size = s.size
print("init err %.4e" % float(F.taylor_green_error(s.velocity_field, size, 1000)))
s._sync_prev()
start = copy.deepcopy(s.velocity_field.state_dict())
# torch reference arrangement of _advect_velocity (model_split.py:88-120)
ref = S.FusedSiren(2, 2, 6, 64).cuda(); ref.load_state_dict(start)
prev = s.velocity_field_prev
opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
lo = torch.tensor(size[0::2], device="cuda"); hi = torch.tensor(size[1::2], device="cuda")
for it in range(K):
    x = s.sample_random(4096)
    with torch.no_grad():
        pu = S.envelope_reference(s.env, x, prev.forward_reference(x))
        back = torch.max(torch.min(x - pu*s.dt, hi), lo)
        tgt = S.envelope_reference(s.env, back, prev.forward_reference(back))
    loss = torch.mean((S.envelope_reference(s.env, x, ref.forward_reference(x)) - tgt)**2)
    opt.zero_grad(); loss.backward(); opt.step()
    if it in (0, 1, 2, 5, 10, 50, 100, 500, 1000, 2000, K - 1):
        print("torch  it %5d loss %.3e" % (it, loss.item()), flush=True)
print("torch  err after fit %.4e" % float(F.taylor_green_error(ref, size, 1000)))
# ours
s.max_n_iters = K
for chk in (1, 2, 3, 6, 11, 51, 101, 501, 1001, 2001, K):
    s.velocity_field.load_state_dict(start); s._sync_prev()
    it, loss = s.advect_velocity(chk)
    print("fused  it %5d loss %.3e  err %.4e" % (chk - 1, float(loss), float(F.taylor_green_error(s.velocity_field, size, 1000))), flush=True)
