"""Are the fit iterations (fp32 forward, tcgen05 delta chain / weight gradients, Adam) disturbed by a tcgen05 inference forward
running concurrently on another stream?  Same data, same start: weights after N iterations, quiet vs concurrent."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
pkg = util.package(); S = pkg.load_siren()
torch.manual_seed(0)
prev = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine", tensor_cores=True).cuda()
env = S.wall_envelope((0.0, 6.28, 0.0, 6.28), 1e-3)
xb = [torch.rand(4096, 2, device="cuda")*6.28 for _ in range(8)]
with torch.no_grad():
    tb = [prev(x, envelope=env) + 1e-3*torch.sin(x) for x in xb]
big = torch.rand(327680, 2, device="cuda")*6.28
side = torch.cuda.Stream()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400

def run(concurrent):
    net = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine", tensor_cores=True).cuda()
    net.load_state_dict(prev.state_dict())
    fit = S.DirectFit(net, 1e-5, env, max_batch=4096)
    torch.cuda.synchronize()
    for it in range(N):
        if concurrent and it % 4 == 0:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                y = prev(big, envelope=env); y = prev(big, envelope=env)
        fit.iterate(xb[it % 8], tb[it % 8])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return [p.detach().clone() for p in net.parameters()], fit.loss.item()

a, la = run(False); b, lb = run(False); c, lc = run(True); d, ld = run(True)
diff = lambda u, v: max((p - q).abs().max().item() for p, q in zip(u, v))
move = diff(a, [p.detach() for p in prev.parameters()])
print("moved by %.3e; quiet vs quiet %.3e; quiet vs concurrent %.3e, %.3e; losses %.4e %.4e %.4e %.4e" % (move, diff(a, b), diff(a, c), diff(a, d), la, lb, lc, ld))
