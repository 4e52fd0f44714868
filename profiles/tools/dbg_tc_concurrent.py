"""Is the tcgen05 inference forward exact when it runs concurrently with fit iterations on another stream?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
pkg = util.package(); S = pkg.load_siren()
torch.manual_seed(0)
prev = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine", tensor_cores=True).cuda()
net = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine", tensor_cores=True).cuda()
env = S.wall_envelope((0.0, 6.28, 0.0, 6.28), 1e-3)
fit = S.DirectFit(net, 1e-5, env, max_batch=4096)
xb = torch.rand(4096, 2, device="cuda")*6.28; tb = torch.rand(4096, 2, device="cuda")
side = torch.cuda.Stream()
for m in (81920, 163840, 327680):
    x = torch.rand(m, 2, device="cuda")*6.28
    with torch.no_grad():
        prev.tensor_cores = False
        ref = prev(x, envelope=env)
        prev.tensor_cores = True
        quiet = prev(x, envelope=env)
        torch.cuda.synchronize()
        worst = 0.0; bad = 0
        for rep in range(20):
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                y = prev(x, envelope=env)
            for _ in range(12):
                fit.iterate(xb, tb)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            dmax = (y - quiet).abs().max().item()
            worst = max(worst, dmax); bad += int(dmax > 0)
    print("batch %6d: |tc - fp32| max %.3e   concurrent vs quiet tc: max %.3e, %d of 20 runs differ" % (m, (quiet - ref).abs().max().item(), worst, bad), flush=True)
