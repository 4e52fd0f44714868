"""One forward of the karman-shaped network (2 -> 128 -> [128]x2 -> 2) on the 1,004,004-point divergence grid through
the tcgen05 kernel, and one fit iteration at batch 16384 through the split fp32 kernels, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, __graft_entry__ as ge
pkg = ge.load_package(); S = pkg.load_siren()
net = S.FusedSiren(2, 2, 2, 128, nonlinearity="sine", tensor_cores=True).cuda()
x = torch.rand(1004004, 2, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = net(x)
fit = S.DirectFit(net, 1e-5, None, max_batch=16384)
xb = torch.rand(16384, 2, device="cuda"); tb = torch.rand(16384, 2, device="cuda")
for _ in range(3):
    fit.iterate(xb, tb)
torch.cuda.synchronize()
print("done")
# hidden = 64 at batch 16384: the one-kernel backward (sirenBackwardFusedTc) and the large-batch inference forward
net64 = S.FusedSiren(3, 3, 5, 64, nonlinearity="sine", tensor_cores=True).cuda()
fit64 = S.DirectFit(net64, 1e-5, None, max_batch=16384)
x64 = torch.rand(16384, 3, device="cuda"); t64 = torch.rand(16384, 3, device="cuda")
for _ in range(3):
    fit64.iterate(x64, t64)
torch.cuda.synchronize()
print("done 64")
