import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
import test_gpu_siren as T
pkg = util.package(); siren = pkg.load_siren()
shape = (2, 64, 3, 2); n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = T._net(siren, shape, seed=91); x = T._coords(n, shape[0], seed=92)
target = torch.sin(3*x[:, :1]).expand(-1, shape[3]).contiguous()
with torch.no_grad():
    y = net(x)
gy = ((y - target)*(2.0/y.numel())).contiguous()
_, gW32, gb32 = T._raw_backward(siren, net, x, gy, None, tc=False)
L = siren._lib(); lin = net._linears()
W = [m.weight.detach().contiguous() for m in lin]; b = [m.bias.detach().contiguous() for m in lin]
sh = siren._shape_of(W, 30.0)
z = torch.empty(((sh.n_hidden_layers + 1)*sh.hidden, n), device=x.device); yy = torch.empty((n, sh.out_dim), device=x.device)
siren._check(L.nmc_siren_forward(C.byref(sh), siren._ptrs(W), siren._ptrs(b), x.data_ptr(), n, yy.data_ptr(), z.data_ptr(), None, siren._stream()))
gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(v) for v in b]
siren._check(L.nmc_siren_backward_fused_tc(C.byref(sh), siren._ptrs(W), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(), siren._ptrs(gW), siren._ptrs(gb), None, siren._stream()))
torch.cuda.synchronize()
for l, (a, r) in enumerate(zip(gW, gW32)):
    print("gW[%d] max|diff| %.3e  max|ref| %.3e  max|ours| %.3e  corr %.4f" % (l, (a - r).abs().max().item(), r.abs().max().item(), a.abs().max().item(),
          float(torch.corrcoef(torch.stack([a.flatten(), r.flatten()]))[0, 1]) if a.abs().max() > 0 else 0.0))
    if 1 <= l <= sh.n_hidden_layers:
        at = a.t()
        print("        vs transposed ref: max|diff| %.3e" % (a - r.t()).abs().max().item())
for l, (a, r) in enumerate(zip(gb, gb32)):
    print("gb[%d] max|diff| %.3e  max|ref| %.3e  max|ours| %.3e" % (l, (a - r).abs().max().item(), r.abs().max().item(), a.abs().max().item()))
if os.environ.get("NMC_DUMP"):
    d = gW[1].flatten().reshape(128, 32)
    ref = gW32[1]
    print("dump: nonzero lanes", [int(i) for i in torch.nonzero(d.abs().sum(1) > 0).flatten().tolist()][:40])
    print("dump lane0 cols0-7", d[0, :8].tolist())
    print("ref row0 cols0-7  ", ref[0, :8].tolist())
    print("ref col0 rows0-7  ", ref[:8, 0].tolist())
