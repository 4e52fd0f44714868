"""Phase timeline of the first tile of sirenBackwardFusedTc (CTA 0; thread 0 = the MMA issuer, thread 64 = an epilogue-only
warp; clock64 stamps of the -DNMC_TC_TRACE build of csrc/siren_tc_fused_bwd.cu).
usage: NMC_LIBNMCFS=<pkg>/build/variants/libnmcfs_ftrace.so python profiles/tools/fused_bwd_trace.py [layers] [n]
tags: 1 tile start, 2 layer-L deltas computed, 3 stored (+ bias sums), 4 last-layer batch issued, per hidden layer: 10 loop top,
11 gradient batch of the previous layer complete, 12 operands stored + fences, 13 barrier, 14 chain + gradient batches issued,
15 chain complete, 16 TMEM read + cosine factor, 17 K-major store, 18 bias sums, 19 next activations; 20 loop end, 21 tile end,
22 gradient tiles written."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
pkg = util.package(); S = pkg.load_siren(); L = S._lib()
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
net = S.FusedSiren(3, 3, layers, 64, nonlinearity="sine", tensor_cores=True).cuda()
lin = net._linears()
W = [m.weight.detach().contiguous() for m in lin]; b = [m.bias.detach().contiguous() for m in lin]
sh = S._shape_of(W, 30.0)
x = torch.rand(n, 3, device="cuda")*2 - 1
z = torch.empty(((layers + 1)*64, n), device="cuda"); y = torch.empty((n, 3), device="cuda")
S._check(L.nmc_siren_forward(C.byref(sh), S._ptrs(W), S._ptrs(b), x.data_ptr(), n, y.data_ptr(), z.data_ptr(), None, S._stream()))
gy = (y*(2.0/y.numel())).contiguous()
gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(v) for v in b]
for _ in range(3):
    S._check(L.nmc_siren_backward_fused_tc(C.byref(sh), S._ptrs(W), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(), S._ptrs(gW), S._ptrs(gb), None, S._stream()))
torch.cuda.synchronize()
buf = (C.c_longlong*256)()
for slot in (0, 1):
    k = L.nmc_siren_trace_read_fused(slot, buf, 127)
    print("thread %d: %d stamps" % (0 if slot == 0 else 64, k))
    t0 = prev = buf[1]
    for i in range(k):
        tag, t = buf[2*i], buf[2*i + 1]
        print("  tag %2d  +%6d cycles  (total %7d)" % (tag, t - prev, t - t0))
        prev = t
