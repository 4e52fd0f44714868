"""Backward pass of a fit iteration, H = 64 networks: the two-kernel tcgen05 form (delta chain + weight gradients) against the
one-kernel form (csrc/siren_tc_fused_bwd.cu).  CUDA events, 10 warm-up launches, 100 timed."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
pkg = util.package(); S = pkg.load_siren(); L = S._lib()
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for name, (i, h, l, o, n) in {"taylorgreen": (2, 64, 6, 2, 4096), "smoke3d": (3, 64, 5, 3, 16384), "tg16k": (2, 64, 6, 2, 16384)}.items():
    if only and name != only:
        continue
    torch.manual_seed(0)
    net = S.FusedSiren(i, o, l, h, nonlinearity="sine", tensor_cores=True).cuda()
    lin = net._linears()
    W = [m.weight.detach().contiguous() for m in lin]; b = [m.bias.detach().contiguous() for m in lin]
    sh = S._shape_of(W, 30.0)
    x = torch.rand(n, i, device="cuda")*2 - 1
    z = torch.empty(((l + 1)*h, n), device="cuda"); y = torch.empty((n, o), device="cuda")
    S._check(L.nmc_siren_forward(C.byref(sh), S._ptrs(W), S._ptrs(b), x.data_ptr(), n, y.data_ptr(), z.data_ptr(), None, S._stream()))
    gy = (y*(2.0/y.numel())).contiguous()
    gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(v) for v in b]
    dZ = torch.empty(((l + 1)*h + o)*n, device="cuda")

    def two():
        S._check(L.nmc_siren_backward_tc(C.byref(sh), S._ptrs(W), S._ptrs(b), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(), dZ.data_ptr(), None, S._stream()))
        S._check(L.nmc_siren_weight_grads_tc(C.byref(sh), x.data_ptr(), n, dZ.data_ptr(), z.data_ptr(), S._ptrs(gW), S._ptrs(gb), S._stream()))

    def one():
        S._check(L.nmc_siren_backward_fused_tc(C.byref(sh), S._ptrs(W), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(), S._ptrs(gW), S._ptrs(gb), None, S._stream()))

    res = {}
    for key, fn in (("two_kernels_us", two), ("fused_us", one)):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[key] = e0.elapsed_time(e1)/reps*1e3
    print(json.dumps({"shape": name, "batch": n, **res}), flush=True)
