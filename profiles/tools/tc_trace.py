"""Phase timeline of one tile of sirenForwardTc (CTA 0, thread 0; clock64 stamps of the -DNMC_TC_TRACE build, see
csrc/siren_tc.cu): usage  NMC_LIBNMCFS=<pkg>/build/variants/libnmcfs_trace.so python profiles/tools/tc_trace.py [H] [layers] [n]
tags: 1 tile start, 2 chunk start, 3 weights staged, 4 after fence + barrier, 5 MMAs issued + commit (thread 0),
6 next weights requested, 7 MMAs complete (mbarrier), 8 epilogues done, 9 tile end."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
pkg = ge.load_package()
S = pkg.load_siren()
net = S.FusedSiren(3, 3, layers, H, tensor_cores=True).cuda()
x = torch.rand(n, 3, device="cuda")*2 - 1
with torch.no_grad():
    for _ in range(3):
        y = net(x)
torch.cuda.synchronize()
L = pkg.capi.lib()
buf = (C.c_longlong*256)()
k = L.nmc_siren_trace_read(buf, 127)
names = {1: "tile start", 2: "chunk start", 3: "weights staged", 4: "fence+barrier", 5: "MMAs issued", 6: "next W requested", 7: "MMAs complete", 8: "epilogues done", 9: "tile end"}
t0 = buf[1]
prev = t0
print("H=%d layers=%d n=%d : %d stamps" % (H, layers, n, k))
for i in range(k):
    tag, t = buf[2*i], buf[2*i + 1]
    print("%-18s +%6d cycles  (total %7d)" % (names.get(tag, str(tag)), t - prev, t - t0))
    prev = t

# one fit iteration: the weight-gradient kernel's FMA-path CTA (slot 0: 1 tile start, 2 tiles done, 3 atomics done) and
# tensor-core CTA (slot 1: 1 start, 2 stage start, 3 previous MMAs done, 4 operands staged, 5 fence + barrier, 6 MMAs issued,
# 7 loop end, 8 MMAs complete, 9 reductions done)
fit = S.DirectFit(net, 1e-4, None, max_batch=n)
target = torch.sin(x)
for _ in range(3):
    fit.iterate(x, target)
torch.cuda.synchronize()
for slot in (0, 1, 2):  # 2 = delta chain: 1 tile start, 2 chunk start, 3 weights staged, 4 fence + barrier, 5 MMAs issued, 6 MMAs complete, 7 tile end
    k = L.nmc_siren_trace_read_wgrad(slot, buf, 127)
    print("weight-gradient kernel, slot %d: %d stamps" % (slot, k))
    t0 = prev = buf[1]
    for i in range(k):
        tag, t = buf[2*i], buf[2*i + 1]
        print("  tag %d  +%6d cycles  (total %7d)" % (tag, t - prev, t - t0))
        prev = t
