#!/usr/bin/env python
"""bench_siren.py -- the neural-field side of the hot path (input set M-MLP):
inference throughput on a divergence-grid-sized batch and the cost of one Adam iteration of the fit loops,
for the four network shapes of the shipped configs, fused kernels vs the reference's stock PyTorch ops.
Prints one JSON line per shape.  CUDA events, 5 warm-up iterations, inputs larger than L2 for the grid case."""
import argparse
import json
import sys
import os

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

SHAPES = {"taylorgreen": (2, 64, 6, 2, 4096), "karman": (2, 128, 2, 2, 16384), "smoke3d": (3, 64, 5, 3, 16384), "karman3d": (3, 128, 2, 3, 16384)}


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1004004)  # 1002 x 1002 divergence grid (model_split.py:230-243)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    S = ge.load_package().load_siren()
    torch.backends.cuda.matmul.allow_tf32 = False
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    for name, (i, h, l, o, batch) in SHAPES.items():
        if args.only and name != args.only:
            continue
        torch.manual_seed(0)
        net = S.FusedSiren(i, o, l, h, nonlinearity="sine").cuda()
        flops = 2*(i*h + l*h*h + h*o)  # per sample, forward
        xg = torch.rand(args.grid, i, device="cuda")*2 - 1
        with torch.no_grad():
            t_ref = timeit(lambda: net.forward_reference(xg), args.iters)
            t_f32 = timeit(lambda: net(xg), args.iters)
            net.tensor_cores = True
            t_tc = timeit(lambda: net(xg), args.iters)
            net.tensor_cores = False
        # one fit iteration of _advect_velocity (model_split.py:88-120): 2 no-grad forwards, 1 forward with grad, backward, Adam
        prev = S.FusedSiren(i, o, l, h, nonlinearity="sine").cuda()
        opt_f = S.FusedAdam(list(net.parameters()), lr=1e-5)
        net_r = S.FusedSiren(i, o, l, h, nonlinearity="sine").cuda()
        opt_r = torch.optim.Adam(net_r.parameters(), lr=1e-5)
        xb = torch.rand(batch, i, device="cuda")*2 - 1

        def it_fused():
            with torch.no_grad():
                pu = prev(xb); back = (xb - pu[:, :i]*0.05).clamp(-1, 1); adv = prev(back)
            loss = ((net(xb) - adv)**2).mean()
            opt_f.zero_grad(); loss.backward(); opt_f.step()

        def it_ref():
            with torch.no_grad():
                pu = prev.forward_reference(xb); back = (xb - pu[:, :i]*0.05).clamp(-1, 1); adv = prev.forward_reference(back)
            loss = ((net_r.forward_reference(xb) - adv)**2).mean()
            opt_r.zero_grad(); loss.backward(); opt_r.step()
        t_it_f = timeit(it_fused, 50)
        t_it_r = timeit(it_ref, 50)
        # the same iteration the way the time-stepper runs it (stepper.py): DirectFit (no autograd; tcgen05 forward /
        # delta chain / weight gradients, fused MSE and Adam), captured once in a CUDA graph and replayed
        net_g = S.FusedSiren(i, o, l, h, nonlinearity="sine", tensor_cores=True).cuda()
        fit = S.DirectFit(net_g, 1e-5, None, max_batch=batch)

        def it_direct():
            with torch.no_grad():
                pu = prev(xb); back = (xb - pu[:, :i]*0.05).clamp(-1, 1); adv = prev(back)
            fit.iterate(xb, adv)
        prev.tensor_cores = True
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                it_direct()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            it_direct()
        t_it_g = timeit(graph.replay, 200)
        del graph
        tf = lambda ms: args.grid*flops/(ms*1e-3)/1e12  # noqa: E731
        print(json.dumps({"shape": name, "net": {"in": i, "hidden": h, "hidden_layers": l, "out": o}, "grid_points": args.grid,
                          "forward_ms": {"torch_fp32": t_ref, "fused_fp32": t_f32, "fused_tcgen05_3xtf32": t_tc},
                          "forward_tflops_algorithmic": {"torch_fp32": tf(t_ref), "fused_fp32": tf(t_f32), "fused_tcgen05_3xtf32": tf(t_tc)},
                          "tensor_roofline": {"achieved_tflops_issued": 3*tf(t_tc), "peak_bf16_tflops": peaks["bf16_tflops"],
                                              "note": "3 TF32 MMAs per algorithmic product; TF32 peak is half the bf16 figure"},
                          "fit_iteration_ms": {"batch": batch, "torch": t_it_r, "fused_autograd_eager": t_it_f, "direct_fit_graph_replay": t_it_g}}), flush=True)


if __name__ == "__main__":
    main()
