"""Run under torchrun with 2+ GPUs (tests/test_gpu_stepper.py launches it when they are there): the distributed
stepper keeps the ranks' networks identical, gathers every rank's pressure samples, and its fit follows the
single-GPU fit on the union of the ranks' batches (same samples fed to both)."""
import math
import os
import sys
from importlib import import_module

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import util  # noqa: E402


def say(rank, msg):
    print("[rank %d] %s" % (rank, msg), flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("NMC_DIST_CHECK_TIMEOUT", "150")), exit=True)  # a hung collective must not eat the GPU budget
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = util.package()
    st = import_module(pkg.__name__ + ".stepper")
    S = pkg.load_siren()
    cfg = util.load_case("taylorgreen_active")
    kw = dict(scene_size=pkg.workloads.scene_size_from_obj(cfg["scene"]["boundary"]), grid_resolution=120, wost_resolution=64, sample_resolution=32, max_n_iters=30,
              check_every=10, lr=1e-4, seed=3, device=local, reset_wts=True)
    tg = lambda x: torch.stack([torch.sin(x[:, 0])*torch.cos(x[:, 1]), -torch.cos(x[:, 0])*torch.sin(x[:, 1])], dim=-1)  # noqa: E731

    say(rank, 'process group up')
    # 1. data-parallel DirectFit == single-process DirectFit on the concatenated batch
    torch.manual_seed(100)
    net_a = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine").cuda()
    net_b = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine").cuda()
    net_b.load_state_dict(net_a.state_dict())
    fit_a = S.DirectFit(net_a, 1e-4, None, max_batch=1024, distributed=True)
    fit_b = S.DirectFit(net_b, 1e-4, None, max_batch=1024*world)
    fit_a.sync_parameters()
    g = torch.Generator(device="cuda").manual_seed(7)
    for _ in range(20):
        x_all = torch.rand(1024*world, 2, device="cuda", generator=g)*6.28
        t_all = tg(x_all)
        fit_a.iterate(x_all[rank*1024:(rank + 1)*1024].contiguous(), t_all[rank*1024:(rank + 1)*1024].contiguous())
        fit_b.iterate(x_all, t_all)
    for a, b in zip(net_a.parameters(), net_b.parameters()):
        assert (a - b).abs().max().item() <= 2e-3*1e-4*20 + 1e-7, (a - b).abs().max().item()   # << the 20 lr-sized steps taken

    say(rank, 'data-parallel fit matches')
    # 2. the distributed step: identical weights on every rank, gathered pressure samples, finite losses
    for graph in (False, True):
        s = st.SplitStepper(cfg, use_cuda_graph=graph, distributed=True, **kw)
        say(rank, 'stepper built graph=%s' % graph)
        s.fit_initial(tg, 50, lr=1e-3)
        say(rank, 'initial fit done')
        out = s.step()
        say(rank, 'step done')
        assert math.isfinite(out["advect_loss"].item()) and math.isfinite(out["project_loss"].item())
        assert s.last["pressure_samples"].shape[0] == 64*64 and s.last["grad_p"].shape == (64*64, 2)
        flat = torch.cat([p.detach().reshape(-1) for p in s.velocity_field.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(flat, ref), "rank %d diverged from rank 0 (graph=%s)" % (rank, graph)
        ps = s.last["pressure_samples"].clone()
        dist.broadcast(ps, src=0)
        assert torch.equal(ps, s.last["pressure_samples"])
        # the ranks trained on different samples: their shards of the gathered pressure set differ
        assert not torch.equal(s.last["pressure_samples"][:2048], s.last["pressure_samples"][2048:4096])
        # early stop is a collective decision: rank 0 alone below the threshold must not stop anybody
        mine = torch.tensor(1e-12 if rank == 0 else 1.0, device="cuda")
        assert s._stop_now(mine) is False
        assert s._stop_now(torch.tensor(1e-12, device="cuda")) is True
        s.close()   # drops the graphs that captured the gradient all_reduce
    # 3. replicated fits (fit_parallel="replicated"): every rank runs the whole fit on identical samples, only the
    #    pressure solve is sharded; identical weights (re-broadcast) and identical full pressure sets on every rank
    s = st.SplitStepper(cfg, use_cuda_graph=True, distributed=True, fit_parallel="replicated", **kw)
    s.fit_initial(tg, 50, lr=1e-3)
    for _ in range(2):
        out = s.step()
    assert math.isfinite(out["advect_loss"].item()) and math.isfinite(out["project_loss"].item())
    assert s.last["pressure_samples"].shape[0] == 64*64 and s.last["grad_p"].shape == (64*64, 2)
    for t in (torch.cat([p.detach().reshape(-1) for p in s.velocity_field.parameters()]), s.last["pressure_samples"], s.last["grad_p"]):
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "replicated fits: rank %d differs from rank 0" % rank
    assert s.last["grad_p"].abs().sum().item() > 0
    s.close()
    say(rank, 'replicated fits ok')
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("DIST_STEPPER_OK world=%d" % world, flush=True)
    dist.destroy_process_group()   # returns because close() released the NCCL-capturing graphs
    if rank == 0:
        print("DIST_STEPPER_CLEAN_EXIT", flush=True)


if __name__ == "__main__":
    main()
