"""Device-resident split step (stepper.py) against the pieces it replaces: divergence grid vs a stock-PyTorch
evaluation, device-resident pressure solve vs the host-buffer entry point, and the fit loops (eager and
CUDA-graph replay) driving the loss down."""
import math
from importlib import import_module

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stepper(**kw):
    pkg = util.package()
    st = import_module(pkg.__name__ + ".stepper")
    cfg = util.load_case("taylorgreen_active")
    args = dict(scene_size=pkg.workloads.scene_size_from_obj(cfg["scene"]["boundary"]), grid_resolution=200, wost_resolution=64, sample_resolution=32,
                max_n_iters=40, check_every=10, seed=3, device=0)
    args.update(kw)
    return pkg, st.SplitStepper(cfg, **args)


def _tg(x):  # Taylor-Green vortex
    return torch.stack([torch.sin(x[:, 0])*torch.cos(x[:, 1]), -torch.cos(x[:, 0])*torch.sin(x[:, 1])], dim=-1)


def test_divergence_grid_matches_stock_autograd():
    pkg, s = _stepper(use_cuda_graph=False)
    div = s.divergence_grid()
    assert tuple(div.shape) == tuple(s.grid_shape) and 201 <= min(div.shape) and max(div.shape) == 202  # sample_uniform_2D: int(200 * ratio) + 2 along the shorter side
    x = s.grid_samples.detach().clone().requires_grad_(True)
    u = s.velocity_field_prev.forward_reference(x)*s.envelope(x)
    ref = 0.0
    for i in range(2):
        ref = ref + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=True)[0][:, i]
    ref = (-ref).reshape(div.shape)
    assert (div - ref).abs().max().item() <= 3e-4*ref.abs().max().item() + 1e-7
    # orientation: rows <-> y, columns <-> x (image.h:70-75): grid_samples[i*W + j] = (x_j, y_i)
    g = s.grid_samples.reshape(div.shape[0], div.shape[1], 2)
    assert g[0, 5, 0] > g[0, 4, 0] and g[5, 0, 1] > g[4, 0, 1] and g[0, 5, 1] == g[0, 4, 1]


def test_device_resident_pressure_solve_equals_host_entry_point():
    pkg = util.package()
    pkg, s = _stepper(use_cuda_graph=False, mode=pkg.capi.MODE_DETERMINISTIC)
    pts = s.sample_random(2000).contiguous()
    p, g = s.pressure_solve(pts)
    div = s.last["div"].cpu().numpy()
    sc = pkg.Scene(s.cfg["scene"], div, device=0)
    ph, gh, _, _ = pkg.zombie.wost_array(sc, s.cfg["solver"], s.cfg["output"], pts.cpu().numpy(), mode=pkg.capi.MODE_DETERMINISTIC, seed=3)
    assert np.array_equal(p.cpu().numpy(), ph) and np.array_equal(g.cpu().numpy(), gh)


@pytest.mark.parametrize("graph", [False, True])
def test_step_runs_and_fits(graph):
    pkg, s = _stepper(use_cuda_graph=graph, max_n_iters=60, lr=1e-4)
    s.fit_initial(_tg, 150, lr=2e-3)
    x = s.sample_random(4096)
    before = torch.mean((s.query_velocity(x) - _tg(x))**2).item()
    out = s.step()
    assert out["advect_iters"] == 60 and out["project_iters"] == 60
    assert math.isfinite(out["advect_loss"].item()) and math.isfinite(out["project_loss"].item())
    assert s.last["walks"] > 0 and torch.isfinite(s.last["grad_p"]).all()
    after = torch.mean((s.query_velocity(x) - _tg(x))**2).item()
    assert math.isfinite(after) and after < 10*max(before, 1e-3)  # one small step keeps the field near Taylor-Green
    _, s2 = _stepper(use_cuda_graph=graph, max_n_iters=5, lr=1e-4)
    s2.fit_initial(_tg, 150, lr=2e-3)
    s2._sync_prev()
    _, l5 = s2.advect_velocity(5)
    _, l200 = s2.advect_velocity(200)
    assert l200.item() <= l5.item()*1.5


def _reference_arrangement_step(pkg, s, net, prev, lr, oracle_lib, n_iters, seed):
    """One NeuralFluidSplit.step (model_split.py:44-62, adv_ref = 0) the way the reference runs it: stock PyTorch
    ops for the network (forward_reference = nn.Linear + sin), torch.optim.Adam, autograd divergence, the
    divergence grid moved to the host, the CPU solver (oracle, bit-exact with the reference's), grad p moved back."""
    env = s.envelope
    n = s.sample_resolution**2
    prev.load_state_dict(net.state_dict())
    opt = torch.optim.Adam(net.parameters(), lr=lr)  # create_optimizer() at the top of every _training_loop (base.py:133)
    for _ in range(n_iters):  # _advect_velocity, model_split.py:88-120
        x = s.sample_random(n)
        with torch.no_grad():
            pu = prev.forward_reference(x)*env(x)
            back = torch.clamp(x - pu*s.dt, min=s._lo, max=s._hi)
            adv = prev.forward_reference(back)*env(back)
        loss = torch.mean((net.forward_reference(x)*env(x) - adv)**2)
        opt.zero_grad(); loss.backward(); opt.step()
    prev.load_state_dict(net.state_dict())
    x = s.grid_samples.detach().clone().requires_grad_(True)  # get_divergence, model_split.py:230-243
    u = prev.forward_reference(x)*env(x)
    div = 0.0
    for i in range(2):
        div = div + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=(i == 0))[0][:, i]
    div = (-div).reshape(s.grid_shape).detach().cpu().numpy()
    pts = s.sample_random(s.wost_resolution**2)
    osc = oracle_lib.OracleScene(2, s.cfg["scene"], div)  # wost_pressure, model_split.py:185-228
    _, g, _ = osc.wost(s.cfg["solver"], s.cfg["output"], pts.cpu().numpy(), seed=seed, nthreads=8)
    grad_p = torch.from_numpy(g).to(pts.device)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    for _ in range(n_iters):  # _project_velocity, model_split.py:246-284
        idx = torch.randint(0, pts.shape[0] - 1, (n,), device=pts.device)
        xs = pts[idx]
        with torch.no_grad():
            target = prev.forward_reference(xs)*env(xs) - grad_p[idx]
        loss = torch.mean((net.forward_reference(xs)*env(xs) - target)**2)
        opt.zero_grad(); loss.backward(); opt.step()
    prev.load_state_dict(net.state_dict())


def test_taylor_green_velocity_error_tracks_the_reference_arrangement(oracle_lib):
    """north_star, second criterion: the end-of-run Taylor-Green velocity error (move_density.py:143-146 metric)
    of the device-resident stepper stays within 1 % of the reference arrangement's, from the same initial
    network, on a reduced configuration (solver-active scene, 3 time steps).
    Part A removes the sampling noise so that 1 % is a meaningful bar: both arrangements draw the same training
    samples (same torch seed, eager launches) and the pressure solve runs in deterministic mode with the seed
    the CPU solver gets, so the only differences left are the fused kernels' rounding and the solver's 1e-5.
    Part B is the production setting (default mode, CUDA-graph replay): different samples and walks, so the
    errors agree only up to the run-to-run spread of either arrangement (measured: +-35 % at this size)."""
    lr, K, steps = 2e-5, 120, 3
    kw = dict(max_n_iters=K, lr=lr, dt=0.01, grid_resolution=300, wost_resolution=96, sample_resolution=48, early_stop=False, seed=11)
    pkg, s = _stepper(use_cuda_graph=False, mode=util.package().capi.MODE_DETERMINISTIC, **kw)
    F = pkg.load_fields(); S = pkg.load_siren()
    size = s.size
    s.fit_initial(_tg, 2500, lr=3e-4)
    init = {k: v.clone() for k, v in s.velocity_field.state_dict().items()}
    ref_net = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine").cuda(); ref_prev = S.FusedSiren(2, 2, 6, 64, nonlinearity="sine").cuda()
    ref_net.load_state_dict(init)
    e0 = F.taylor_green_error(s.velocity_field, size, 500)
    assert e0 < 2e-3, "initial fit did not converge: %g" % e0
    assert e0 == pytest.approx(F.taylor_green_error(ref_net, size, 500), rel=1e-3)   # same weights, fused vs stock evaluation
    torch.manual_seed(77)
    ours, seeds = [], []
    for t in range(steps):
        seeds.append(int(s.opts.seed))
        s.step()
        ours.append(F.taylor_green_error(s.velocity_field, size, 500))
    torch.manual_seed(77)
    ref = []
    for t in range(steps):
        _reference_arrangement_step(pkg, s, ref_net, ref_prev, lr, oracle_lib, K, seed=seeds[t])
        ref.append(F.taylor_green_error(ref_net, size, 500))
    print("taylor-green velocity error, same samples: initial %.4e, ours %s, reference arrangement %s" % (e0, ours, ref))
    for a, b in zip(ours, ref):
        assert abs(a - b) <= 0.01*b, (ours, ref)
    assert abs(ref[-1] - e0) > 0.05*e0   # the steps moved the field by far more than the tolerance

    # Part B: production setting from the same initial network
    pkg, s2 = _stepper(use_cuda_graph=True, **kw)
    s2.velocity_field.load_state_dict(init); s2._sync_prev()
    fast = []
    for t in range(steps):
        s2.step()
        fast.append(F.taylor_green_error(s2.velocity_field, size, 500))
    print("default mode + graph replay:", fast)
    assert all(math.isfinite(e) for e in fast)
    assert 0.5*ref[-1] <= fast[-1] <= 2.0*ref[-1], (fast, ref)


def test_chunked_fit_targets_are_the_targets_of_their_batches():
    """Graph mode: batches and targets of the next `target_chunk` iterations are computed in one pass on a second stream into a
    ring of two chunks and fetched by the captured iteration (stepper._loop, chunked=...).  Every ring slot holds a batch
    inside the domain together with the target of exactly that batch, slots differ from each other and from fit to fit, the
    iteration reads the slot of its index, and the fit reaches the loss level of the fit that computes its targets inside
    the iteration (NMC_TARGET_CHUNK=0 arrangement)."""
    kw = dict(max_n_iters=40, lr=1e-4, dt=0.01, grid_resolution=120, wost_resolution=48, sample_resolution=64, early_stop=False, seed=3)
    pkg, s = _stepper(use_cuda_graph=True, **kw)
    s.target_chunk = 8
    s.fit_initial(_tg, 300, lr=3e-4)
    F = pkg.load_fields()
    lo = torch.tensor(s.size[0::2], device="cuda"); hi = torch.tensor(s.size[1::2], device="cuda")

    def expected_target(x):
        with torch.no_grad():
            pu = s.query_velocity(x, use_prev=True)
            return s.query_velocity(F.backtrace(x, pu, s.dt, s.size[0::2], s.size[1::2]), use_prev=True)
    seen = []
    for n_it in (7, 21):   # within one chunk; three chunks through both halves of the ring
        s._sync_prev()
        it, loss = s.advect_velocity(n_it)
        torch.cuda.synchronize()
        assert it == n_it and math.isfinite(loss.item()) and s._fit.opt.step_dev.item() == n_it
        ring = s._rings["advect"]
        last = n_it - 1
        assert torch.equal(ring["x"], ring["X"][last % 16]) and torch.equal(ring["t"], ring["T"][last % 16])   # the slot of the last iteration
        for k in range(8*(last//8), last + 1):                       # the chunk in use
            x, t = ring["X"][k % 16], ring["T"][k % 16]
            assert (x >= lo).all() and (x < hi).all()
            assert (t - expected_target(x)).abs().max().item() <= 1e-5
        assert (ring["X"][last % 16] == ring["X"][(last - 1) % 16]).float().mean().item() < 1e-3
        seen.append(ring["x"].clone())
    assert (seen[0] == seen[1]).float().mean().item() < 1e-3
    # projection fit: the target is (u_prev(x), grad p) with x, grad p rows of the pressure samples
    s._sync_prev()
    it, loss = s.project_velocity(12)
    torch.cuda.synchronize()
    ring = s._rings["project"]
    ps, pg = s.last["pressure_samples"], s.last["grad_p"]
    x, t, g = ring["X"][3], ring["T"][3], ring["S"][3]
    with torch.no_grad():
        assert (t - s.query_velocity(x, use_prev=True)).abs().max().item() <= 1e-5
    match = (x[:64, None, :] == ps[None, :, :]).all(dim=2)             # every fetched sample is one of the pressure samples, with its gradient
    assert match.any(dim=1).all()
    assert torch.equal(g[:64], pg[match.float().argmax(dim=1)])
    # same fit with the targets computed inside the iteration: different batches, same loss level
    s2_pkg, s2 = _stepper(use_cuda_graph=True, **kw)
    s2.target_chunk = 0
    s2.velocity_field.load_state_dict(s.velocity_field_prev.state_dict()); s2._sync_prev()
    s.velocity_field.load_state_dict(s.velocity_field_prev.state_dict())
    _, la = s.advect_velocity(40); _, lb = s2.advect_velocity(40)
    assert "advect" not in s2._rings
    assert 0.5 <= la.item()/lb.item() <= 2.0, (la.item(), lb.item())


def _karman_stepper(**kw):
    pkg = util.package()
    st = import_module(pkg.__name__ + ".stepper")
    cfg = util.load_case("karman")
    centre, radius, size = util.karman_obstacle(cfg["output"]["boundaryDistanceMask"])
    args = dict(scene_size=size, hidden_features=128, num_hidden_layers=2, dt=0.05, lr=1e-5, grid_resolution=250, wost_resolution=64,
                sample_resolution=32, bdry_eps=3e-2, max_n_iters=40, check_every=10, boundary="karman", obstacle=(centre, radius),
                karman_vel=0.5, seed=5, device=0)
    args.update(kw)
    return pkg, st.SplitStepper(cfg, **args)


@pytest.mark.parametrize("graph", [False, True])
def test_karman_configuration_steps(graph):
    """examples/karman shape (SIREN 2x128, inlet strip + no-slip cylinder + wall envelope, --reset_wts 1, samples
    inside the cylinder unused): the divergence grid equals stock autograd through the reference's query_velocity
    (the obstacle weight is differentiated, base.py:352-358), and a step runs end to end on the device."""
    pkg, s = _karman_stepper(use_cuda_graph=graph, reset_wts=True)
    s.fit_initial(s.karman_initial_velocity, 200, lr=1e-3)
    div = s.divergence_grid()
    x = s.grid_samples.detach().clone().requires_grad_(True)
    u = s.apply_envelope_reference(x, s.velocity_field_prev.forward_reference(x))
    ref = 0.0
    for i in range(2):
        ref = ref + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=True)[0][:, i]
    ref = (-ref).reshape(s.grid_shape)
    assert tuple(div.shape) == tuple(ref.shape) and div.shape[1] > div.shape[0]   # channel: more columns (x) than rows (y)
    assert (div - ref).abs().max().item() <= 5e-4*ref.abs().max().item() + 1e-6
    xs = s.sample_random(20000)
    assert xs.shape == (20000, 2) and (s.obstacle_distance(xs) > 0).float().mean().item() > 0.9999
    assert s.sample_random(20000, keep_shape=False).shape[0] < 20000
    w0 = s.velocity_field.net[0].weight.detach().clone()
    out = s.step()
    assert out["advect_iters"] == 40 and out["project_iters"] == 40
    assert math.isfinite(out["advect_loss"].item()) and math.isfinite(out["project_loss"].item())
    assert (s.obstacle_distance(s.last["pressure_samples"]) > 0).all() and s.last["pressure_samples"].shape[0] < 64*64
    assert s.last["walks"] > 0 and torch.isfinite(s.last["grad_p"]).all()
    assert (s.velocity_field.net[0].weight.detach() - w0).abs().max().item() > 1e-3   # reset_wts: re-initialised before each fit
    inlet = torch.tensor([[s.size[0] + 0.5*s.eps, 0.5*(s.size[2] + s.size[3])]], device="cuda")
    assert s.query_velocity(inlet)[0, 0].item() == pytest.approx(0.5, rel=1e-5)   # inlet strip: u = karman_vel (far from the cylinder)


@pytest.mark.parametrize("boundary", ["walls", "smoke_obs"])
def test_3d_configuration_steps(boundary):
    """examples/smoke3d / vortex_collide (wall envelope) and smoke_obs (inlet ball + no-slip sphere + walls) shapes:
    SIREN 5x64 3->3, 82^3-style divergence grid laid out [x][y][z] for zombie3d, 3D pressure solve on the device."""
    pkg = util.package()
    st = import_module(pkg.__name__ + ".stepper")
    cfg = util.load_case("smoke3d")
    kw = dict(scene_size=(-1.0, 1.0)*3, hidden_features=64, num_hidden_layers=5, dt=0.05, lr=1e-5, grid_resolution=30, wost_resolution=48,
              sample_resolution=32, bdry_eps=1e-2, max_n_iters=30, check_every=10, boundary=boundary, reset_wts=True, seed=7, device=0)
    if boundary == "smoke_obs":
        kw["obstacle"] = ((0.0, 0.0, -0.3), 0.1)   # src/3d/main.py:85-91
    s = st.SplitStepper(cfg, **kw)
    swirl = lambda x: torch.stack([-x[:, 1], x[:, 0], 0.2*torch.ones_like(x[:, 0])], dim=-1)*0.3  # noqa: E731
    s.fit_initial(swirl, 150, lr=1e-3)
    div = s.divergence_grid()
    assert tuple(div.shape) == (32, 32, 32)
    x = s.grid_samples.detach().clone().requires_grad_(True)
    u = s.apply_envelope_reference(x, s.velocity_field_prev.forward_reference(x))
    ref = 0.0
    for i in range(3):
        ref = ref + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=True)[0][:, i]
    ref = (-ref).reshape(32, 32, 32)
    assert (div - ref).abs().max().item() <= 5e-4*ref.abs().max().item() + 1e-6
    g = s.grid_samples.reshape(32, 32, 32, 3)   # [x][y][z]: the first index moves x only
    assert g[5, 0, 0, 0] > g[4, 0, 0, 0] and g[5, 0, 0, 1] == g[4, 0, 0, 1] and g[0, 0, 5, 2] > g[0, 0, 4, 2]
    out = s.step()
    assert out["advect_iters"] == 30 and out["project_iters"] == 30
    assert math.isfinite(out["advect_loss"].item()) and math.isfinite(out["project_loss"].item())
    assert s.last["grad_p"].shape == (48*48, 3) and torch.isfinite(s.last["grad_p"]).all()
    assert 0.95*500*48*48 < s.last["walks"] <= 500*48*48   # points inside the boundary mask are not walked
    # device-resident 3D solve == host entry point on the same divergence grid (deterministic mode)
    s.opts.mode = pkg.capi.MODE_DETERMINISTIC
    pts = s.sample_random(500).contiguous()
    p, gp = s.pressure_solve(pts)
    sc = pkg.Scene(s.cfg["scene"], s.last["div"].cpu().numpy(), device=0)
    ph, gh, _, _ = pkg.zombie.wost_array(sc, s.cfg["solver"], s.cfg["output"], pts.cpu().numpy(), mode=pkg.capi.MODE_DETERMINISTIC, seed=int(s.opts.seed))
    assert np.array_equal(p.cpu().numpy(), ph) and np.array_equal(gp.cpu().numpy(), gh)


def test_distributed_stepper_two_gpus():
    """Data-parallel fits (gradient all_reduce) + sharded pressure solve (all_gather) under torchrun; needs 2 GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(here, "dist_stepper_check.py")], capture_output=True, text=True, timeout=400)
    assert r.returncode == 0 and "DIST_STEPPER_OK" in r.stdout and "DIST_STEPPER_CLEAN_EXIT" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
