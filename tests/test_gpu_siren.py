"""Fused SIREN kernels (csrc/siren.cu, csrc/siren_tc.cu) against a plain PyTorch fp32 evaluation of the same
network (the reference's MLP: nn.Sequential of Linear / sin(30 .), src/2d/models/networks.py:24-68)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

# (in, hidden, hidden layers, out): the four network shapes of the shipped configs
SHAPES = [(2, 64, 6, 2), (2, 128, 2, 2), (3, 64, 5, 3), (3, 128, 2, 3)]


@pytest.fixture(scope="module")
def siren():
    pkg = util.package()
    assert pkg.capi.device_count() > 0
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return pkg.load_siren()


def _net(siren, shape, seed=0, tensor_cores=False):
    torch.manual_seed(seed)
    i, h, l, o = shape
    return siren.FusedSiren(i, o, l, h, nonlinearity="sine", tensor_cores=tensor_cores).cuda()


def _coords(n, dim, seed=1):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return torch.rand((n, dim), generator=g, device="cuda")*2 - 1


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", [4096, 1000, 1])
def test_forward_matches_torch(siren, shape, n):
    net = _net(siren, shape)
    x = _coords(n, shape[0])
    with torch.no_grad():
        y = net(x)
        ref = net.forward_reference(x)
    assert y.shape == ref.shape
    # fp32 with a different summation order; sin(30 .) amplifies rounding layer by layer
    assert (y - ref).abs().max().item() <= 2e-5*max(ref.abs().max().item(), 1e-3) + 2e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_backward_matches_torch_autograd(siren, shape):
    net = _net(siren, shape, seed=3)
    n = 2048 + 37
    x = _coords(n, shape[0], seed=5).requires_grad_(True)
    target = torch.sin(3*x[:, :1]).expand(-1, shape[3]).detach()
    loss = ((net(x) - target)**2).mean()
    grads = torch.autograd.grad(loss, [x] + list(net.parameters()))
    xr = x.detach().clone().requires_grad_(True)
    loss_r = ((net.forward_reference(xr) - target)**2).mean()
    grads_r = torch.autograd.grad(loss_r, [xr] + list(net.parameters()))
    assert abs(loss.item() - loss_r.item()) <= 1e-5*abs(loss_r.item()) + 1e-9
    for g, r in zip(grads, grads_r):
        assert g.shape == r.shape
        assert (g - r).abs().max().item() <= 2e-4*r.abs().max().item() + 1e-9, (tuple(g.shape), (g - r).abs().max().item(), r.abs().max().item())


def test_divergence_through_fused_network(siren):
    """get_divergence (model_split.py:230-243) differentiates the network w.r.t. its input coordinates."""
    net = _net(siren, (2, 64, 6, 2), seed=7)
    x = _coords(3000, 2, seed=9).requires_grad_(True)
    u = net(x)
    div = sum(torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=True)[0][:, i] for i in range(2))
    xr = x.detach().clone().requires_grad_(True)
    ur = net.forward_reference(xr)
    div_r = sum(torch.autograd.grad(ur[:, i], xr, torch.ones_like(ur[:, i]), retain_graph=True)[0][:, i] for i in range(2))
    assert (div - div_r).abs().max().item() <= 2e-4*div_r.abs().max().item()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", [128*300 + 5, 77])
def test_tensor_core_forward_matches_fp32_kernel(siren, shape, n):
    """tcgen05 path with the 3xTF32 split vs the exact-fp32 kernel and vs torch."""
    net = _net(siren, shape, seed=11)
    x = _coords(n, shape[0], seed=13)
    with torch.no_grad():
        ref = net.forward_reference(x)
        y32 = net(x)
        net.tensor_cores = True
        ytc = net(x)
        net.tensor_cores = False
    scale = max(ref.abs().max().item(), 1e-3)
    assert (ytc - y32).abs().max().item() <= 1e-4*scale, (ytc - y32).abs().max().item()/scale
    assert (ytc - ref).abs().max().item() <= 1e-4*scale


def test_fused_adam_matches_torch(siren):
    torch.manual_seed(0)
    p0 = [torch.randn(64, 64, device="cuda"), torch.randn(64, device="cuda"), torch.randn(2, 64, device="cuda")]
    a = [torch.nn.Parameter(p.clone()) for p in p0]
    b = [torch.nn.Parameter(p.clone()) for p in p0]
    oa = siren.FusedAdam(a, lr=1e-3)
    ob = torch.optim.Adam(b, lr=1e-3)
    for it in range(12):
        for pa, pb in zip(a, b):
            g = torch.randn_like(pb)*(1 + it)
            pa.grad = g.clone(); pb.grad = g.clone()
        oa.step(); ob.step()
    for pa, pb in zip(a, b):
        assert (pa - pb).abs().max().item() <= 2e-6*pb.abs().max().item() + 1e-7


def test_fit_loop_tracks_reference(siren):
    """A short advection-style fit (model_split.py:88-120 shape: MSE against a frozen target network, Adam lr 1e-5
    scaled up so that 60 iterations move the loss) with fused kernels vs stock ops from the same initial weights."""
    shape = (2, 64, 6, 2)
    net_f, net_r = _net(siren, shape, seed=21), _net(siren, shape, seed=21)
    tgt = _net(siren, shape, seed=22)
    opt_f = siren.FusedAdam(list(net_f.parameters()), lr=1e-4)
    opt_r = torch.optim.Adam(net_r.parameters(), lr=1e-4)
    lf = lr_ = None
    for it in range(60):
        x = _coords(4096, 2, seed=100 + it)
        with torch.no_grad():
            t = tgt.forward_reference(x)
        lf = ((net_f(x) - t)**2).mean(); opt_f.zero_grad(); lf.backward(); opt_f.step()
        lr_ = ((net_r.forward_reference(x) - t)**2).mean(); opt_r.zero_grad(); lr_.backward(); opt_r.step()
    assert abs(lf.item() - lr_.item()) <= 2e-3*abs(lr_.item())


def _torch_envelope(x, size, eps):
    w = []
    for i in range(x.shape[1]):
        lo, hi = size[2*i], size[2*i + 1]
        w.append(torch.min((x[:, i] - lo).abs().clamp(min=0, max=eps), (x[:, i] - hi).abs().clamp(min=0, max=eps))/eps)
    return torch.stack(w, dim=-1).detach()


@pytest.mark.parametrize("shape", [(2, 64, 6, 2), (3, 128, 2, 3)])
def test_fused_envelope_matches_reference_query_velocity(siren, shape):
    """query_velocity's wall weights (base.py:179-187) fused into the kernels: forward (fp32 and tensor-core)
    and backward (parameter gradients and the detached-weight input gradient used by the divergence)."""
    net = _net(siren, shape, seed=31)
    dim = shape[0]
    size = (-1.0, 1.0)*dim
    eps = 0.2  # wide enough that many samples sit inside the ramp
    env = siren.wall_envelope(size, eps)
    x = _coords(3000, dim, seed=33).requires_grad_(True)
    y = net(x, envelope=env)
    loss = (y**2).sum()
    grads = torch.autograd.grad(loss, [x] + list(net.parameters()))
    xr = x.detach().clone().requires_grad_(True)
    yr = net.forward_reference(xr)*_torch_envelope(xr, size, eps)
    grads_r = torch.autograd.grad((yr**2).sum(), [xr] + list(net.parameters()))
    scale = yr.abs().max().item()
    assert (y - yr).abs().max().item() <= 3e-5*scale
    for g, r in zip(grads, grads_r):
        assert (g - r).abs().max().item() <= 3e-4*r.abs().max().item() + 1e-9
    with torch.no_grad():
        net.tensor_cores = True
        ytc = net(x.detach(), envelope=env)
        net.tensor_cores = False
    assert (ytc - yr).abs().max().item() <= 1e-4*scale


@pytest.mark.parametrize("scenario", ["karman", "smoke_obs", "karman3d", "smoke"])
def test_general_envelope_matches_reference_query_velocity(siren, scenario):
    """The karman (base.py:169-181), smoke_obs (3d base.py:224-244), karman3d (3d base.py:257-275, cylinder of 3d
    main.py:92-98) and smoke (3d base.py:197-222, per-step inlet noise) branches of query_velocity inside the kernels:
    region override, obstacle weight (NOT detached: its gradient reaches x), wall weights; forward on both kernels
    and backward against stock autograd on the torch transcription of the reference (siren.envelope_reference)."""
    if scenario == "karman":
        shape, size, eps = (2, 128, 2, 2), (-1.0, 1.0, -0.4, 0.4), 0.15
        env = siren.karman_envelope(size, eps, centre=(-0.3, 0.05), radius=0.12, karman_vel=0.5)
    elif scenario == "smoke_obs":
        shape, size, eps = (3, 64, 5, 3), (-1.0, 1.0)*3, 0.2
        env = siren.smoke_obs_envelope(size, eps, centre=(0.1, 0.0, 0.2), radius=0.25, inlet_centre=(0.0, 0.0, -0.6), inlet_radius=0.3)
    elif scenario == "karman3d":
        shape, size, eps = (3, 128, 2, 3), (-1.0, 1.0)*3, 0.2
        env = siren.karman3d_envelope(size, eps, centre_xz=(0.1, -0.3), radius=0.25, karman_vel=0.5)
    else:
        shape, size, eps = (3, 64, 5, 3), (-1.0, 1.0)*3, 0.2
        seed = torch.tensor([17], dtype=torch.int32, device="cuda")
        env = siren.smoke_envelope(size, eps, seed, inlet_radius=0.45)
    dim = shape[0]
    net = _net(siren, shape, seed=41)
    lo = torch.tensor(size[0::2], device="cuda"); hi = torch.tensor(size[1::2], device="cuda")
    g = torch.Generator(device="cuda").manual_seed(43)
    x = (torch.rand(6000, dim, device="cuda", generator=g)*(hi - lo)*1.04 + lo - 0.02*(hi - lo)).requires_grad_(True)
    y = net(x, envelope=env)
    w = torch.randn(y.shape, device="cuda", generator=g)
    grads = torch.autograd.grad((y*w).sum() + (y**2).sum(), [x] + list(net.parameters()))
    xr = x.detach().clone().requires_grad_(True)
    yr = siren.envelope_reference(env, xr, net.forward_reference(xr))
    grads_r = torch.autograd.grad((yr*w).sum() + (yr**2).sum(), [xr] + list(net.parameters()))
    scale = yr.abs().max().item()
    # every part of the envelope is exercised by the sample set
    if scenario == "smoke":
        inlet = torch.linalg.norm(x.detach() - torch.tensor([0.0, 0.0, -0.6], device="cuda"), dim=-1) < 0.45
        assert inlet.sum() > 100
        interior = inlet & (x.detach().abs() < 1 - eps).all(dim=-1)   # wall weights are 1 there: u = 0.1 * noise
        un = yr.detach()[interior, 0]/0.1
        assert interior.sum() > 100 and un.min() < -0.9 and un.max() > 0.9 and abs(un.mean().item()) < 0.2   # uniform in [-1, 1)
        assert torch.allclose(yr.detach()[interior, 2], 0.2 + 0.1*un, atol=1e-6)
        seed.fill_(18)   # the next time step draws new noise (the kernels read the seed from device memory)
        y2 = net(x.detach(), envelope=env)
        assert (y2[inlet] - y.detach()[inlet]).abs().max().item() > 0.05 and torch.equal(y2[~inlet], y.detach()[~inlet])
        seed.fill_(17)
    else:
        sel = torch.tensor([float(((env.sphere_axes or 7) >> i) & 1) for i in range(dim)], device="cuda")
        dist = torch.linalg.norm((x.detach() - torch.tensor([env.sphere_c[i] for i in range(dim)], device="cuda"))*sel, dim=-1) - env.sphere_r
        assert ((dist > 0) & (dist < eps)).sum() > 100 and (dist < 0).sum() > 20
    assert (y - yr).abs().max().item() <= 3e-5*scale
    gx, gxr = grads[0], grads_r[0]
    assert (gx - gxr).abs().max().item() <= 5e-4*gxr.abs().max().item() + 1e-9
    for a, b in zip(grads[1:], grads_r[1:]):
        assert (a - b).abs().max().item() <= 5e-4*b.abs().max().item() + 1e-9
    # the obstacle weight contributes to dL/dx: without it the input gradient would be off by far more than the tolerance
    if scenario != "smoke":
        yd = siren.envelope_reference(env, xr, net.forward_reference(xr).detach())
        gobs = torch.autograd.grad((yd*w).sum() + (yd**2).sum(), xr)[0]
        assert gobs.abs().max().item() > 50*5e-4*gxr.abs().max().item()
    with torch.no_grad():
        net.tensor_cores = True
        ytc = net(x.detach(), envelope=env)
        net.tensor_cores = False
    assert (ytc - yr).abs().max().item() <= 1e-4*scale
    with pytest.raises(RuntimeError, match="eps"):
        bad = siren.wall_envelope(size, 0.0)
        net(x.detach(), envelope=bad)


def test_direct_fit_iteration_equals_autograd_iteration(siren):
    """DirectFit.iterate (no autograd, 5 launches) vs loss.backward() + FusedAdam on the same data."""
    shape = (2, 64, 6, 2)
    a, b = _net(siren, shape, seed=41), _net(siren, shape, seed=41)
    env = siren.wall_envelope((-1.0, 1.0, -1.0, 1.0), 0.1)
    fit = siren.DirectFit(a, lr=1e-3, envelope=env, max_batch=4096)
    opt = siren.FusedAdam(list(b.parameters()), lr=1e-3)
    for it in range(5):
        x = _coords(4096, 2, seed=50 + it)
        t = torch.sin(2*x)
        fit.iterate(x, t)
        loss = torch.mean((b(x, envelope=env) - t)**2)
        opt.zero_grad(); loss.backward(); opt.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert (pa - pb).abs().max().item() <= 1e-5*pb.abs().max().item() + 1e-8


def test_graph_replayed_fit_matches_torch_adam(siren):
    """The fit iteration captured once in a CUDA graph and replayed (stepper.py) must be torch.optim.Adam step for
    step: Adam's bias corrections depend on the step number, which therefore lives in device memory
    (nmc_adam_step_device) instead of being frozen into the graph at capture time."""
    shape = (2, 64, 6, 2)
    net = _net(siren, shape, seed=51)
    ref = _net(siren, shape, seed=51)
    x = _coords(4096, 2, seed=52)
    target = torch.stack([torch.sin(3*x[:, 0]), torch.cos(2*x[:, 1])], dim=-1)
    lr, iters = 1e-4, 60
    fit = siren.DirectFit(net, lr, None, max_batch=4096)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fit.iterate(x, target)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fit.iterate(x, target)
    for _ in range(iters - 3):
        graph.replay()
    torch.cuda.synchronize()
    assert int(fit.opt.step_dev.item()) == iters
    opt = torch.optim.Adam(ref.parameters(), lr=lr)
    for _ in range(iters):
        loss = torch.mean((ref.forward_reference(x) - target)**2)
        opt.zero_grad(); loss.backward(); opt.step()
    moved = 0.0
    fresh = _net(siren, shape, seed=51)
    for a, b, c in zip(net.parameters(), ref.parameters(), fresh.parameters()):
        step = (b - c).abs().max().item()
        moved = max(moved, step)
        assert (a - b).abs().max().item() <= 0.02*step + 1e-7   # frozen corrections would be off by a factor of ~5
    assert moved > 10*lr
    fit.opt.reset()
    assert int(fit.opt.step_dev.item()) == 0 and not fit.opt.m.any()


@pytest.mark.parametrize("shape", [(2, 128, 2, 2), (3, 64, 5, 3)])
def test_tensor_core_training_forward_equals_fp32_path(siren, shape):
    """DirectFit at batch 16384 takes the tcgen05 forward, which also writes the pre-activations for the backward
    chain (3xTF32 split: fp32-level accuracy): same parameter trajectory as the fp32 split kernels."""
    a = _net(siren, shape, seed=61, tensor_cores=True)
    b = _net(siren, shape, seed=61, tensor_cores=False)
    start = [p.detach().clone() for p in a.parameters()]
    x = _coords(16384, shape[0], seed=62)
    target = torch.sin(x[:, :1]*3.0).repeat(1, shape[3])
    env = siren.wall_envelope((-1.0, 1.0)*shape[0], 0.1)
    fa = siren.DirectFit(a, 1e-4, env, max_batch=16384)
    fb = siren.DirectFit(b, 1e-4, env, max_batch=16384)
    assert fa.tensor_cores and not fb.tensor_cores
    for _ in range(6):
        da = fa.iterate(x, target); db = fb.iterate(x, target)
    assert (da - db).abs().max().item() <= 2e-5*db.abs().max().item() + 1e-7
    za, zb = fa.z[: fb.z.numel()], fb.z
    assert (za - zb).abs().max().item() <= 2e-5*zb.abs().max().item()      # saved pre-activations of the last iteration
    _assert_same_trajectory(a, b, start)


def _assert_same_trajectory(a, b, start):
    """Two implementations of the same Adam fit: the parameters agree to 2 % of the distance moved, except for the few
    entries whose gradient is ~0 next to the tensor's other entries -- Adam normalises every entry by its own running
    magnitude, so an absolute gradient difference of 1e-5 x max|g| (summation order, 3xTF32 vs fp32 FMA) can move such
    an entry by a sizeable fraction of lr per step in either implementation."""
    for p, q, r in zip(a.parameters(), b.parameters(), start):
        moved = (q - r).abs().max().item()
        d = (p - q).abs()
        assert (d <= 0.02*moved + 1e-8).float().mean().item() >= 0.999
        assert d.max().item() <= 0.25*moved + 1e-8


def _raw_backward(siren, net, x, gy, env, tc):
    """dZ of the delta chain and the flat weight / bias gradients, through the C ABI, fp32 kernels or tcgen05 kernels."""
    import ctypes as C
    L = siren._lib()
    lin = net._linears()
    W = [m.weight.detach().contiguous() for m in lin]; b = [m.bias.detach().contiguous() for m in lin]
    sh = siren._shape_of(W, 30.0)
    n = x.shape[0]
    H, Lh = sh.hidden, sh.n_hidden_layers
    y = torch.empty((n, sh.out_dim), device=x.device)
    z = torch.empty(((Lh + 1)*H, n), device=x.device)
    siren._check(L.nmc_siren_forward(C.byref(sh), siren._ptrs(W), siren._ptrs(b), x.data_ptr(), n, y.data_ptr(), z.data_ptr(), env, siren._stream()))
    dZ = torch.zeros(((Lh + 1)*H + sh.out_dim, n), device=x.device)
    gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(v) for v in b]
    if tc:
        siren._check(L.nmc_siren_backward_tc(C.byref(sh), siren._ptrs(W), siren._ptrs(b), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(),
                                             dZ.data_ptr(), env, siren._stream()))
        siren._check(L.nmc_siren_weight_grads_tc(C.byref(sh), x.data_ptr(), n, dZ.data_ptr(), z.data_ptr(), siren._ptrs(gW), siren._ptrs(gb), siren._stream()))
    else:
        A = torch.empty(((Lh + 1)*H, n), device=x.device)
        siren._check(L.nmc_siren_backward(C.byref(sh), siren._ptrs(W), siren._ptrs(b), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(),
                                          dZ.data_ptr(), A.data_ptr(), None, env, siren._stream()))
        siren._check(L.nmc_siren_weight_grads(C.byref(sh), x.data_ptr(), n, dZ.data_ptr(), A.data_ptr(), siren._ptrs(gW), siren._ptrs(gb), siren._stream()))
    torch.cuda.synchronize()
    return dZ, gW, gb


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", [4096, 16384, 128*9 + 52, 60])
@pytest.mark.parametrize("with_env", [False, True])
def test_tensor_core_backward_matches_fp32_kernels_and_autograd(siren, shape, n, with_env):
    """tcgen05 delta chain + weight gradients (csrc/siren_tc_bwd.cu, 3xTF32) vs the exact-fp32 kernels (every delta,
    every gradient entry) and vs torch autograd through the reference's stock-op network (base.py:83-96)."""
    if with_env and n not in (4096, 60):
        pytest.skip("envelope covered at two batch sizes")
    net = _net(siren, shape, seed=71)
    x = _coords(n, shape[0], seed=72)
    env = siren.wall_envelope((-1.0, 1.0)*shape[0], 0.1) if with_env else None
    target = torch.sin(3*x[:, :1]).expand(-1, shape[3]).contiguous()
    with torch.no_grad():
        y = net(x, envelope=env)
    gy = ((y - target)*(2.0/y.numel())).contiguous()
    dZ32, gW32, gb32 = _raw_backward(siren, net, x, gy, env, tc=False)
    dZtc, gWtc, gbtc = _raw_backward(siren, net, x, gy, env, tc=True)
    H, Lh = shape[1], shape[2]
    for l in range(Lh + 1):
        a, b = dZtc[l*H:(l + 1)*H], dZ32[l*H:(l + 1)*H]
        assert (a - b).abs().max().item() <= 2e-5*b.abs().max().item() + 1e-12, ("dZ", l, (a - b).abs().max().item(), b.abs().max().item())
    assert torch.equal(dZtc[(Lh + 1)*H:], dZ32[(Lh + 1)*H:])
    for l, (a, b) in enumerate(zip(gWtc, gW32)):
        assert (a - b).abs().max().item() <= 3e-5*b.abs().max().item() + 1e-12, ("gW", l, (a - b).abs().max().item(), b.abs().max().item())
    for l, (a, b) in enumerate(zip(gbtc, gb32)):
        assert (a - b).abs().max().item() <= 3e-5*b.abs().max().item() + 1e-12, ("gb", l, (a - b).abs().max().item(), b.abs().max().item())
    # torch autograd through stock ops
    ref_out = siren.envelope_reference(env, x, net.forward_reference(x))
    loss = ((ref_out - target)**2).mean()
    lin = net._linears()
    gr = torch.autograd.grad(loss, [m.weight for m in lin] + [m.bias for m in lin])
    for a, r in zip(gWtc + gbtc, gr):
        assert (a - r).abs().max().item() <= 2e-4*r.abs().max().item() + 1e-9, (tuple(a.shape), (a - r).abs().max().item(), r.abs().max().item())


@pytest.mark.parametrize("shape", [(2, 64, 6, 2), (3, 128, 2, 3)])
def test_tensor_core_backward_in_direct_fit(siren, shape):
    """DirectFit with the tcgen05 backward (default for tensor_cores networks) follows the fp32-kernel fit step for step."""
    a = _net(siren, shape, seed=81, tensor_cores=True)
    b = _net(siren, shape, seed=81, tensor_cores=False)
    start = [p.detach().clone() for p in a.parameters()]
    n = 4096
    x = _coords(n, shape[0], seed=82)
    target = torch.cos(x[:, :1]*2.0).repeat(1, shape[3])
    fa = siren.DirectFit(a, 1e-4, None, max_batch=n)
    fb = siren.DirectFit(b, 1e-4, None, max_batch=n)
    assert fa.tc_backward and not fb.tc_backward
    for _ in range(8):
        fa.iterate(x, target); fb.iterate(x, target)
    _assert_same_trajectory(a, b, start)


@pytest.mark.parametrize("shape", [(2, 64, 6, 2), (3, 64, 5, 3), (2, 64, 1, 2)])
@pytest.mark.parametrize("n", [4096, 16384, 128*150 + 52, 60])
@pytest.mark.parametrize("with_env", [False, True])
def test_fused_tensor_core_backward_matches_fp32_kernels_and_autograd(siren, shape, n, with_env):
    """The one-kernel backward of the hidden = 64 networks (csrc/siren_tc_fused_bwd.cu: delta chain and every gradient, the
    delta / activation buffers read K-major for the chain and MN-major for the gradients, bias gradients from a column of ones,
    gradient tiles accumulated in TMEM across a CTA's tiles) vs the exact-fp32 kernels and torch autograd."""
    import ctypes as C
    if with_env and n not in (4096, 60):
        pytest.skip("envelope covered at two batch sizes")
    net = _net(siren, shape, seed=91)
    x = _coords(n, shape[0], seed=92)
    env = siren.wall_envelope((-1.0, 1.0)*shape[0], 0.1) if with_env else None
    target = torch.sin(3*x[:, :1]).expand(-1, shape[3]).contiguous()
    with torch.no_grad():
        y = net(x, envelope=env)
    gy = ((y - target)*(2.0/y.numel())).contiguous()
    _, gW32, gb32 = _raw_backward(siren, net, x, gy, env, tc=False)
    L = siren._lib()
    lin = net._linears()
    W = [m.weight.detach().contiguous() for m in lin]; b = [m.bias.detach().contiguous() for m in lin]
    sh = siren._shape_of(W, 30.0)
    z = torch.empty(((sh.n_hidden_layers + 1)*sh.hidden, n), device=x.device)
    yy = torch.empty((n, sh.out_dim), device=x.device)
    siren._check(L.nmc_siren_forward(C.byref(sh), siren._ptrs(W), siren._ptrs(b), x.data_ptr(), n, yy.data_ptr(), z.data_ptr(), env, siren._stream()))
    gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(v) for v in b]
    siren._check(L.nmc_siren_backward_fused_tc(C.byref(sh), siren._ptrs(W), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(),
                                               siren._ptrs(gW), siren._ptrs(gb), env, siren._stream()))
    torch.cuda.synchronize()
    for l, (a, r) in enumerate(zip(gW, gW32)):
        assert (a - r).abs().max().item() <= 3e-5*r.abs().max().item() + 1e-12, ("gW", l, (a - r).abs().max().item(), r.abs().max().item())
    for l, (a, r) in enumerate(zip(gb, gb32)):
        assert (a - r).abs().max().item() <= 3e-5*r.abs().max().item() + 1e-12, ("gb", l, (a - r).abs().max().item(), r.abs().max().item())
    ref_out = siren.envelope_reference(env, x, net.forward_reference(x))
    loss = ((ref_out - target)**2).mean()
    gr = torch.autograd.grad(loss, [m.weight for m in lin] + [m.bias for m in lin])
    for a, r in zip(gW + gb, gr):
        assert (a - r).abs().max().item() <= 2e-4*r.abs().max().item() + 1e-9


def _counters(step, epoch):
    return (torch.full((), step, dtype=torch.int64, device="cuda"), torch.full((), epoch, dtype=torch.int64, device="cuda"))


@pytest.mark.parametrize("dim", [2, 3])
def test_fit_sample_uniform_kernel(siren, dim):
    """One-launch batch draw of a captured fit iteration (csrc/fit_glue.cu; sample_in_training 'random', base.py:225-241):
    inside the box, uniform, a new draw for every (epoch, step) read from device memory, reproducible for the same key,
    and -- with an obstacle -- redrawn once when inside the ball."""
    lo, hi = [-1.0, 0.5, 2.0][:dim], [3.0, 1.5, 2.5][:dim]
    n = 200000
    st, ep = _counters(3, 7)
    a = siren.fit_sample_uniform(n, lo, hi, st, ep, seed=11)
    assert a.shape == (n, dim)
    for k in range(dim):
        assert a[:, k].min().item() >= lo[k] and a[:, k].max().item() < hi[k]
        u = (a[:, k] - lo[k])/(hi[k] - lo[k])
        assert abs(u.mean().item() - 0.5) < 4/np.sqrt(12*n) and abs(u.var().item() - 1/12) < 1e-3
        hist = torch.histc(u, bins=64, min=0, max=1)
        assert (hist - n/64).abs().max().item() < 5*np.sqrt(n/64)
    if dim > 1:  # the coordinates are independent
        c = torch.corrcoef(a.t())
        assert (c - torch.eye(dim, device="cuda")).abs().max().item() < 0.01
    assert torch.equal(a, siren.fit_sample_uniform(n, lo, hi, st, ep, seed=11))
    st2, ep2 = _counters(4, 7)
    b = siren.fit_sample_uniform(n, lo, hi, st2, ep2, seed=11)
    assert (a == b).float().mean().item() < 1e-3
    st3, ep3 = _counters(3, 8)
    assert (a == siren.fit_sample_uniform(n, lo, hi, st3, ep3, seed=11)).float().mean().item() < 1e-3
    assert (a == siren.fit_sample_uniform(n, lo, hi, st, ep, seed=12)).float().mean().item() < 1e-3
    # graph replay: the counters are read at run time
    out = torch.empty((n, dim), device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        siren.fit_sample_uniform(n, lo, hi, st, ep, seed=11, out=out)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        siren.fit_sample_uniform(n, lo, hi, st, ep, seed=11, out=out)
    st.fill_(4)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, b)
    # obstacle: a ball covering ~20 % of the box (2D) -- after one redraw ~4 % of the points are still inside
    c, r = [0.5*(l + h) for l, h in zip(lo, hi)], 0.25
    st.fill_(3)
    o = siren.fit_sample_uniform(n, lo, hi, st, ep, seed=11, obstacle=(c, r))
    inside = ((o - torch.tensor(c, device="cuda")).norm(dim=1) <= r).float().mean().item()
    frac = ((a - torch.tensor(c, device="cuda")).norm(dim=1) <= r).float().mean().item()
    assert frac > 0 and abs(inside - frac*frac) < 4*np.sqrt(frac*frac/n) + 1e-4
    keep = (a - torch.tensor(c, device="cuda")).norm(dim=1) > r
    assert torch.equal(o[keep], a[keep])   # points outside the obstacle are the first draw


def test_fit_gather_kernel(siren):
    """The projection fit's batch in one launch (model_split.py:272-277): rows floor(u count) of the samples and of grad p."""
    cap, n, dim = 5000, 100000, 3
    src_x = torch.arange(cap*dim, device="cuda", dtype=torch.float32).reshape(cap, dim).contiguous()
    src_g = -src_x
    count = torch.tensor(3000.0, device="cuda")
    st, ep = _counters(9, 2)
    x, g = siren.fit_gather(n, src_x, src_g, count, st, ep, seed=5)
    idx = (x[:, 0]/dim).long()
    assert torch.equal(x, src_x[idx]) and torch.equal(g, src_g[idx])
    assert idx.min().item() >= 0 and idx.max().item() <= 2999
    hist = torch.bincount(idx, minlength=3000).float()
    assert (hist > 0).float().mean().item() > 0.99 and abs(idx.float().mean().item() - 1499.5) < 4*3000/np.sqrt(12*n)
    st.fill_(10)
    x2, _ = siren.fit_gather(n, src_x, src_g, count, st, ep, seed=5)
    assert (x2[:, 0] == x[:, 0]).float().mean().item() < 0.01
    count.fill_(1.0e9)   # the clamp to the buffer
    x3, _ = siren.fit_gather(n, src_x, src_g, count, st, ep, seed=5)
    assert (x3[:, 0]/dim).long().max().item() == cap - 1


@pytest.mark.parametrize("count", [3*4096, 2*1000 + 1])
def test_mse_grad_fit_kernel(siren, count):
    """nmc_mse_grad_fit: the loss kernel that also subtracts grad p from the target, clears the flat gradient buffer and advances
    Adam's device-side step; nmc_adam_update_device then equals torch.optim.Adam for the advanced step."""
    import ctypes as C
    L = siren._lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    y, t, s = (torch.randn(count, generator=g, device="cuda") for _ in range(3))
    diff, gy = torch.empty_like(y), torch.empty_like(y)
    loss = torch.zeros((), device="cuda")
    zero = torch.ones(21123, device="cuda")
    step = torch.full((), 4, dtype=torch.int64, device="cuda")
    flag = torch.zeros((), dtype=torch.int32, device="cuda")
    siren._check(L.nmc_mse_grad_fit(y.data_ptr(), t.data_ptr(), s.data_ptr(), count, diff.data_ptr(), gy.data_ptr(), loss.data_ptr(),
                                    zero.data_ptr(), zero.numel(), step.data_ptr(), 1.1e-10, flag.data_ptr(), siren._stream()))
    assert flag.item() == 0   # the loss is far above the early-stop threshold
    ref = y - (t - s)
    assert torch.equal(diff, ref) and torch.allclose(gy, ref*(2.0/count), rtol=1e-6, atol=0)
    assert loss.item() == pytest.approx((ref.double()**2).mean().item(), rel=1e-5)
    assert zero.abs().max().item() == 0.0 and step.item() == 5
    siren._check(L.nmc_mse_grad_fit(y.data_ptr(), t.data_ptr(), None, count, diff.data_ptr(), gy.data_ptr(), loss.data_ptr(),
                                    None, 0, None, 0.0, None, siren._stream()))
    assert torch.equal(diff, y - t) and step.item() == 5
    # early stop decided on the device: the flag is set at the threshold and stays set
    siren._check(L.nmc_mse_grad_fit(y.data_ptr(), y.data_ptr(), None, count, diff.data_ptr(), gy.data_ptr(), loss.data_ptr(),
                                    None, 0, None, 1.1e-10, flag.data_ptr(), siren._stream()))
    assert loss.item() == 0.0 and flag.item() == 1
    siren._check(L.nmc_mse_grad_fit(y.data_ptr(), t.data_ptr(), None, count, diff.data_ptr(), gy.data_ptr(), loss.data_ptr(),
                                    None, 0, None, 1.1e-10, flag.data_ptr(), siren._stream()))
    assert loss.item() > 1.0e-3 and flag.item() == 1
    # Adam with the device-side counter, not advanced by the update itself
    p = torch.randn(1000, generator=g, device="cuda"); grad = torch.randn(1000, generator=g, device="cuda")
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step.fill_(0)
    for it in range(3):
        pr.grad = grad.clone(); opt.step()
        step += 1
        siren._check(L.nmc_adam_update_device(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8,
                                              step.data_ptr(), None, siren._stream()))
        assert step.item() == it + 1
    assert torch.allclose(p, pr.detach(), rtol=1e-5, atol=1e-7)
    before = p.clone()   # a set flag freezes the parameters, a clear one does not
    siren._check(L.nmc_adam_update_device(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8,
                                          step.data_ptr(), flag.data_ptr(), siren._stream()))
    assert torch.equal(p, before)
    flag.zero_()
    siren._check(L.nmc_adam_update_device(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8,
                                          step.data_ptr(), flag.data_ptr(), siren._stream()))
    assert not torch.equal(p, before)


def test_direct_fit_stops_moving_at_the_early_stop_threshold(siren):
    """base.py:148 leaves the fit at the first iteration with loss <= 1.1e-10.  DirectFit takes that decision on the device (no host
    round trip per iteration): a fit whose target is the network's own output -- the projection of examples/taylorgreen as
    shipped, evaluated by ANOTHER kernel (tcgen05 at the chunk size, fp32 at the batch size), i.e. a loss of ~1e-13 and a
    tiny non-zero gradient that Adam would normalise to full-size steps -- leaves the parameters exactly where they were."""
    net = _net(siren, (2, 64, 6, 2), seed=5, tensor_cores=True)
    x = _coords(4096, 2, seed=6)
    big = _coords(32768, 2, seed=7); big[:4096] = x
    with torch.no_grad():
        target = net(big)[:4096].contiguous()      # 32768 samples: the tcgen05 forward; the training forward at 4096 is the fp32 kernel
    fit = siren.DirectFit(net, 1e-5, None, max_batch=4096)
    start = [p.detach().clone() for p in net.parameters()]
    fit.stop_threshold = 1.1e-10
    for _ in range(5):
        fit.iterate(x, target)
    assert 0.0 <= fit.loss.item() <= 1.1e-10 and fit.opt.stop_flag.item() == 1
    assert all(torch.equal(a, b) for a, b in zip(start, net.parameters()))
    fit.opt.reset(); fit.stop_threshold = 0.0       # without the gate Adam walks away at lr per step
    for _ in range(5):
        fit.iterate(x, target)
    if fit.loss.item() > 0.0:
        assert max((a - b).abs().max().item() for a, b in zip(start, net.parameters())) > 1e-5


@pytest.mark.parametrize("with_sub,count", [(True, 3*4096), (False, 2*4096 + 2)])
def test_adam_update_and_next_fetch_in_one_launch(siren, with_sub, count):
    """csrc/siren.cu adamFetchKernel (between the iterations of an unrolled fit graph): the same bits as nmc_adam_update_device
    followed by nmc_fit_fetch -- parameters, both moments and the three fetched buffers -- for a 16-byte-aligned and an odd slot
    size, with and without the projection fit's third ring, and with the early-stop flag set (no update, the fetch still runs)."""
    import torch
    S = siren
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    slots = 8
    ring_x, ring_t = torch.randn(slots, count, device=dev), torch.randn(slots, count, device=dev)
    ring_s = torch.randn(slots, count, device=dev) if with_sub else None
    for stopped in (False, True):
        outs = []
        for merged in (False, True):
            torch.manual_seed(5)
            params = [torch.nn.Parameter(torch.randn(64, 3, device=dev)), torch.nn.Parameter(torch.randn(64, device=dev)),
                      torch.nn.Parameter(torch.randn(4, 64, 64, device=dev))]
            opt = S.FusedAdam(params, lr=1e-3)
            opt.g.copy_(torch.randn_like(opt.g)); opt.m.copy_(torch.rand_like(opt.m)*0.1); opt.v.copy_(torch.rand_like(opt.v)*0.01)
            opt.step_dev.fill_(13)                 # slot 13 % 8 = 5
            opt.stop_flag.fill_(1 if stopped else 0)
            ox, ot = torch.zeros(count, device=dev), torch.zeros(count, device=dev)
            os_ = torch.zeros(count, device=dev) if with_sub else None
            before = opt.flat.clone()
            if merged:
                opt.update_flat_fetch(True, ring_x, ring_t, ring_s, ox, ot, os_)
            else:
                opt.update_flat(True)
                S.fit_fetch(ring_x, ring_t, ring_s, opt.step_dev, ox, ot, os_)
            torch.cuda.synchronize()
            assert torch.equal(ox, ring_x[5]) and torch.equal(ot, ring_t[5]) and (not with_sub or torch.equal(os_, ring_s[5]))
            assert torch.equal(opt.flat, before) == stopped
            outs.append((opt.flat.clone(), opt.m.clone(), opt.v.clone()))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
