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
    args = dict(scene_size=(0.0, 2*math.pi, 0.0, 2*math.pi), grid_resolution=200, wost_resolution=64, sample_resolution=32,
                max_n_iters=40, check_every=10, seed=3, device=0)
    args.update(kw)
    return pkg, st.SplitStepper(cfg, **args)


def _tg(x):  # Taylor-Green vortex
    return torch.stack([torch.sin(x[:, 0])*torch.cos(x[:, 1]), -torch.cos(x[:, 0])*torch.sin(x[:, 1])], dim=-1)


def test_divergence_grid_matches_stock_autograd():
    pkg, s = _stepper(use_cuda_graph=False)
    div = s.divergence_grid()
    assert tuple(div.shape) == (202, 202)
    x = s.grid_samples.detach().clone().requires_grad_(True)
    u = s.velocity_field_prev.forward_reference(x)*s.envelope(x)
    ref = 0.0
    for i in range(2):
        ref = ref + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=True)[0][:, i]
    ref = (-ref).reshape(202, 202)
    assert (div - ref).abs().max().item() <= 3e-4*ref.abs().max().item() + 1e-7
    # orientation: rows <-> y, columns <-> x (image.h:70-75): grid_samples[i*W + j] = (x_j, y_i)
    g = s.grid_samples.reshape(202, 202, 2)
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
