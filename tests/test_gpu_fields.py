"""Grid post-processing kernels (include/nmcfs_fields.h) against the reference's own CPU tools: scipy.ndimage.
map_coordinates with the arguments src/2d/move_density.py:92-95 and src/3d/move_density.py:184-185 use, and numpy."""
import math

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ndi = pytest.importorskip("scipy.ndimage")


@pytest.fixture(scope="module")
def F():
    pkg = util.package()
    assert pkg.capi.device_count() > 0
    return pkg.load_fields()


def _reference_step(d, vel, dt, size, mode):
    """move_density.py:98-101,130-136 (2D) / 3D :188-190,207-215, verbatim arithmetic in float64."""
    n = d.shape[0]
    coords = np.indices(d.shape).transpose(tuple(range(1, d.ndim + 1)) + (0,)).astype(float)
    coords = coords/n*(size[1] - size[0]) + size[0]
    back = [coords[..., a] - dt*vel[..., a] for a in range(d.ndim)]
    back_pos = (np.stack(back) - size[0])*n/(size[1] - size[0])
    return ndi.map_coordinates(d, back_pos, order=1, prefilter=False, mode=mode, cval=0)


@pytest.mark.parametrize("dim,mode,n", [(2, "constant", 257), (3, "nearest", 48), (2, "nearest", 64), (3, "constant", 33)])
def test_density_advection_matches_map_coordinates(F, dim, mode, n):
    rng = np.random.default_rng(dim*100 + n)
    size = (-1.0, 1.0)
    d = rng.random((n,)*dim).astype(np.float32)
    # velocities large enough that many nodes trace back across several cells and out of the grid
    vel = (rng.standard_normal((n,)*dim + (dim,))*3.0).astype(np.float32)
    dt = 0.05
    want = _reference_step(d.astype(np.float64), vel.astype(np.float64), dt, size, mode)
    got = F.advect_density(torch.from_numpy(d).cuda(), torch.from_numpy(vel).cuda(), dt, [size[0]]*dim, [size[1] - size[0]]*dim, mode).cpu().numpy()
    # positions are computed in fp32 (like the kernel's inputs): a node whose back-traced position sits within
    # rounding of a cell or grid boundary may interpolate from the neighbouring cell (continuous) or, in
    # 'constant' mode at the outer edge, flip to 0 (discontinuous): allow a handful of those
    err = np.abs(got - want)
    assert np.isfinite(got).all()
    assert (err > 2e-4).mean() < 2e-4, "fraction off: %g, max %g" % ((err > 2e-4).mean(), err.max())
    assert np.median(err) < 1e-5
    if mode == "constant":
        assert ((want == 0) == (got == 0)).mean() > 0.9995


def test_identity_and_edge_cases(F):
    d = torch.rand(17, 23, device="cuda")
    z = torch.zeros(17, 23, 2, device="cuda")
    # zero velocity: a copy up to the rounding of (i/n*extent + lo - lo)*n/extent, as in the reference's own arithmetic
    assert torch.allclose(F.advect_density(d, z, 0.1, [0.0, 0.0], [1.0, 1.0]), d, atol=2e-5)
    one = torch.rand(1, 1, device="cuda")
    assert torch.allclose(F.advect_density(one, torch.zeros(1, 1, 2, device="cuda"), 0.1, [0.0, 0.0], [1.0, 1.0]), one)
    with pytest.raises(ValueError):
        F.advect_density(d, torch.zeros(17, 23, 3, device="cuda"), 0.1, [0.0, 0.0], [1.0, 1.0])
    # uniform shift by exactly one cell along axis 0: out[i] = d[i-1], first row traced outside -> 0 (constant)
    n = 32
    d = torch.rand(n, n, device="cuda")
    v = torch.zeros(n, n, 2, device="cuda"); v[..., 0] = 1.0
    out = F.advect_density(d, v, 2.0/n, [-1.0, -1.0], [2.0, 2.0], "constant")
    assert torch.allclose(out[1:], d[:-1], atol=1e-5) and (out[0].abs() < 1e-5).all()
    out = F.advect_density(d, v, 2.0/n, [-1.0, -1.0], [2.0, 2.0], "nearest")
    assert torch.allclose(out[0], d[0], atol=1e-5)


def test_taylor_green_error_metric(F):
    """The error metric of move_density.py:103-106,143-146 for a 'network' that is the analytic field plus a known
    perturbation, against numpy in double."""
    size = (-1.0, 1.0, -1.0, 1.0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor([0.01, -0.02]))

        def forward(self, x):
            return F.taylor_green_velocity(x, size) + self.w*x

    net = Net().cuda()
    n = 300
    e = F.taylor_green_error(net, size, n)
    g = np.indices((n, n)).transpose(1, 2, 0).astype(float)/n*2 - 1
    want = np.mean(np.sum((np.array([0.01, -0.02])*g)**2, axis=-1))
    assert e == pytest.approx(want, rel=2e-3)
    a = torch.rand(1000, 3, device="cuda"); b = torch.rand(1000, 3, device="cuda")
    assert F.mean_squared_error(a, b) == pytest.approx(((a - b).double()**2).sum(-1).mean().item(), rel=1e-6)
    assert math.isfinite(e)


def test_backtrace_and_fused_mse(F):
    """The two glue kernels of the fit iteration: clamp(x - dt u, lo, hi) and (diff, dL/dy, loss) of the MSE."""
    for dim in (2, 3):
        x = torch.rand(5000, dim, device="cuda")*2 - 1
        u = torch.randn(5000, dim, device="cuda")*3
        lo, hi = [-1.0, -0.5, -0.25][:dim], [1.0, 0.75, 0.5][:dim]
        want = torch.clamp(x - u*0.1, min=torch.tensor(lo, device="cuda"), max=torch.tensor(hi, device="cuda"))
        assert torch.allclose(F.backtrace(x, u, 0.1, lo, hi), want, atol=1e-6)
    pkg = util.package()
    S = pkg.load_siren()
    import ctypes as C
    L = S._lib()
    y = torch.randn(4096, 3, device="cuda"); t = torch.randn(4096, 3, device="cuda")
    diff = torch.empty_like(y); gy = torch.empty_like(y); loss = torch.full((), 123.0, device="cuda")   # no zero-fill needed
    assert L.nmc_mse_grad(y.data_ptr(), t.data_ptr(), y.numel(), diff.data_ptr(), gy.data_ptr(), loss.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert torch.equal(diff, y - t) and torch.allclose(gy, (y - t)*(2.0/y.numel()), rtol=1e-6, atol=0)
    assert loss.item() == pytest.approx(((y - t).double()**2).mean().item(), rel=1e-5)
