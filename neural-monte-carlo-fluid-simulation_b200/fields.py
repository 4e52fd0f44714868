"""Grid post-processing on the device (include/nmcfs_fields.h): the second stage of the reference's demos.

    advect_density      src/2d/move_density.py:92-95,130-136 (map_coordinates order=1, mode='constant', cval=0)
                        src/3d/move_density.py:184-185,211-215 (mode='nearest')
    taylor_green_*      src/2d/sources.py:19-32 (initial / steady-state velocity), move_density.py:98-106,143-146 (error)
The reference moves the 1000^2 (200^3) velocity grid to the host every time step and interpolates with scipy;
here the grid, the network evaluation and the reduction stay in HBM.  No CPU fallback: CUDA tensors only.
"""
import ctypes as C
import math

import torch

from . import capi

FIELDS_EXPORTS = ["nmc_fields_last_error", "nmc_advect_density", "nmc_backtrace", "nmc_sum_squared_error"]
_ready = False


def _lib():
    global _ready
    L = capi.lib()
    if not _ready:
        L.nmc_fields_last_error.restype = C.c_char_p
        L.nmc_advect_density.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_float, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p]
        L.nmc_sum_squared_error.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.nmc_backtrace.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        _ready = True
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError("libnmcfs fields: " + _lib().nmc_fields_last_error().decode())


def _cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s must be a CUDA tensor (there is no CPU path)" % name)
    return t.contiguous().float()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def advect_density(density, velocity, dt, lo, extent, mode="constant"):
    """density [n0,n1(,n2)], velocity [n0,n1(,n2),dim] on the node grid x = lo + index/n*extent -> advected density."""
    d = _cuda_f32(density, "density"); v = _cuda_f32(velocity, "velocity")
    dim = d.dim()
    if dim not in (2, 3) or tuple(v.shape) != tuple(d.shape) + (dim,):
        raise ValueError("density must be 2D/3D and velocity must have shape density.shape + (dim,)")
    out = torch.empty_like(d)
    shape = (C.c_int*3)(*d.shape, *([1]*(3 - dim)))
    flo = (C.c_float*3)(*[float(x) for x in lo], *([0.0]*(3 - dim)))
    fex = (C.c_float*3)(*[float(x) for x in extent], *([1.0]*(3 - dim)))
    m = {"constant": 0, "nearest": 1}[mode]
    with torch.cuda.device(d.device):
        _check(_lib().nmc_advect_density(dim, shape, d.data_ptr(), v.data_ptr(), float(dt), flo, fex, m, out.data_ptr(), _stream()))
    return out


def backtrace(x, u, dt, lo, hi):
    """clamp(x - dt*u, lo, hi) in one launch (the back-traced positions of _advect_velocity, model_split.py:97-103)."""
    a = _cuda_f32(x, "x"); b = _cuda_f32(u, "u")
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    dim = a.shape[-1]
    out = torch.empty_like(a)
    flo = (C.c_float*3)(*[float(v) for v in lo], *([0.0]*(3 - dim)))
    fhi = (C.c_float*3)(*[float(v) for v in hi], *([0.0]*(3 - dim)))
    with torch.cuda.device(a.device):
        _check(_lib().nmc_backtrace(dim, a.data_ptr(), b.data_ptr(), a.numel()//dim, float(dt), flo, fhi, out.data_ptr(), _stream()))
    return out


def mean_squared_error(u, u_ref):
    """mean over samples of ||u - u_ref||^2 (the reference's np.mean(np.linalg.norm(grid_vel - true, axis=-1)**2))."""
    a = _cuda_f32(u, "u"); b = _cuda_f32(u_ref, "u_ref")
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    dim = a.shape[-1]; n = a.numel()//dim
    acc = torch.zeros((), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _check(_lib().nmc_sum_squared_error(dim, a.data_ptr(), b.data_ptr(), n, acc.data_ptr(), _stream()))
    return acc.item()/max(n, 1)


def node_grid(n, scene_size, device):
    """np.indices((N,N)) / N * (size[1]-size[0]) + size[0]  (move_density.py:98-101): [N,N,2], axis 0 <-> x.
    The reference uses the x-extent for both axes (square domains); so does this."""
    i = torch.arange(n, device=device, dtype=torch.float32)
    g = torch.stack(torch.meshgrid(i, i, indexing="ij"), dim=-1)
    return g/n*(scene_size[1] - scene_size[0]) + scene_size[0]


def taylor_green_velocity(samples, scene_size):
    """sources.py:19-32: u = sin x cos y, v = -cos x sin y on the domain rescaled to (0, 2 pi)^2."""
    x = (samples[..., 0] - scene_size[0])/(scene_size[1] - scene_size[0])*(2*math.pi)
    y = (samples[..., 1] - scene_size[2])/(scene_size[3] - scene_size[2])*(2*math.pi)
    return torch.stack([torch.sin(x)*torch.cos(y), -torch.cos(x)*torch.sin(y)], dim=-1)


def taylor_green_error(network, scene_size, n=1000):
    """Velocity error of move_density.py:103-106,124-146: the raw network (no envelope) on the N^2 node grid against
    the steady Taylor-Green field, mean of squared norms."""
    dev = next(network.parameters()).device
    g = node_grid(n, scene_size, dev)
    with torch.no_grad():
        u = network(g.reshape(-1, 2).contiguous())
    dom = (g - scene_size[0])/(scene_size[1] - scene_size[0])*(2*math.pi)   # grid_coords_domain (:100)
    true = torch.stack([torch.sin(dom[..., 0])*torch.cos(dom[..., 1]), -torch.cos(dom[..., 0])*torch.sin(dom[..., 1])], dim=-1)
    return mean_squared_error(u, true.reshape(-1, 2))


class DensityTracker:
    """The density loop of move_density.py: d <- interpolate(d, back-traced node positions) once per time step."""

    def __init__(self, density, scene_size, dt, mode="constant"):
        self.d = _cuda_f32(density, "density").clone()
        self.size, self.dt, self.mode = tuple(float(s) for s in scene_size), float(dt), mode
        n, dim = self.d.shape[0], self.d.dim()
        i = torch.arange(n, device=self.d.device, dtype=torch.float32)
        g = torch.stack(torch.meshgrid(*([i]*dim), indexing="ij"), dim=-1)
        self.nodes = (g/n*(self.size[1] - self.size[0]) + self.size[0]).reshape(-1, dim).contiguous()
        self.lo = [self.size[0]]*dim
        self.extent = [self.size[1] - self.size[0]]*dim

    def step(self, network):
        with torch.no_grad():
            vel = network(self.nodes).reshape(*self.d.shape, self.d.dim())
        self.d = advect_density(self.d, vel, self.dt, self.lo, self.extent, self.mode)
        return self.d
