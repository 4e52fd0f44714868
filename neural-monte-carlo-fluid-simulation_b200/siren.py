"""Fused SIREN velocity network: drop-in for the reference's `MLP` (src/2d/models/networks.py:24-68).

`FusedSiren` keeps the reference module's parameter names and layout (`net.0.weight`, `net.0.bias`,
`net.2.weight`, ... -- an nn.Sequential of nn.Linear and Sine), so `state_dict()` / checkpoints
(src/2d/models/base.py:102-127) are interchangeable, and the same initialisation (sine_init,
first_layer_sine_init).  `forward` runs ONE CUDA kernel (csrc/siren.cu) instead of 2(L+2) stock kernels;
`backward` runs one kernel that produces every parameter gradient and, when the coordinates require grad
(the divergence of src/2d/utils/diff_ops.py:45-51), the input gradient.  Inference batches can use the
tensor-core kernel (csrc/siren_tc.cu) with `tensor_cores=True`.  CUDA tensors only: there is no CPU path.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import capi


class Shape(C.Structure):
    _fields_ = [("in_dim", C.c_int), ("out_dim", C.c_int), ("hidden", C.c_int), ("n_hidden_layers", C.c_int), ("w0", C.c_float)]


class Envelope(C.Structure):
    """nmc_siren_envelope: the boundary envelope of query_velocity evaluated inside the kernels (include/nmcfs_siren.h)."""
    _fields_ = [("kind", C.c_int), ("lo", C.c_float*3), ("hi", C.c_float*3), ("eps", C.c_float),
                ("wall_mask", C.c_int), ("has_sphere", C.c_int), ("sphere_c", C.c_float*3), ("sphere_r", C.c_float),
                ("region_kind", C.c_int), ("region_mask", C.c_int), ("region_lo", C.c_float*3), ("region_hi", C.c_float*3),
                ("region_vel", C.c_float*3), ("sphere_axes", C.c_int), ("region_noise", C.c_float*3), ("noise_seed", C.c_void_p)]


def wall_envelope(size, eps):
    """taylorgreen / vortex_collide (base.py:179-187): wall weights on every component.
    size = (x0, x1, y0, y1[, z0, z1]) as the reference's scene_size."""
    e = Envelope()
    e.kind, e.eps = 1, float(eps)
    for i in range(len(size)//2):
        e.lo[i], e.hi[i] = float(size[2*i]), float(size[2*i + 1])
    return e


def general_envelope(size, eps, wall_mask, sphere=None, region=None, sphere_axes=0, region_noise=None, noise_seed=None):
    """kind 2.  sphere = (centre, radius) of the no-slip obstacle or None; region = ("box", lo, hi, mask, vel) or
    ("ball", centre, radius, mask, vel) or None, `mask` = bit set of the overridden components.  sphere_axes: bit set of
    the coordinates entering the obstacle distance (0 = all).  region_noise: per-component amplitude of the per-sample
    uniform(-1, 1) inlet noise; noise_seed: a uint32 CUDA tensor holding the time step (kept alive by the caller)."""
    e = wall_envelope(size, eps)
    e.kind, e.wall_mask = 2, int(wall_mask)
    e.sphere_axes = int(sphere_axes)
    if region_noise is not None:
        for i, v in enumerate(region_noise):
            e.region_noise[i] = float(v)
    if noise_seed is not None:
        e.noise_seed = register_seed(noise_seed).data_ptr()
    if sphere is not None:
        c, r = sphere
        e.has_sphere, e.sphere_r = 1, float(r)
        for i, v in enumerate(c):
            e.sphere_c[i] = float(v)
    if region is not None:
        kind, a, b, mask, vel = region
        e.region_kind, e.region_mask = {"box": 1, "ball": 2}[kind], int(mask)
        for i, v in enumerate(a):
            e.region_lo[i] = float(v)
        if kind == "box":
            for i, v in enumerate(b):
                e.region_hi[i] = float(v)
        else:
            e.region_hi[0] = float(b)
        for i, v in enumerate(vel):
            e.region_vel[i] = float(v)
    return e


def karman_envelope(size, eps, centre, radius, karman_vel):
    """src/2d/models/base.py:169-181: u = karman_vel in the inlet strip [x0, x0 + eps], no-slip cylinder, wall weight on v."""
    lo = (size[0], -3.0e38)
    hi = (size[0] + eps, 3.0e38)
    return general_envelope(size, eps, wall_mask=0b10, sphere=(centre, radius), region=("box", lo, hi, 0b01, (karman_vel, 0.0)))


def smoke_obs_envelope(size, eps, centre, radius, inlet_centre=(0.0, 0.0, -0.6), inlet_radius=0.1, inlet_w=1.0):
    """src/3d/models/base.py:224-244: w = 1 inside the inlet ball, no-slip sphere obstacle, wall weights on u, v, w."""
    return general_envelope(size, eps, wall_mask=0b111, sphere=(centre, radius),
                            region=("ball", inlet_centre, inlet_radius, 0b100, (0.0, 0.0, inlet_w)))


def karman3d_envelope(size, eps, centre_xz, radius, karman_vel):
    """src/3d/models/base.py:257-275 with the obstacle of src/3d/main.py:92-98: w = karman_vel in the inlet slab
    z in [z0, z0 + eps], no-slip cylinder along y (distance in the x-z plane), wall weights on u and v only."""
    big = 3.0e38
    lo = (-big, -big, size[4])
    hi = (big, big, size[4] + eps)
    return general_envelope(size, eps, wall_mask=0b011, sphere=((centre_xz[0], 0.0, centre_xz[1]), radius),
                            region=("box", lo, hi, 0b100, (0.0, 0.0, karman_vel)), sphere_axes=0b101)


def smoke_envelope(size, eps, noise_seed, inlet_centre=(0.0, 0.0, -0.6), inlet_radius=0.1):
    """src/3d/models/base.py:197-222 (--src smoke): inside the inlet ball (u, v, w) = (0.01 r, 0.01 r, 0.2 + 0.01 r) with
    r uniform in [-10, 10] per sample (numpy re-seeded with the time step), wall weights on u, v, w, no obstacle."""
    return general_envelope(size, eps, wall_mask=0b111, region=("ball", inlet_centre, inlet_radius, 0b111, (0.0, 0.0, 0.2)),
                            region_noise=(0.1, 0.1, 0.1), noise_seed=noise_seed)


_SEEDS = {}


def register_seed(t):
    """Keeps a noise-seed tensor (uint32 / int32, one element, CUDA) alive and findable by its device address."""
    _SEEDS[t.data_ptr()] = t
    return t


def _env_noise_reference(env, samples):
    """The kernels' inlet-noise hash (csrc/siren_env.cuh envNoise) with torch integer ops, for envelope_reference."""
    M = 0xFFFFFFFF
    h0 = 0x7F4A7C15
    if env.noise_seed:
        seed = int(_SEEDS[env.noise_seed].item()) & M
        h0 = (seed*0x9E3779B9 + 0x7F4A7C15) & M
    h = torch.full(samples.shape[:-1], h0, dtype=torch.int64, device=samples.device)
    bits = samples.detach().contiguous().view(torch.int32).to(torch.int64) & M
    for i in range(samples.shape[-1]):
        h = h ^ bits[..., i]
        h = (h*0x85EBCA6B) & M
        h = h ^ (h >> 13)
        h = (h*0xC2B2AE35) & M
        h = h ^ (h >> 16)
    return (h >> 8).to(torch.float32)*(2.0/16777216.0) - 1.0


def envelope_reference(env, samples, net_vel):
    """The same envelope with stock torch ops, written like the reference's query_velocity (autograd flows through the
    obstacle weight, the wall weights are detached).  For tests and for callers who want the un-fused path."""
    if env is None or env.kind == 0:
        return net_vel
    dim = samples.shape[-1]
    eps = env.eps
    vel = net_vel.clone()
    if env.kind == 2 and env.region_kind:
        if env.region_kind == 1:
            m = torch.ones_like(samples[..., 0], dtype=torch.bool)
            for i in range(dim):
                m = m & (samples[..., i] >= env.region_lo[i]) & (samples[..., i] <= env.region_hi[i])
        else:
            c = torch.tensor([env.region_lo[i] for i in range(dim)], device=samples.device, dtype=samples.dtype)
            m = torch.linalg.norm(samples - c, dim=-1) < env.region_hi[0]
        noisy = any(env.region_noise[j] != 0.0 for j in range(3))
        un = _env_noise_reference(env, samples) if noisy else None
        for j in range(vel.shape[-1]):
            if (env.region_mask >> j) & 1:
                val = torch.full_like(vel[..., j], env.region_vel[j])
                if noisy:
                    val = torch.addcmul(val, un, torch.full_like(un, env.region_noise[j]))
                vel[..., j] = torch.where(m, val, vel[..., j])
    if env.kind == 2 and env.has_sphere:
        c = torch.tensor([env.sphere_c[i] for i in range(dim)], device=samples.device, dtype=samples.dtype)
        axes = env.sphere_axes & 7 or 7
        sel = torch.tensor([float((axes >> i) & 1) for i in range(dim)], device=samples.device, dtype=samples.dtype)
        dist = torch.linalg.norm((samples - c)*sel, dim=-1) - env.sphere_r
        vel = vel*(torch.clamp(dist, 0, eps)/eps).unsqueeze(-1)
    mask = 7 if env.kind == 1 else env.wall_mask
    ws = []
    for j in range(vel.shape[-1]):
        if (mask >> j) & 1 and j < dim:
            ws.append(torch.min((samples[..., j] - env.lo[j]).abs().clamp(min=0, max=eps), (samples[..., j] - env.hi[j]).abs().clamp(min=0, max=eps))/eps)
        else:
            ws.append(torch.ones_like(samples[..., 0]))
    return torch.stack(ws, dim=-1).detach()*vel


_configured = False


def _lib():
    global _configured
    L = capi.lib()
    if not _configured:
        vp = C.c_void_p
        L.nmc_siren_last_error.restype = C.c_char_p
        L.nmc_siren_forward.argtypes = [C.POINTER(Shape), C.POINTER(vp), C.POINTER(vp), vp, C.c_int64, vp, vp, C.POINTER(Envelope), vp]
        L.nmc_siren_forward_tc.argtypes = [C.POINTER(Shape), C.POINTER(vp), C.POINTER(vp), vp, C.c_int64, vp, vp, C.POINTER(Envelope), vp]
        L.nmc_siren_backward.argtypes = [C.POINTER(Shape), C.POINTER(vp), C.POINTER(vp), vp, C.c_int64, vp, vp,
                                         vp, vp, vp, C.POINTER(Envelope), vp]
        L.nmc_siren_weight_grads.argtypes = [C.POINTER(Shape), vp, C.c_int64, vp, vp, C.POINTER(vp), C.POINTER(vp), vp]
        L.nmc_siren_backward_tc.argtypes = [C.POINTER(Shape), C.POINTER(vp), C.POINTER(vp), vp, C.c_int64, vp, vp, vp, C.POINTER(Envelope), vp]
        L.nmc_siren_weight_grads_tc.argtypes = [C.POINTER(Shape), vp, C.c_int64, vp, vp, C.POINTER(vp), C.POINTER(vp), vp]
        L.nmc_siren_backward_fused_tc.argtypes = [C.POINTER(Shape), C.POINTER(vp), vp, C.c_int64, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(Envelope), vp]
        L.nmc_adam_step.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, vp]
        L.nmc_adam_step_device.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
        L.nmc_mse_grad.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp]
        L.nmc_mse_grad_fit.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int64, vp, C.c_float, vp, vp]
        L.nmc_adam_update_device.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp]
        f3 = C.POINTER(C.c_float)
        L.nmc_fit_sample_uniform.argtypes = [C.c_int, f3, f3, C.c_int64, vp, vp, vp, C.c_uint64, f3, vp]
        L.nmc_fit_gather.argtypes = [C.c_int, C.c_int64, vp, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_uint64, vp]
        L.nmc_fit_fetch.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        L.nmc_adam_update_fetch.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp,
                                            C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        _configured = True
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError("libnmcfs siren: " + _lib().nmc_siren_last_error().decode())


def _ptrs(tensors):
    return (C.c_void_p*len(tensors))(*[t.data_ptr() for t in tensors])


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _shape_of(weights, w0):
    return Shape(weights[0].shape[1], weights[-1].shape[0], weights[0].shape[0], len(weights) - 2, float(w0))


class _SirenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w0, tensor_cores, need_grad, env, *params):
        n_layers = len(params)//2
        W = [p.contiguous() for p in params[:n_layers]]
        b = [p.contiguous() for p in params[n_layers:]]
        if not x.is_cuda:
            raise RuntimeError("FusedSiren needs CUDA tensors (there is no CPU path)")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
        n = x2.shape[0]
        sh = _shape_of(W, w0)
        y = torch.empty((n, sh.out_dim), device=x.device, dtype=torch.float32)
        L = _lib()
        with torch.cuda.device(x.device):
            if need_grad:
                z = torch.empty(((sh.n_hidden_layers + 1)*sh.hidden, n), device=x.device, dtype=torch.float32)
                if tensor_cores and n >= 16384 and sh.n_hidden_layers >= 1:
                    _check(L.nmc_siren_forward_tc(C.byref(sh), _ptrs(W), _ptrs(b), x2.data_ptr(), n, y.data_ptr(), z.data_ptr(), env, _stream()))
                else:
                    _check(L.nmc_siren_forward(C.byref(sh), _ptrs(W), _ptrs(b), x2.data_ptr(), n, y.data_ptr(), z.data_ptr(), env, _stream()))
                ctx.save_for_backward(x2, z, *W, *b)
            elif tensor_cores and n >= 16384:  # smaller batches: the split fp32 kernel beats the per-tile latency of the tcgen05 one
                _check(L.nmc_siren_forward_tc(C.byref(sh), _ptrs(W), _ptrs(b), x2.data_ptr(), n, y.data_ptr(), None, env, _stream()))
            else:
                _check(L.nmc_siren_forward(C.byref(sh), _ptrs(W), _ptrs(b), x2.data_ptr(), n, y.data_ptr(), None, env, _stream()))
        ctx.w0, ctx.n_layers, ctx.lead, ctx.x_needs, ctx.env = w0, n_layers, lead, ctx.needs_input_grad[0], env
        return y.reshape(*lead, sh.out_dim)

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        x2, z = saved[0], saved[1]
        W = list(saved[2:2 + ctx.n_layers]); b = list(saved[2 + ctx.n_layers:])
        sh = _shape_of(W, ctx.w0)
        n = x2.shape[0]
        gy2 = gy.reshape(n, sh.out_dim).contiguous().float()
        gx = torch.empty_like(x2) if ctx.x_needs else None
        need_params = any(ctx.needs_input_grad[5:])  # False e.g. for the divergence of a frozen network: only dL/dx is wanted
        if not need_params and gx is None:
            return (None,)*(5 + 2*ctx.n_layers)
        dZ, A = _backward_chain(sh, W, b, x2, n, z, gy2, gx, ctx.env, need_params)
        gxr = gx.reshape(*ctx.lead, sh.in_dim) if gx is not None else None
        if not need_params:
            return (gxr, None, None, None, None) + (None,)*(2*ctx.n_layers)
        gW0, gb0, gWh, gbh, gWl, gbl = _param_grads(sh, x2, n, dZ, A)
        gW = [gW0] + list(gWh.unbind(0)) + [gWl]
        gb = [gb0] + list(gbh.unbind(0)) + [gbl]
        return (gxr, None, None, None, None, *gW, *gb)


def _backward_chain(sh, W, b, x2, n, z, gy2, gx, env, need_params=True):
    """One kernel: per-layer deltas dZ and activations A (csrc/siren.cu sirenBackwardChain / sirenBackwardSplit).
    need_params=False: only dL/dx is produced (no dZ / A buffers: 2 * (L+1) * H * n floats not written)."""
    rows = (sh.n_hidden_layers + 1)*sh.hidden
    dZ = A = None
    if need_params:
        dZ = torch.empty((rows + sh.out_dim, n), device=x2.device, dtype=torch.float32)
        A = torch.empty((rows, n), device=x2.device, dtype=torch.float32)
    with torch.cuda.device(x2.device):
        _check(_lib().nmc_siren_backward(C.byref(sh), _ptrs(W), _ptrs(b), x2.data_ptr(), n, z.data_ptr(), gy2.data_ptr(),
                                         dZ.data_ptr() if need_params else None, A.data_ptr() if need_params else None,
                                         gx.data_ptr() if gx is not None else None, env, _stream()))
    return dZ, A


def _param_grads(sh, x2, n, dZ, A, out=None):
    """All weight / bias gradients in one launch (csrc/siren.cu sirenWeightGrad): batch-dimension GEMMs
    dW_0 = dZ_0 x, dW_l = dZ_l A_{l-1}^T, dW_last = gy'^T A_L^T, db_l = rowsum(dZ_l), accumulated with atomics
    into zero-filled buffers.  `out` = (gW0, gb0, gWh[L,H,H], gbh[L,H], gWl, gbl) already zeroed, or None."""
    H, Lh = sh.hidden, sh.n_hidden_layers
    if out is None:
        dev = x2.device
        out = (torch.zeros((H, sh.in_dim), device=dev), torch.zeros(H, device=dev), torch.zeros((Lh, H, H), device=dev),
               torch.zeros((Lh, H), device=dev), torch.zeros((sh.out_dim, H), device=dev), torch.zeros(sh.out_dim, device=dev))
    gW0, gb0, gWh, gbh, gWl, gbl = out
    gW = [gW0] + [gWh[i] for i in range(Lh)] + [gWl]
    gb = [gb0] + [gbh[i] for i in range(Lh)] + [gbl]
    with torch.cuda.device(x2.device):
        _check(_lib().nmc_siren_weight_grads(C.byref(sh), x2.data_ptr(), n, dZ.data_ptr(), A.data_ptr(), _ptrs(gW), _ptrs(gb), _stream()))
    return out


class Sine(nn.Module):
    """Placeholder with the reference's name so that `net` has the same module indices (networks.py:15-21)."""

    def forward(self, input):
        return torch.sin(30*input)


def sine_init(m):  # networks.py:78-83
    with torch.no_grad():
        if hasattr(m, "weight"):
            num_input = m.weight.size(-1)
            m.weight.uniform_(-np.sqrt(6/num_input)/30, np.sqrt(6/num_input)/30)


def first_layer_sine_init(m):  # networks.py:85-90
    with torch.no_grad():
        if hasattr(m, "weight"):
            num_input = m.weight.size(-1)
            m.weight.uniform_(-1/num_input, 1/num_input)


class FusedSiren(nn.Module):
    """MLP(in_features, out_features, num_hidden_layers, hidden_features, nonlinearity='sine') with fused kernels."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features, outermost_linear=True,
                 nonlinearity="sine", weight_init=None, tensor_cores=False):
        super().__init__()
        if nonlinearity != "sine" or not outermost_linear:
            raise NotImplementedError("the time-stepper only uses nonlinearity='sine' with a linear last layer (run.sh)")
        if hidden_features not in (64, 128):
            raise NotImplementedError("hidden_features must be 64 or 128 (the shipped configs)")
        self.weight_init = weight_init if weight_init is not None else sine_init
        self.first_layer_init = None if weight_init is not None else first_layer_sine_init
        layers = [nn.Linear(in_features, hidden_features), Sine()]
        for _ in range(num_hidden_layers):
            layers.extend([nn.Linear(hidden_features, hidden_features), Sine()])
        layers.append(nn.Linear(hidden_features, out_features))
        self.net = nn.Sequential(*layers)
        self.net.apply(self.weight_init)
        if self.first_layer_init is not None:
            self.net[0].apply(self.first_layer_init)
        self.tensor_cores = tensor_cores

    def _linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def forward(self, coords, weights=None, envelope=None):
        """envelope: optional siren.Envelope fused into the kernels (detached weights, as in the reference)."""
        lin = self._linears()
        # grad mode is off inside autograd.Function.forward, so decide here whether activations must be saved
        need_grad = torch.is_grad_enabled() and (coords.requires_grad or any(p.requires_grad for p in self.parameters()))
        out = _SirenFn.apply(coords, 30.0, self.tensor_cores, need_grad, envelope, *[m.weight for m in lin], *[m.bias for m in lin])
        if weights is not None:
            out = out*weights
        return out

    def forward_reference(self, coords):
        """The reference's own evaluation (stock PyTorch ops) -- used by the parity tests."""
        return self.net(coords)


class FusedAdam:
    """torch.optim.Adam semantics (lr, betas=(0.9, 0.999), eps=1e-8) with one kernel per step over a flat copy
    of the parameters' gradients (base.py:61-77 creates torch.optim.Adam with default betas/eps)."""

    def __init__(self, params, lr=1e-5, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.step_count = lr, betas, eps, 0
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev); self.m = torch.zeros(n, device=dev); self.v = torch.zeros(n, device=dev)
        self.g = torch.zeros(n, device=dev)
        self.step_dev = torch.zeros((), dtype=torch.int64, device=dev)  # Adam's t lives on the device (CUDA-graph replay)
        self.stop_flag = torch.zeros((), dtype=torch.int32, device=dev)  # set by the loss kernel at the early-stop threshold (DirectFit.stop_threshold)
        with torch.no_grad():  # re-home the parameters as views into one flat buffer
            off = 0
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k] = p.reshape(-1)
                p.data = self.flat[off:off + k].view_as(p)
                off += k

        self.grad_views = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.grad_views.append(self.g[off:off + k].view_as(p))
            off += k

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def reset(self):
        """A fresh optimizer on the same parameters (create_optimizer() at the top of every fit, base.py:133)."""
        self.m.zero_(); self.v.zero_(); self.step_dev.zero_(); self.stop_flag.zero_()
        self.step_count = 0

    def step_flat(self):
        """Adam step when the gradients were written straight into `grad_views` (no autograd).  The step counter is
        advanced on the device, so the call can be captured once in a CUDA graph and replayed."""
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            _check(_lib().nmc_adam_step_device(self.flat.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat.numel(),
                                               self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(), _stream()))

    def update_flat(self, gated=False):
        """step_flat without the increment: the counter was advanced by nmc_mse_grad_fit earlier in the iteration.
        gated: no update once the loss kernel has set `stop_flag` (the fit reached the early-stop threshold)."""
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            _check(_lib().nmc_adam_update_device(self.flat.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat.numel(),
                                                 self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(),
                                                 self.stop_flag.data_ptr() if gated else None, _stream()))

    def update_flat_fetch(self, gated, ring_x, ring_t, ring_s, out_x, out_t, out_s):
        """update_flat and the fit_fetch of the NEXT iteration in one launch (csrc/siren.cu: adamFetchKernel)."""
        self.step_count += 1
        slots, count = ring_x.shape[0], out_x.numel()
        assert ring_x.is_contiguous() and ring_t.is_contiguous() and ring_x.numel() == slots*count and ring_t.shape == ring_x.shape
        with torch.cuda.device(self.flat.device):
            _check(_lib().nmc_adam_update_fetch(self.flat.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat.numel(),
                                                self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(),
                                                self.stop_flag.data_ptr() if gated else None, count, slots, ring_x.data_ptr(), ring_t.data_ptr(),
                                                ring_s.data_ptr() if ring_s is not None else None, out_x.data_ptr(), out_t.data_ptr(),
                                                out_s.data_ptr() if ring_s is not None else None, _stream()))

    def step(self):
        self.step_count += 1
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is not None:
                self.g[off:off + k] = p.grad.reshape(-1)
            else:
                self.g[off:off + k].zero_()
            off += k
        with torch.cuda.device(self.flat.device):
            _check(_lib().nmc_adam_step_device(self.flat.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat.numel(),
                                               self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(), _stream()))


def fit_sample_uniform(n, lo, hi, step_dev, epoch_dev, seed, obstacle=None, out=None):
    """One launch: n points uniform in the box [lo, hi) (sample_in_training 'random', base.py:225-241); obstacle = (centre, radius):
    a point inside the ball is redrawn once.  The draw is keyed by (seed, epoch, step) read from device memory at run time, so the
    call can be captured in a CUDA graph (csrc/fit_glue.cu)."""
    dim = len(lo)
    dev = step_dev.device
    if out is None:
        out = torch.empty((n, dim), device=dev)
    f = C.c_float*dim
    obs = None
    if obstacle is not None:
        obs = (C.c_float*(dim + 1))(*[float(v) for v in obstacle[0]], float(obstacle[1]))
    with torch.cuda.device(dev):
        _check(_lib().nmc_fit_sample_uniform(dim, f(*[float(v) for v in lo]), f(*[float(v) for v in hi]), n, out.data_ptr(), step_dev.data_ptr(),
                                             epoch_dev.data_ptr(), int(seed) & (2**64 - 1), obs, _stream()))
    return out


def fit_gather(n, src_x, src_g, count_dev, step_dev, epoch_dev, seed):
    """One launch: the projection fit's batch, rows idx = floor(u * count) of the pressure samples and of grad p
    (model_split.py:272-277)."""
    dim = src_x.shape[1]
    assert src_x.is_contiguous() and src_g.is_contiguous() and src_g.shape == src_x.shape and count_dev.dtype == torch.float32
    out_x = torch.empty((n, dim), device=src_x.device); out_g = torch.empty((n, dim), device=src_x.device)
    with torch.cuda.device(src_x.device):
        _check(_lib().nmc_fit_gather(dim, n, src_x.data_ptr(), src_g.data_ptr(), count_dev.data_ptr(), src_x.shape[0], out_x.data_ptr(), out_g.data_ptr(),
                                     step_dev.data_ptr(), epoch_dev.data_ptr(), int(seed) & (2**64 - 1), _stream()))
    return out_x, out_g


def fit_fetch(ring_x, ring_t, ring_s, step_dev, out_x, out_t, out_s):
    """One launch: slot (step % slots) of the target ring ([slots, n, dim] each) -> the fixed buffers of a captured iteration."""
    slots = ring_x.shape[0]
    count = out_x.numel()
    assert ring_x.is_contiguous() and ring_t.is_contiguous() and ring_x.numel() == slots*count and ring_t.shape == ring_x.shape
    with torch.cuda.device(out_x.device):
        _check(_lib().nmc_fit_fetch(count, slots, ring_x.data_ptr(), ring_t.data_ptr(), ring_s.data_ptr() if ring_s is not None else None,
                                    step_dev.data_ptr(), out_x.data_ptr(), out_t.data_ptr(), out_s.data_ptr() if ring_s is not None else None, _stream()))


class DirectFit:
    """One Adam iteration of an MSE fit without autograd: forward (saving pre-activations), dL/dy, delta-chain
    kernel, batched GEMMs straight into one flat gradient buffer, one Adam kernel.  Replaces update_network
    (base.py:83-96).  Flat layout: W_0, b_0, W_1..W_L (one [L,H,H] block), b_1..b_L, W_last, b_last."""

    def __init__(self, net, lr, envelope=None, max_batch=16384, group=None, distributed=False):
        """distributed: data-parallel fit over the ranks of `group` (torch.distributed, NCCL): every rank passes its
        own shard of the batch to iterate(); the flat gradient buffer (21-34 k floats for the shipped shapes) is
        averaged with ONE all_reduce per iteration and every rank applies the same Adam step.  The parameters are
        broadcast from rank 0 by sync_parameters()."""
        self.net, self.env = net, envelope
        self.tensor_cores = bool(getattr(net, "tensor_cores", False))
        self.group, self.world = group, 1
        if distributed:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        lin = net._linears()
        self.W = [m.weight for m in lin]; self.b = [m.bias for m in lin]
        order = [self.W[0], self.b[0]] + self.W[1:-1] + self.b[1:-1] + [self.W[-1], self.b[-1]]
        self.opt = FusedAdam(order, lr=lr)
        self.sh = _shape_of(self.W, 30.0)
        H, Lh, g = self.sh.hidden, self.sh.n_hidden_layers, self.opt.g
        off = [0]

        def take(shape):
            k = 1
            for v in shape:
                k *= v
            t = g[off[0]:off[0] + k].view(*shape); off[0] += k
            return t
        self.out = (take((H, self.sh.in_dim)), take((H,)), take((Lh, H, H)), take((Lh, H)), take((self.sh.out_dim, H)), take((self.sh.out_dim,)))
        assert off[0] == g.numel()
        self.z = torch.empty((Lh + 1)*H*max_batch, device=g.device)
        # tensor-core backward (delta chain + weight gradients): on with the network's tensor_cores flag; NMC_SIREN_TC_BWD=0
        # selects the fp32 kernels for A/B measurements
        self.tc_backward = self.tensor_cores and os.environ.get("NMC_SIREN_TC_BWD", "1") != "0"
        self.tc_backward_min = int(os.environ.get("NMC_SIREN_TC_BWD_MIN", "4096"))
        self.fused_backward = os.environ.get("NMC_SIREN_FUSED_BWD", "1") != "0"  # hidden = 64: the one-kernel backward (NMC_SIREN_FUSED_BWD=0: two kernels)
        # one tile chain per CTA: pays once the batch fills the GPU (>= 128 tiles); below, the weight-gradient kernel's layer-parallel grid wins
        self.fused_backward_min = int(os.environ.get("NMC_SIREN_FUSED_BWD_MIN", "16384"))
        self.tc_forward_min = int(os.environ.get("NMC_SIREN_TC_FWD_MIN", "16384"))
        self.dz = torch.empty(((Lh + 1)*H + self.sh.out_dim)*max_batch, device=g.device) if self.tc_backward else None
        self.max_batch = max_batch
        self.loss = torch.zeros((), device=g.device)  # mean squared error of the last iterate() call
        # > 0: the reference's early stop (base.py:148, loss <= 1.1e-10) decided on the device -- the loss kernel sets opt.stop_flag,
        # later updates are skipped until opt.reset(); the host reads the flag whenever it likes (stepper._stop_now)
        self.stop_threshold = 0.0

    def iterate(self, x, target, sub=None):
        return self.finish(x, self.forward(x), target, sub)

    def forward(self, x):
        """First half of an iteration: the training forward (saves the pre-activations).  Independent of the target, so a
        caller may compute the target on another stream meanwhile (stepper.py)."""
        n = x.shape[0]
        assert n <= self.max_batch and x.is_contiguous()
        sh = self.sh
        y = torch.empty((n, sh.out_dim), device=x.device)
        z = self.z[: (sh.n_hidden_layers + 1)*sh.hidden*n]  # [layer*H + neuron][sample], stride n
        with torch.cuda.device(x.device):
            if self.tensor_cores and n >= self.tc_forward_min and sh.n_hidden_layers >= 1:  # tcgen05 forward that also writes the pre-activations
                _check(_lib().nmc_siren_forward_tc(C.byref(sh), _ptrs(self.W), _ptrs(self.b), x.data_ptr(), n, y.data_ptr(), z.data_ptr(), self.env, _stream()))
            else:
                _check(_lib().nmc_siren_forward(C.byref(sh), _ptrs(self.W), _ptrs(self.b), x.data_ptr(), n, y.data_ptr(), z.data_ptr(), self.env, _stream()))
        return y

    def finish(self, x, y, target, sub=None, fetch_next=None):
        """Second half: loss, dL/dy, delta chain, weight gradients, (all_reduce,) Adam.  Returns y - target.
        The fit target is target - sub when `sub` is given (the projection fit's u_prev - grad p).
        fetch_next = (ring_x, ring_t, ring_s, out_x, out_t, out_s): the Adam launch also fetches the next iteration's ring slot."""
        n = x.shape[0]
        sh = self.sh
        z = self.z[: (sh.n_hidden_layers + 1)*sh.hidden*n]
        diff = torch.empty_like(y); gy = torch.empty_like(y)
        tgt = target.contiguous()
        gated = self.stop_threshold > 0 and self.world == 1   # data-parallel fits: the decision is collective (stepper.collective_stop)
        if sub is not None:
            sub = sub.contiguous()
            assert sub.shape == tgt.shape
        with torch.cuda.device(x.device):  # diff, dL/dy, the loss, the zero-fill of the gradient buffer and Adam's t += 1 in one launch
            _check(_lib().nmc_mse_grad_fit(y.data_ptr(), tgt.data_ptr(), sub.data_ptr() if sub is not None else None, y.numel(), diff.data_ptr(),
                                           gy.data_ptr(), self.loss.data_ptr(), self.opt.g.data_ptr(), self.opt.g.numel(), self.opt.step_dev.data_ptr(),
                                           float(self.stop_threshold), self.opt.stop_flag.data_ptr() if gated else None, _stream()))
        if self.tc_backward and n >= self.tc_backward_min and n % 4 == 0 and sh.n_hidden_layers >= 1:
            gW0, gb0, gWh, gbh, gWl, gbl = self.out
            gW = [gW0] + [gWh[i] for i in range(sh.n_hidden_layers)] + [gWl]
            gb = [gb0] + [gbh[i] for i in range(sh.n_hidden_layers)] + [gbl]
        if self.tc_backward and self.fused_backward and n >= self.fused_backward_min and sh.hidden == 64 and 1 <= sh.n_hidden_layers <= 6:
            # one kernel: delta chain and all gradients, deltas / activations in shared memory (csrc/siren_tc_fused_bwd.cu)
            with torch.cuda.device(x.device):
                _check(_lib().nmc_siren_backward_fused_tc(C.byref(sh), _ptrs(self.W), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(),
                                                          _ptrs(gW), _ptrs(gb), self.env, _stream()))
        elif self.tc_backward and n >= self.tc_backward_min and n % 4 == 0 and sh.n_hidden_layers >= 1:
            # tcgen05 delta chain + weight gradients (csrc/siren_tc_bwd.cu); the activations are recomputed from z
            dZ = self.dz[: ((sh.n_hidden_layers + 1)*sh.hidden + sh.out_dim)*n]
            with torch.cuda.device(x.device):
                _check(_lib().nmc_siren_backward_tc(C.byref(sh), _ptrs(self.W), _ptrs(self.b), x.data_ptr(), n, z.data_ptr(), gy.data_ptr(),
                                                    dZ.data_ptr(), self.env, _stream()))
                _check(_lib().nmc_siren_weight_grads_tc(C.byref(sh), x.data_ptr(), n, dZ.data_ptr(), z.data_ptr(), _ptrs(gW), _ptrs(gb), _stream()))
        else:
            dZ, A = _backward_chain(sh, self.W, self.b, x, n, z, gy, None, self.env)
            _param_grads(sh, x, n, dZ, A, out=self.out)
        if self.world > 1:  # mean over the global batch = mean over ranks of the local means (equal shard sizes)
            import torch.distributed as dist
            dist.all_reduce(self.opt.g, op=dist.ReduceOp.AVG, group=self.group)
        if fetch_next is not None:
            self.opt.update_flat_fetch(gated, *fetch_next)
        else:
            self.opt.update_flat(gated)
        return diff

    def close(self):
        """Releases the scratch buffers; the network keeps its (flat-backed) parameters."""
        self.z = None
        self.dz = None
        self.out = None

    def sync_parameters(self, src=0):
        """Broadcast the flat parameter buffer from rank `src` (after initialisation / re-initialisation)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(self.opt.flat, src=src, group=self.group)
