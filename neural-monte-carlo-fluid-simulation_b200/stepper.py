"""Device-resident operator-split time step: advect fit -> divergence grid -> Monte Carlo pressure solve ->
projection fit, with every array staying in HBM.

Mirrors the reference's NeuralFluidSplit.step for adv_ref = 0 (src/2d/models/model_split.py:44-62):
    _advect_velocity   model_split.py:88-120      semi-Lagrangian backtrace, MSE fit by Adam
    get_divergence     model_split.py:230-243     -div u on the (res+2)^2 grid, autograd w.r.t. the coordinates
    wost_pressure      model_split.py:185-228     Scene(sceneConfig, div) + wost(...)
    _project_velocity  model_split.py:246-284     fit u <- u_prev - grad p on the pressure samples
    query_velocity     base.py:158-224            network + boundary envelope (taylorgreen branch :179-187)
    _training_loop     base.py:129-152            max_n_iters Adam iterations, early stop at loss <= 1.1e-10
What changes is where the data lives: the divergence grid is written into the scene with
nmc_scene_set_source(src_is_device = 1), the solve runs through nmc_wost_solve_device on torch's stream and
grad p never leaves the GPU (the reference round-trips all three through numpy / Python lists,
model_split.py:194, :272).  The boundary structure is built once, not once per step (:191).
One fit iteration = 2 no-grad forwards + 1 forward/backward of the fused SIREN kernels + 1 fused Adam kernel,
optionally replayed from a CUDA graph (no per-iteration host synchronisation; the reference calls loss.item()
every iteration, base.py:142 -- the early-stop test here runs every `check_every` iterations).
"""
import ctypes as C

import os

import torch

from . import capi, zombie, fields
from .siren import (FusedSiren, DirectFit, fit_sample_uniform, fit_gather, fit_fetch, wall_envelope, karman_envelope, smoke_obs_envelope, karman3d_envelope, smoke_envelope,
                    envelope_reference)


def sample_uniform_2d(resolution, size, device, with_boundary=True):
    """utils/model_utils.py:3-20 (normalize=True, meshgrid indexing 'xy': result[i][j] = (x_j, y_i))."""
    if (size[1] - size[0]) > (size[3] - size[2]):
        res_x, res_y = resolution, int(resolution*(size[3] - size[2])/(size[1] - size[0]))
    else:
        res_x, res_y = int(resolution*(size[1] - size[0])/(size[3] - size[2])), resolution
    x = torch.linspace(0.5, res_x - 0.5, res_x, device=device)
    y = torch.linspace(0.5, res_y - 0.5, res_y, device=device)
    if with_boundary:
        x = torch.cat([torch.tensor([0.0], device=device), x, torch.tensor([res_x*1.0], device=device)])
        y = torch.cat([torch.tensor([0.0], device=device), y, torch.tensor([res_y*1.0], device=device)])
    coords = torch.stack(torch.meshgrid(x, y, indexing="xy"), dim=-1)
    coords[..., 0] = coords[..., 0]/res_x*(size[1] - size[0]) + size[0]
    coords[..., 1] = coords[..., 1]/res_y*(size[3] - size[2]) + size[2]
    return coords


def sample_uniform_3d(resolution, size, device, with_boundary=True):
    """src/3d/utils/model_utils.py:3-34 (meshgrid indexing 'ij': result[i][j][k] = (x_i, y_j, z_k), the layout
    zombie3d's Scene expects).  The reference sizes the z axis with res_y samples (:19); for the cubic domains of
    all 3D examples the three resolutions coincide, which is what this implements."""
    ext = [size[1] - size[0], size[3] - size[2], size[5] - size[4]]
    m = min(ext)
    res = [int(resolution*e/m) for e in ext]
    axes = []
    for r in res:
        a = torch.linspace(0.5, r - 0.5, r, device=device)
        if with_boundary:
            a = torch.cat([torch.tensor([0.0], device=device), a, torch.tensor([r*1.0], device=device)])
        axes.append(a)
    coords = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    for k in range(3):
        coords[..., k] = coords[..., k]/res[k]*ext[k] + size[2*k]
    return coords


def collective_stop(loss, world, group=None, threshold=1.1e-10):
    """True when the mean over ranks of `loss` (a scalar tensor, each rank's shard MSE) is at or below the
    reference's early-stop threshold (base.py:148).  One all_reduce, issued by every rank."""
    if world > 1:
        import torch.distributed as dist
        loss = loss.detach().clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        loss = loss/world
    return loss.item() <= threshold


class SplitStepper:
    """Operator-split stepper on a box domain `scene_size` = (x0, x1, y0, y1[, z0, z1]); 2D or 3D by its length.
    boundary = 'taylorgreen' / 'walls': wall weights on every component (taylorgreen, vortex_collide branches);
    'karman' (2D): inlet strip, no-slip cylinder `obstacle` = (centre, radius), wall weight on v, samples inside the
    cylinder unused (base.py:169-181, 239-241); 'smoke_obs' (3D): inlet ball, no-slip sphere `obstacle`, wall weights
    (3d base.py:224-244); 'karman3d' (3D): inlet slab on w, no-slip cylinder along y `obstacle` = ((cx, cz), radius), wall
    weights on u and v (3d base.py:257-275, main.py:92-98); 'smoke' (3D): noisy inlet ball re-drawn every time step, wall
    weights (3d base.py:197-222); None: no envelope.  reset_wts re-initialises the network before every fit
    (create_optimizer(reset=True), model_split.py:44-62; the karman and 3D examples run with --reset_wts 1)."""

    def __init__(self, wost_config, scene_size, hidden_features=64, num_hidden_layers=6, dt=0.001, lr=1e-5,
                 sample_resolution=64, wost_resolution=512, grid_resolution=1000, bdry_eps=1e-3, max_n_iters=10000,
                 early_stop=True, check_every=100, boundary="taylorgreen", mode=capi.MODE_FAST, seed=0, device=0,
                 use_cuda_graph=True, tensor_cores=True, init_velocity=None, init_iters=0, obstacle=None, karman_vel=0.5,
                 reset_wts=False, distributed=False, fit_parallel="data"):
        """distributed: one process per GPU under torch.distributed (NCCL).  The fits are data-parallel (each rank
        draws 1/world of every batch, one gradient all_reduce per iteration), the pressure samples are drawn and
        solved per rank and all-gathered (points, p, grad p), the divergence grid and the networks are replicated.
        Every rank ends a step with identical weights."""
        self.dev = torch.device("cuda", device)
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # fit_parallel (distributed runs): "data" = data-parallel fits, every rank 1/world of each batch and one gradient
        # all_reduce per iteration; "replicated" = every rank runs the whole fit on identical samples (the same random
        # streams on all ranks) and only the pressure solve is sharded -- the better choice while a fit iteration is
        # latency-bound (measured at 2 GPUs, smoke3d, K = 1000: 2.9 steps/s data-parallel against 4.2 on one GPU).  The
        # weights are re-broadcast from rank 0 after every fit (the gradient reductions use atomics: last-bit differences).
        if fit_parallel not in ("data", "replicated"):
            raise ValueError("fit_parallel must be 'data' or 'replicated'")
        self.fit_world = self.world if fit_parallel == "data" else 1
        self.cfg = wost_config
        self.size = tuple(float(v) for v in scene_size)
        self.dt, self.lr, self.eps = dt, lr, bdry_eps
        self.sample_resolution, self.wost_resolution, self.grid_resolution = sample_resolution, wost_resolution, grid_resolution
        self.max_n_iters, self.early_stop, self.check_every = max_n_iters, early_stop, check_every
        self.boundary, self.use_graph = boundary, use_cuda_graph
        # fit target on a second stream, parallel to the training forward (NMC_OVERLAP_TARGETS=0: one stream, for A/B runs)
        self.overlap_targets = os.environ.get("NMC_OVERLAP_TARGETS", "1") != "0"
        # captured iterations draw their batches with one launch keyed by device-side counters (csrc/fit_glue.cu) instead of
        # three to eight torch launches; without CUDA graphs the batches come from torch's generator (NMC_FIT_GLUE=0: always)
        self.fused_glue = bool(use_cuda_graph) and os.environ.get("NMC_FIT_GLUE", "1") != "0"
        # fit targets of the next `target_chunk` iterations computed in one pass on a second stream (NMC_TARGET_CHUNK=0: inside the iteration)
        self.target_chunk = int(os.environ.get("NMC_TARGET_CHUNK", "40"))
        # with chunked targets an iteration is one stream of six kernels; `graph_unroll` of them are captured in a second graph, which
        # removes the ~4 us between two graph launches (taylorgreen: 68 -> 64 us per iteration).  Must divide the chunk and check_every.
        self.graph_unroll = int(os.environ.get("NMC_GRAPH_UNROLL", "20"))
        # inside the unrolled graph the Adam update of iteration i and the ring fetch of iteration i + 1 are one launch (NMC_MERGE_ADAM_FETCH=0: two)
        self.merge_adam_fetch = os.environ.get("NMC_MERGE_ADAM_FETCH", "1") != "0"
        torch.manual_seed(seed)
        self.dim = dim = len(self.size)//2
        if dim not in (2, 3):
            raise ValueError("scene_size must have 4 (2D) or 6 (3D) entries")
        mk = lambda: FusedSiren(dim, dim, num_hidden_layers, hidden_features, nonlinearity="sine", tensor_cores=tensor_cores).to(self.dev)  # noqa: E731
        self.velocity_field, self.velocity_field_prev = mk(), mk()
        for p in self.velocity_field_prev.parameters():
            p.requires_grad_(False)
        # Scene(config, div): built once; the source grid is replaced every step
        grid = (sample_uniform_2d if dim == 2 else sample_uniform_3d)(grid_resolution, self.size, self.dev)
        self.grid_shape = tuple(grid.shape[:dim])
        self.grid_samples = grid.reshape(-1, dim).contiguous()
        self.scene = zombie.Scene(wost_config["scene"], torch.zeros(self.grid_shape).numpy(), device=device)
        self.opts = zombie.solver_opts(wost_config["solver"], wost_config["output"], mode=mode, seed=seed)
        self.timestep, self.seed = 0, seed
        self.last = {}
        self._fit, self._graphs, self._proj = None, {}, None
        self._rings = {}
        self._epoch = torch.zeros((), dtype=torch.int64, device=self.dev)  # fits started so far: part of the key of the captured draws
        if self.world > 1:  # identical initial weights (same seed above); data-parallel fits: different training samples per rank from here on
            torch.manual_seed(seed*7919 + 1 + (self.rank if self.fit_world > 1 else 0))
        self.obstacle, self.karman_vel, self.reset_wts = obstacle, float(karman_vel), bool(reset_wts)
        if boundary in ("taylorgreen", "walls"):
            self.env = wall_envelope(self.size, bdry_eps)
        elif boundary in ("karman", "smoke_obs"):
            if obstacle is None or dim != (2 if boundary == "karman" else 3):
                raise ValueError("boundary='karman' (2D) / 'smoke_obs' (3D) needs obstacle=(centre, radius)")
            if boundary == "karman":
                self.env = karman_envelope(self.size, bdry_eps, obstacle[0], obstacle[1], karman_vel)
            else:
                self.env = smoke_obs_envelope(self.size, bdry_eps, obstacle[0], obstacle[1])
            self._obs_c = torch.tensor([float(c) for c in obstacle[0]], device=self.dev)
        elif boundary == "karman3d":
            if obstacle is None or dim != 3:
                raise ValueError("boundary='karman3d' needs obstacle=((cx, cz), radius) and a 3D scene")
            self.env = karman3d_envelope(self.size, bdry_eps, obstacle[0], obstacle[1], karman_vel)
        elif boundary == "smoke":
            if dim != 3:
                raise ValueError("boundary='smoke' is the 3D smoke plume")
            # the reference re-seeds numpy with the time step on every query (3d base.py:203); the kernels read this cell
            self._noise_seed = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.env = smoke_envelope(self.size, bdry_eps, self._noise_seed)
        elif boundary is None:
            self.env = None
        else:
            raise ValueError("boundary must be 'taylorgreen', 'walls', 'karman', 'smoke_obs', 'karman3d', 'smoke' or None")
        self._lo = torch.tensor(self.size[0::2], device=self.dev)
        self._hi = torch.tensor(self.size[1::2], device=self.dev)
        if init_velocity is not None and init_iters > 0:
            self.fit_initial(init_velocity, init_iters)

    # ---- network + envelope (base.py:158-224, taylorgreen branch) ---------------------------------------------
    def envelope(self, samples):
        """Multiplicative wall weights of the taylorgreen branch as a tensor (tests); see apply_envelope_reference."""
        if self.boundary not in ("taylorgreen", "walls"):
            return None
        s, e = self.size, self.eps
        w = [torch.min((samples[..., k] - s[2*k]).abs().clamp(min=0, max=e), (samples[..., k] - s[2*k + 1]).abs().clamp(min=0, max=e))/e
             for k in range(self.dim)]
        return torch.stack(w, dim=-1).detach()

    def query_velocity(self, samples, use_prev=False):
        """network x envelope in ONE kernel (the envelope is fused: include/nmcfs_siren.h nmc_siren_envelope)."""
        net = self.velocity_field_prev if use_prev else self.velocity_field
        return net(samples, envelope=self.env)

    def apply_envelope_reference(self, samples, net_vel):
        """query_velocity's boundary treatment with stock torch ops (what the fused kernels replace)."""
        return envelope_reference(self.env, samples, net_vel)

    def obstacle_distance(self, samples):
        """circle_obstable_functions (main.py:101-104): signed distance to the cylinder."""
        return torch.linalg.norm(samples - self._obs_c, dim=-1) - self.obstacle[1]

    def sample_random(self, n, keep_shape=True, fused=False, seed_xor=0, out=None):
        """sample_in_training with the 'random' pattern (base.py:225-241, utils/model_utils.py:22-31).  With an
        obstacle the reference drops the samples inside it (a batch a fraction of a percent smaller); here a sample
        inside is redrawn once so that the batch keeps its shape (CUDA-graph replay), or, with keep_shape=False,
        dropped exactly like the reference."""
        if fused and keep_shape:  # inside a captured fit iteration only: the key is the fit's (epoch, iteration) pair
            # data-parallel fits: every rank its own stream; replicated fits: the same samples on every rank
            seed = (self.seed*0x9E3779B1 + 0x632BE5AB*(self.rank if self.fit_world > 1 else 0)) ^ seed_xor
            return fit_sample_uniform(n, self.size[0::2], self.size[1::2], self._fit.opt.step_dev, self._epoch, seed,
                                      obstacle=self.obstacle if self.boundary == "karman" else None, out=out)
        x = torch.rand(n, self.dim, device=self.dev)*(self._hi - self._lo) + self._lo
        if self.boundary == "karman":
            if keep_shape:
                x2 = torch.rand(n, self.dim, device=self.dev)*(self._hi - self._lo) + self._lo
                x = torch.where((self.obstacle_distance(x) > 0).unsqueeze(-1), x, x2)
            else:
                x = x[self.obstacle_distance(x) > 0]
        return x

    # ---- fit loops ---------------------------------------------------------------------------------------------
    def _reset_weights(self):
        net = self.velocity_field
        with torch.no_grad():
            net.net.apply(net.weight_init)
            if net.first_layer_init is not None:
                net.net[0].apply(net.first_layer_init)

    def _loop(self, iteration, n_iters, key=None, chunked=None):
        """_training_loop (base.py:129-152) without autograd and without a host sync per iteration:
        `iteration()` returns (samples, target); the MSE fit step is DirectFit.iterate.  One DirectFit (one flat
        parameter / Adam buffer) serves every fit and is reset where the reference creates a new optimizer
        (base.py:133).  With CUDA graphs the iteration of each phase (`key`) is captured ONCE and replayed in every
        later time step: its inputs live in buffers that persist across steps, and Adam's step counter is on the
        device.  (Capturing per fit costs a synchronise + allocator flush per phase, i.e. tens to hundreds of ms.)"""
        if self._fit is None:
            self._fit = DirectFit(self.velocity_field, self.lr, self.env, max_batch=self.sample_resolution**2, distributed=self.fit_world > 1)
        fit = self._fit
        fit.stop_threshold = 1.1e-10 if (self.early_stop and self.fit_world == 1) else 0.0
        fit.opt.reset()
        self._epoch += 1
        if self.reset_wts:
            self._reset_weights()
            fit.sync_parameters()  # the re-initialisation draws from per-rank random streams
        cached = self._graphs.get(key) if (self.use_graph and key is not None) else None
        it = 0
        # Targets ahead of time (`chunked` = gen(chunk id, m) -> (samples, target, sub or None), graph mode): the target of an
        # iteration depends only on the frozen previous network and on the batch, so the batches and targets of the next
        # `target_chunk` iterations are computed in ONE pass on a second stream -- at 32 x 4096 samples the tcgen05 forward runs at its
        # large-batch rate (2.3 us per 4096 samples against 18.7 us for a launch of its own) -- into a ring of two chunks; the
        # captured iteration fetches its slot (Adam's device-side step modulo the ring size) and is one stream: forward, loss,
        # backward, Adam.  taylorgreen advect: 97 us per iteration with the target inside the iteration, 83 us with the next
        # iteration's target overlapped, see profiles/ for the chunked figure.
        ring = None
        C = self.target_chunk
        if chunked is not None and self.use_graph and self.fused_glue and C >= 4 and key is not None:
            n_b = self.sample_resolution**2//self.fit_world
            ring = self._rings.get(key)
            if ring is None:
                mk = lambda *shape: torch.empty(*shape, device=self.dev)  # noqa: E731
                ring = self._rings[key] = dict(X=mk(2*C, n_b, self.dim), T=mk(2*C, n_b, self.dim), S=None, x=mk(n_b, self.dim), t=mk(n_b, self.dim), s=None,
                                               ids=torch.arange(1 << 16, dtype=torch.int64, device=self.dev), gen=torch.cuda.Stream(device=self.dev),
                                               done=[torch.cuda.Event(), torch.cuda.Event()], used=[torch.cuda.Event(), torch.cuda.Event()])
            main = torch.cuda.current_stream()
            ring["gen"].wait_stream(main)   # the previous network, the epoch counter and (projection) the pressure samples are in place

            def generate(c):
                half = c % 2
                with torch.cuda.stream(ring["gen"]):
                    if c >= 2:
                        ring["gen"].wait_event(ring["used"][half])   # the iterations of chunk c - 2 have read this half
                    X, T, S = chunked(ring["ids"][c % ring["ids"].numel()], C*n_b)
                    ring["X"][half*C:(half + 1)*C].view(-1, self.dim).copy_(X)
                    ring["T"][half*C:(half + 1)*C].view(-1, self.dim).copy_(T)
                    if S is not None:
                        if ring["S"] is None:
                            ring["S"] = torch.empty_like(ring["X"]); ring["s"] = torch.empty_like(ring["x"])
                        ring["S"][half*C:(half + 1)*C].view(-1, self.dim).copy_(S)
                    ring["done"][half].record(ring["gen"])
                if os.environ.get("NMC_CHUNK_SERIAL", "0") == "1":   # experiment: no overlap of generation and training
                    torch.cuda.current_stream().wait_event(ring["done"][half])
            generate(0)
            main.wait_event(ring["done"][0])
            if n_iters > C:
                generate(1)
            chunk_now = 0
        graph_u = None
        if cached is not None:
            graph, loss_buf, graph_u = cached
        else:
            loss_buf = fit.loss  # mean squared error, written by the iteration itself

            side2 = torch.cuda.Stream(device=self.dev) if self.overlap_targets else None

            def one_chunked(fetch=True, fetch_next=False):
                # fetch_next (inside the unrolled graph): the Adam launch of this iteration also copies the ring slot of the next one
                if fetch:
                    fit_fetch(ring["X"], ring["T"], ring["S"], fit.opt.step_dev, ring["x"], ring["t"], ring["s"])
                y = fit.forward(ring["x"])
                fit.finish(ring["x"], y, ring["t"], ring["s"],
                           fetch_next=(ring["X"], ring["T"], ring["S"], ring["x"], ring["t"], ring["s"]) if fetch_next else None)

            def one():
                # `iteration()` returns the samples and a function that computes the fit target from them.  The target
                # (one or two evaluations of the previous network) does not depend on the training forward, so the two run
                # on different streams -- inside the captured graph they become parallel branches.
                samples, make_target = iteration()
                if side2 is None:
                    target = make_target(samples)
                    fit.iterate(samples, *target) if isinstance(target, tuple) else fit.iterate(samples, target)
                    return
                main = torch.cuda.current_stream()
                side2.wait_stream(main)
                with torch.cuda.stream(side2):
                    target = make_target(samples)
                y = fit.forward(samples)
                main.wait_stream(side2)
                fit.finish(samples, y, *target) if isinstance(target, tuple) else fit.finish(samples, y, target)

            if ring is not None:
                one = one_chunked
            graph = None
            if self.use_graph:
                side = torch.cuda.Stream(device=self.dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):  # warm-up outside capture
                        one()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    one()
                it = 3
                U = self.graph_unroll
                if ring is not None and U > 1 and C % U == 0 and self.check_every % U == 0:
                    graph_u = torch.cuda.CUDAGraph()
                    merge = self.merge_adam_fetch and fit.world == 1
                    with torch.cuda.graph(graph_u):
                        for u in range(U):
                            if merge:
                                one_chunked(fetch=(u == 0), fetch_next=(u < U - 1))
                            else:
                                one()
                if key is not None:
                    self._graphs[key] = (graph, loss_buf, graph_u)
        while it < n_iters:
            if ring is not None and it//C != chunk_now:   # first iteration of the next chunk: wait for it, start the one after
                ring["used"][chunk_now % 2].record(torch.cuda.current_stream())
                chunk_now = it//C
                torch.cuda.current_stream().wait_event(ring["done"][chunk_now % 2])
                if (chunk_now + 1)*C < n_iters:
                    generate(chunk_now + 1)
            if graph_u is not None and it >= self.graph_unroll and it % self.graph_unroll == 0 and it + self.graph_unroll <= n_iters:
                graph_u.replay()   # neither a chunk boundary nor an early-stop test falls inside: both are multiples of the unroll
                it += self.graph_unroll
            elif graph is not None:
                graph.replay()
                it += 1
            else:
                one()
                it += 1
            # the reference tests the loss after every iteration (base.py:148); here after the first one (a fit whose target is
            # the network itself -- the projection with a zero pressure gradient -- stops at once, as it does there) and then
            # every `check_every` iterations, each test being a host synchronisation
            if self.early_stop and (it == 1 or it % self.check_every == 0) and self._stop_now(loss_buf):
                break
        if ring is not None:  # a chunk generated ahead of an early stop must not overlap the next fit's set-up
            torch.cuda.current_stream().wait_stream(ring["gen"])
        return it, loss_buf

    def _stop_now(self, loss_buf):
        """Early-stop test of _training_loop (base.py:148).  Data-parallel fits: `loss_buf` is this rank's shard MSE, so
        the decision is taken on the mean over ranks -- every rank issues this all_reduce at the same iteration and
        reaches the same verdict (a rank that stopped alone would leave the others waiting in the gradient all_reduce)."""
        if self._fit is not None and self._fit.stop_threshold > 0 and self.fit_world == 1:
            return bool(self._fit.opt.stop_flag.item() != 0) or loss_buf.item() <= self._fit.stop_threshold
        return collective_stop(loss_buf, self.fit_world)

    def close(self):
        """Drops the captured CUDA graphs (they hold the NCCL gradient all_reduce of the data-parallel fits: while they
        are alive torch.distributed.destroy_process_group() blocks), the fit buffers and the scene."""
        torch.cuda.synchronize(self.dev)
        for key in list(self._graphs):
            graph, _, graph_u = self._graphs.pop(key)
            graph.reset()
            if graph_u is not None:
                graph_u.reset()
        self._graphs = {}
        self._rings = {}
        self._proj = None
        if self._fit is not None:
            self._fit.close()
            self._fit = None
        if self.scene is not None:
            self.scene.handle.close()
            self.scene = None
        torch.cuda.synchronize(self.dev)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def advect_velocity(self, n_iters=None):
        n = self.sample_resolution**2//self.fit_world
        s = self.size

        def make_target(samples):
            with torch.no_grad():
                prev_u = self.query_velocity(samples, use_prev=True)
                back = fields.backtrace(samples, prev_u, self.dt, self.size[0::2], self.size[1::2])
                return self.query_velocity(back, use_prev=True)

        def iteration():
            return self.sample_random(n, fused=self.fused_glue), make_target

        def chunked(chunk_id, m):
            seed = self.seed*0x9E3779B1 + 0x632BE5AB*(self.rank if self.fit_world > 1 else 0)
            x = fit_sample_uniform(m, self.size[0::2], self.size[1::2], chunk_id, self._epoch, seed,
                                   obstacle=self.obstacle if self.boundary == "karman" else None)
            return x, make_target(x), None
        return self._loop(iteration, self.max_n_iters if n_iters is None else n_iters, key="advect", chunked=chunked)

    def divergence_grid(self):
        """-div u_prev on the grid with boundary samples, as the source array the scene expects: 2D [rows(y)][cols(x)]
        (model_split.py:230-243), 3D [x][y][z] (3d model_split.py:232-235)."""
        x = self.grid_samples.detach().clone().requires_grad_(True)
        u = self.query_velocity(x, use_prev=True)
        div = 0.0
        for i in range(self.dim):
            div = div + torch.autograd.grad(u[:, i], x, torch.ones_like(u[:, i]), retain_graph=(i < self.dim - 1))[0][:, i]
        return (-div).reshape(self.grid_shape).contiguous()

    def pressure_solve(self, samples, index_offset=None):
        div = self.divergence_grid()
        stream = torch.cuda.current_stream().cuda_stream  # the grid's producer, the copy and the solve share one stream
        self.scene.handle.set_source_device(div.data_ptr(), div.shape, stream=stream)
        n = samples.shape[0]
        p = torch.empty(n, device=self.dev); g = torch.empty((n, self.dim), device=self.dev)
        st = capi.SolveStats()
        self.scene.handle.solve_device(self.opts, samples.data_ptr(), n, p.data_ptr(), g.data_ptr(), index_offset=self.rank*(self.wost_resolution**2) if index_offset is None else index_offset,
                                       stream=stream, stats=st)
        self.last.update(walks=st.walks_started, wost_ms=st.kernel_ms, div=div)
        return p, g

    def _gather_pressure(self, samples, p, grad_p):
        """all_gather of every rank's (points, p, grad p); shard sizes may differ (obstacle rejection)."""
        import torch.distributed as dist
        cap = self.wost_resolution**2//self.world
        buf = torch.zeros(cap + 1, 2*self.dim + 1, device=self.dev)
        k = samples.shape[0]
        buf[:k, :self.dim] = samples; buf[:k, self.dim] = p; buf[:k, self.dim + 1:] = grad_p
        buf[cap, 0] = float(k)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf)
        parts = [o[: int(o[cap, 0].item())] for o in out]
        full = torch.cat(parts, dim=0)
        return full[:, :self.dim].contiguous(), full[:, self.dim].contiguous(), full[:, self.dim + 1:].contiguous()

    def _solve_shard_of_common_samples(self, samples_all):
        """Replicated fits: every rank holds the same pressure samples, solves its contiguous block and the blocks are
        all-gathered (equal padded blocks, one collective for p and grad p)."""
        import torch.distributed as dist
        from .sharding import shard_bounds
        N = samples_all.shape[0]
        b0, b1 = shard_bounds(N, self.rank, self.world)
        m = -(-N//self.world)
        p, g = self.pressure_solve(samples_all[b0:b1].contiguous(), index_offset=b0)
        loc = torch.zeros(m, self.dim + 1, device=self.dev)
        loc[: b1 - b0, 0] = p; loc[: b1 - b0, 1:] = g
        full = torch.empty(m*self.world, self.dim + 1, device=self.dev)
        dist.all_gather_into_tensor(full, loc)
        parts = []
        for r in range(self.world):
            a0, a1 = shard_bounds(N, r, self.world)
            parts.append(full[r*m: r*m + (a1 - a0)])
        full = torch.cat(parts, dim=0)
        return full[:, 0].contiguous(), full[:, 1:].contiguous()

    def project_velocity(self, n_iters=None):
        replicated = self.world > 1 and self.fit_world == 1
        samples_all = self.sample_random(self.wost_resolution**2//(1 if replicated else self.world), keep_shape=False).contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if replicated:
            p, grad_p = self._solve_shard_of_common_samples(samples_all)
        else:
            p, grad_p = self.pressure_solve(samples_all)
            if self.world > 1:
                samples_all, p, grad_p = self._gather_pressure(samples_all, p, grad_p)
        e1.record(); e1.synchronize()
        self.last["pressure_ms"] = e0.elapsed_time(e1)
        self.last.update(p=p, grad_p=grad_p, pressure_samples=samples_all)
        n, big = self.sample_resolution**2//self.fit_world, samples_all.shape[0]
        if self.use_graph:  # persistent inputs of the captured iteration: samples, gradients and their count
            if self._proj is None:
                cap = self.wost_resolution**2
                self._proj = (torch.zeros(cap, self.dim, device=self.dev), torch.zeros(cap, self.dim, device=self.dev), torch.zeros((), device=self.dev))
            ps, pg, pc = self._proj
            # 2D: randint(0, N - 1) excludes the last point (2d model_split.py:274); 3D: randint(0, N) (3d model_split.py:295)
            ps[:big].copy_(samples_all); pg[:big].copy_(grad_p); pc.fill_(float(big - 1 if self.dim == 2 else big))

            def iteration_fused():
                # the same draw and both gathers in one launch; the target is (u_prev(samples), grad p): the subtraction happens in the loss kernel
                seed = self.seed*0x9E3779B1 + 0x632BE5AB*(self.rank if self.fit_world > 1 else 0) + 0x1B873593
                xs, gs = fit_gather(n, ps, pg, pc, self._fit.opt.step_dev, self._epoch, seed)

                def make_target(samples):
                    with torch.no_grad():
                        return self.query_velocity(samples, use_prev=True), gs
                return xs, make_target

            def iteration():
                # uniform index in [0, big - 2] (the reference's randint(0, big - 1) excludes the last point, :274);
                # the bound is a device scalar so the captured graph serves every step's sample count
                idx = torch.clamp((torch.rand(n, device=self.dev)*pc).long(), max=cap_idx)

                def make_target(samples):
                    with torch.no_grad():
                        return self.query_velocity(samples, use_prev=True) - pg[idx]
                return ps[idx], make_target
            cap_idx = ps.shape[0] - 1
        else:
            def iteration():
                idx = torch.randint(0, big - 1 if self.dim == 2 else big, (n,), device=self.dev)  # 2D excludes the last point (:274)

                def make_target(samples):
                    with torch.no_grad():
                        return self.query_velocity(samples, use_prev=True) - grad_p[idx]
                return samples_all[idx], make_target
        chunked = None
        if self.use_graph and self.fused_glue:
            iteration = iteration_fused

            def chunked(chunk_id, m):
                seed = self.seed*0x9E3779B1 + 0x632BE5AB*(self.rank if self.fit_world > 1 else 0) + 0x1B873593
                xs, gs = fit_gather(m, ps, pg, pc, chunk_id, self._epoch, seed)
                with torch.no_grad():
                    return xs, self.query_velocity(xs, use_prev=True), gs
        return self._loop(iteration, self.max_n_iters if n_iters is None else n_iters, key="project", chunked=chunked)

    def _sync_prev(self):
        self.velocity_field_prev.load_state_dict(self.velocity_field.state_dict())

    def step(self, n_iters=None):
        """NeuralFluidSplit.step, adv_ref = 0 (model_split.py:44-62)."""
        if self.boundary == "smoke":
            self._noise_seed.fill_(self.timestep)
        self._sync_prev()
        it_a, loss_a = self.advect_velocity(n_iters)
        self._rebroadcast_weights()
        self._sync_prev()
        it_p, loss_p = self.project_velocity(n_iters)
        self._rebroadcast_weights()
        self._sync_prev()
        self.timestep += 1
        self.opts.seed = (self.seed + self.timestep) & 0xFFFFFFFFFFFFFFFF
        return {"advect_iters": it_a, "advect_loss": loss_a, "project_iters": it_p, "project_loss": loss_p}

    def _rebroadcast_weights(self):
        """Replicated fits: the ranks' weights agree up to the last bits (atomic gradient reductions) and the ranks may stop
        a fit at different iterations; continue from rank 0's."""
        if self.world > 1 and self.fit_world == 1 and self._fit is not None:
            import torch.distributed as dist
            dist.broadcast(self._fit.opt.flat, src=0)

    def karman_initial_velocity(self, samples):
        """karman_vortex_velocity (sources.py:34-43): (karman_vel, 0) times the obstacle weight."""
        w = torch.clamp(self.obstacle_distance(samples), 0, self.eps)/self.eps
        return torch.stack([self.karman_vel*w, torch.zeros_like(w)], dim=-1)

    def fit_initial(self, velocity_fn, n_iters, lr=1e-4):
        """Fit the network to an analytic initial velocity (main.py: initial condition fit)."""
        opt = torch.optim.Adam(self.velocity_field.parameters(), lr=lr)
        for _ in range(n_iters):
            x = self.sample_random(self.sample_resolution**2)
            loss = torch.mean((self.query_velocity(x) - velocity_fn(x))**2)
            opt.zero_grad(); loss.backward(); opt.step()
        if self.world > 1:  # every rank fitted on its own samples: continue from rank 0's weights
            import torch.distributed as dist
            for prm in self.velocity_field.parameters():
                dist.broadcast(prm.data, src=0)
        self._sync_prev()
        return loss.item()
