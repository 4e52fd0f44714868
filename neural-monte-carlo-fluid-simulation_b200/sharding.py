"""Multi-GPU sharding of the query points.

Every query point is an independent unit of work (walk_on_stars.h:91-95), so the path shards with no
data-path collective: one process per GPU, the scene (boundary structure + source grid) replicated,
points split into contiguous blocks, and each point's RNG keyed by its GLOBAL index so results do not
depend on the number of ranks.  The only exchange is the final gather of N x (1 + dim) floats, done with
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one point."""
    base, rem = divmod(int(n), int(world))
    lo = rank*base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_estimates(p_local, g_local, n, dim, group=None):
    """all_gather of the per-rank estimates into full arrays (tensors on the caller's device).
    Works with unequal shard sizes by padding to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return p_local, g_local
    sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    m = max(sizes)
    buf = torch.zeros((m, 1 + dim), dtype=torch.float32, device=p_local.device)
    k = sizes[rank]
    buf[:k, 0] = p_local.reshape(-1)
    buf[:k, 1:] = g_local.reshape(-1, dim)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
    return full[:, 0].contiguous(), full[:, 1:].contiguous()


def solve_sharded(solve_fn, pts, group=None):
    """Runs `solve_fn(pts_block, index_offset) -> (p, g)` (numpy) on this rank's block and gathers.
    Used by the gloo tests with a CPU stand-in for solve_fn and by bench.py with the CUDA path."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n, dim = pts.shape
    lo, hi = shard_bounds(n, rank, world)
    p, g = solve_fn(pts[lo:hi], lo)
    if world == 1:
        return np.asarray(p), np.asarray(g)
    pt, gt = gather_estimates(torch.as_tensor(np.asarray(p)), torch.as_tensor(np.asarray(g)), n, dim, group)
    return pt.numpy(), gt.numpy()
