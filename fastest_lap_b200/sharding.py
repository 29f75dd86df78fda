"""Sharding of batches of independent problems over the GPUs of one box.

The reference solves one lap (or one steady-state NLP) at a time on one CPU thread; batches (G-G sweeps,
`src/core/applications/steady_state.hpp:682-928`; vehicle-parameter sweeps) are loops.  Independent problems need no
communication while they are evaluated or solved: problem p of B goes to rank floor(p * G / B) (contiguous blocks), every
rank works on its own block on its own GPU, and the only collective is one final gather of the per-problem results.
One process per GPU; the process group is torch.distributed's (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(count, world):
    """Start offsets [world + 1] of the contiguous blocks: problem p belongs to rank floor(p * world / count)."""
    if count < 0 or world < 1:
        raise ValueError("need count >= 0 and world >= 1")
    # smallest p with floor(p * world / count) >= r  is  ceil(r * count / world)
    return [-((-r * count) // world) for r in range(world + 1)]


def shard_range(count, world, rank):
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    b = shard_bounds(count, world)
    return b[rank], b[rank + 1]


def owner_of(problem, count, world):
    return (problem * world) // count


def gather_results(local, count, group=None, device=None):
    """Final gather of per-problem results: `local` is this rank's [n_local, ...] block (numpy); every rank returns the
    full [count, ...] array in problem order.  One all_gather of equal-size padded blocks (ncclAllGather on the GPUs)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b = shard_bounds(count, world)
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.shape[0] != b[rank + 1] - b[rank]:
        raise ValueError("rank %d holds %d problems, its shard has %d" % (rank, local.shape[0], b[rank + 1] - b[rank]))
    tail = local.shape[1:]
    width = max(b[r + 1] - b[r] for r in range(world))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pad = torch.zeros((width,) + tail, dtype=torch.float64, device=device)
    pad[:local.shape[0]] = torch.from_numpy(local).to(device)
    out = torch.empty((world, width) + tail, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    out = out.cpu().numpy()
    return np.concatenate([out[r, :b[r + 1] - b[r]] for r in range(world)], axis=0)
