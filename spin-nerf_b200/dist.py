"""Ray-sharded data parallelism: one process per GPU, rays split contiguously, the only
data-path exchange is one all-reduce of each flat gradient vector per step (NCCL over NVLink on GPUs; the same code
runs on gloo/CPU tensors in tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class RaySharder:
    """Contiguous ray shards per rank: rank r gets rays [r*n/W, (r+1)*n/W) of every batch (equal shards when W divides
    n, so the mean of per-rank MSE means equals the global MSE and averaging gradients over ranks is exact)."""

    def __init__(self, rank=0, world=1):
        self.rank, self.world = int(rank), int(world)

    def bounds(self, n):
        return (n * self.rank) // self.world, (n * (self.rank + 1)) // self.world

    def shard(self, t, dim=0):
        lo, hi = self.bounds(t.shape[dim])
        return t.narrow(dim, lo, hi - lo)


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def allreduce_sum_(flat_grads, group=None):
    """In-place SUM all-reduce of the flat gradient vectors; returns the factor that turns the sum into the mean
    (applied inside the fused Adam as grad_scale, so no extra pass over the gradients)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    for g in flat_grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def allreduce_sum_async(flat, group=None):
    """In-place SUM all-reduce issued asynchronously: the collective waits for the work already queued on the current stream
    (the backward pass that produced `flat`) and runs on the backend's own stream while later launches of the current stream
    compute.  Returns the work handle (`.wait()` orders the current stream behind the collective) or None without a group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)


def broadcast_parameters(nets, src=0, group=None):
    """Replicas must start identical ("initial weight broadcast (or identical seeds)"): one broadcast of
    each network's flat parameter vector from rank `src`.  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for net in nets:
        if net is None:
            continue
        dist.broadcast(net.flat_params(), src=src, group=group)
        net.mark_params_changed()


def frames_for_rank(n_frames, rank, world):
    """render_path sharding (config 5): frames round-robin over ranks."""
    return list(range(rank, n_frames, world))


def gather_frames(local, n_frames, rank, world, group=None, dst=None):
    """local: this rank's frames [ceil(n_frames / world), ...] in its round-robin order (frame rank + j*world at row j,
    unused tail rows arbitrary) -> all n_frames in frame order on every rank (dst=None) or on rank dst only (others get
    None).  One all-gather / gather of equally sized blocks; the interleave back to frame order is a view transpose."""
    per = (n_frames + world - 1) // world
    if local.shape[0] != per:
        raise RuntimeError(f"gather_frames: expected {per} local frame rows, got {local.shape[0]}")
    if world == 1:
        return local[:n_frames]
    if dst is None:
        full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)       # rank-major concatenation
        full = full.view((world,) + tuple(local.shape))
    else:
        parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local.contiguous(), parts, dst=dst, group=group)
        if rank != dst:
            return None
        full = torch.stack(parts, 0)
    # full[r, j] is frame r + j*world  ->  [j, r] flattened is frame order
    return full.transpose(0, 1).reshape((per * world,) + tuple(local.shape[1:]))[:n_frames]
