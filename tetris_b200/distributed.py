"""Multi-GPU plumbing: one process per GPU, envs sharded contiguously, no per-step communication.

Envs are independent, and every env's RNG is keyed by (seed, GLOBAL env id), so a shard of a
bigger job is just a BatchedTetris with `env_offset` set: results do not depend on the sharding.  The only
exchange is the end-of-rollout reduction of the episode statistics -- an int64[16] vector, sums except two
maxima -- done as one all-gather with torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests) plus a
local combine.
"""
import os

import torch

from . import _lib


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process when unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total_envs, rank, world_size):
    """Contiguous env range [offset, offset + count) of `rank`: the first total % world ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_shard(num_columns, num_rows, total_envs, rank=None, world_size=None, **kw):
    """This rank's shard of a `total_envs`-env job as a BatchedTetris (global env ids preserved)."""
    from .batched import BatchedTetris
    r, w, _ = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    offset, count = shard_range(total_envs, rank, world_size)
    return BatchedTetris(num_columns, num_rows, count, env_offset=offset, **kw)


def combine_stats(parts):
    """Combine per-shard statistics vectors on one process (sum, max for the two maxima)."""
    parts = [torch.as_tensor(p, dtype=torch.int64) for p in parts]
    out = torch.stack(parts).sum(0)
    for i in _lib.STATS_MAX_FIELDS:
        out[i] = torch.stack([p[i] for p in parts]).max()
    return out


_gather_buf = {}


def reduce_stats(stats, group=None, out=None):
    """Reduce an episode-statistics vector over the process group: SUM everywhere except MAX for max_ep_lines /
    max_ep_steps.  ONE collective -- an all-gather of the ranks' int64[16] vectors into a preallocated buffer --
    followed by the local combine (tb_combine_stats, one tiny kernel on the same stream; CPU tensors, as in the gloo
    tests, are combined on the host).  Integer arithmetic, so the result is exact and independent of the order.
    Returns `out` (a new tensor when not given); a plain copy when torch.distributed is not initialised."""
    import ctypes as C
    import torch.distributed as dist
    if out is None:
        out = torch.empty_like(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out.copy_(stats)
        return out
    world_size = dist.get_world_size(group)
    key = (stats.device, world_size, stats.numel())
    buf = _gather_buf.get(key)
    if buf is None:
        buf = _gather_buf[key] = torch.empty(world_size * stats.numel(), dtype=torch.int64, device=stats.device)
    dist.all_gather_into_tensor(buf, stats.contiguous(), group=group)          # rank r's vector lands at [16 r, 16 r + 16)
    if stats.is_cuda:
        stream = torch.cuda.current_stream(stats.device).cuda_stream
        _lib.check(_lib.lib().tb_combine_stats(C.c_void_p(buf.data_ptr()), world_size, C.c_void_p(out.data_ptr()),
                                               C.c_void_p(stream)))
    else:
        out.copy_(combine_stats(list(buf.view(world_size, -1))))
    return out
