"""Multi-GPU layout: environments shard by index, one process per GPU, no collective on the step
path.  The only exchange is the per-iteration rollout statistics -- the
numbers RLlib's callbacks aggregate in the reference (paint_ppo.py:36-72) -- reduced with one
small `all_reduce` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist

STAT_KEYS = ('env_steps', 'episodes', 'sum_reward', 'sum_penalty', 'sum_return', 'new_texels',
             'max_episode_len', 'max_step_ms')
_MAX_KEYS = ('max_episode_len', 'max_step_ms')


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def shard_range(total_envs, rank, world_size):
    """Contiguous block of environment indices owned by `rank`: env e lives on GPU
    e // ceil(total / world) ; the last rank may own fewer."""
    if total_envs < 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError('bad shard request')
    per = -(-total_envs // world_size)
    lo = min(rank * per, total_envs)
    hi = min(lo + per, total_envs)
    return lo, hi


def init_process_group(backend=None):
    """Join the default group if torchrun started us; returns (rank, local_rank, world_size)."""
    rank, local_rank, world_size = world()
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size)
    return rank, local_rank, world_size


def allreduce_stats(stats, device=None):
    """Reduce a dict of per-rank rollout statistics over all ranks: sums, except the `max_*`
    keys which take the maximum.  Returns plain Python floats; identity when not distributed."""
    vals = torch.tensor([float(stats.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        is_max = torch.tensor([k in _MAX_KEYS for k in STAT_KEYS], device=vals.device)
        sums = torch.where(is_max, torch.zeros_like(vals), vals)
        maxs = torch.where(is_max, vals, torch.full_like(vals, float('-inf')))
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
        vals = torch.where(is_max, maxs, sums)
    return {k: float(v) for k, v in zip(STAT_KEYS, vals.cpu())}
