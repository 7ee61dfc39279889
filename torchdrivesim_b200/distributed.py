"""Multi-GPU plumbing: one process per GPU, environments sharded along the batch dimension.

The hot path has no data-path exchange: environments never interact (collisions are within an
environment), so each rank owns a contiguous block of environments and the per-map grids are replicated.
The only collective is a SUM all-reduce of a small vector of aggregate infraction metrics
(`allreduce_metrics`), issued with torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: Optional[str] = None, device: Optional[torch.device] = None) -> None:
    rank, world, _ = rank_world()
    if world <= 1 or dist.is_initialized():
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kwargs = {}
    if backend == "nccl" and device is not None:
        kwargs["device_id"] = device
    dist.init_process_group(backend, **kwargs)


def shard_range(n_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [b0, b1) of environments owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_envs, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_indices(n_envs: int, rank: int, world: int, device=None) -> torch.Tensor:
    b0, b1 = shard_range(n_envs, rank, world)
    return torch.arange(b0, b1, device=device)


def infraction_metrics(collision: torch.Tensor, offroad: torch.Tensor, present: torch.Tensor) -> torch.Tensor:
    """[6] float64: collision sum, offroad sum, colliding agents, offroad agents, present agents, agent slots - the
    definition of the metric vector as a torch expression (host-side logic, any device; the gloo tests use it).
    The hot path computes and accumulates the same vector with ONE launch: `ops.infraction_metrics`
    (tds_infraction_metrics, CUDA only)."""
    p = present.to(collision.dtype)
    return torch.stack([(collision * p).sum(), (offroad * p).sum(), ((collision > 0) & present).sum(),
                        ((offroad > 0) & present).sum(), present.sum(), torch.tensor(present.numel(), device=present.device)]
                       ).to(torch.float64)


METRIC_NAMES = ("collision_sum", "offroad_sum", "colliding_agents", "offroad_agents", "present_agents", "agent_slots")


def allreduce_metrics(local: torch.Tensor, async_op: bool = False):
    """SUM over ranks of the metric vector (in place).  Returns the work handle when async_op."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(local, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def metrics_dict(v: torch.Tensor) -> Dict[str, float]:
    return {k: float(x) for k, x in zip(METRIC_NAMES, v.tolist())}
