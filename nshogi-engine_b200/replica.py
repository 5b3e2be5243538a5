"""Multi-GPU host logic: independent replicas, one process per GPU, no data-path collective.

The reference scales by giving each GPU its own ``Infer`` fed from a shared queue (reference
src/mcts/manager.cc:168-179, src/selfplay/main.cc:189-195); leaf evaluations are independent, so
here self-play game sets / evaluation batches are sharded by ``unit_id mod world`` and the only
exchange is the end-of-run reduction of counters and the max-over-ranks of the timed region
(torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Sequence


@dataclass
class RankInfo:
    rank: int
    local_rank: int
    world: int

    @classmethod
    def from_env(cls) -> "RankInfo":
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
                   int(os.environ.get("WORLD_SIZE", "1")))


def shard_units(total_units: int, rank: int, world: int) -> range:
    """Units (games, batches) owned by `rank`: unit_id mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank, total_units, world)


def shard_counts(total_units: int, world: int) -> List[int]:
    return [len(shard_units(total_units, r, world)) for r in range(world)]


COUNTER_KEYS = ("evals", "batches", "legal_moves", "nan_rows", "records", "games")


def aggregate(counters: Dict[str, int], elapsed_ms: float, device=None):
    """Sum the per-rank counters and take the max of the per-rank timed region.  With a single
    process (no initialised process group) this is the identity."""
    import torch
    import torch.distributed as dist

    vec = torch.tensor([int(counters.get(k, 0)) for k in COUNTER_KEYS], dtype=torch.int64, device=device)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {k: int(v) for k, v in zip(COUNTER_KEYS, vec.tolist())}, float(t.item())


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def whole_job_rate(total_units: int, elapsed_ms_max: float) -> float:
    return total_units / (elapsed_ms_max * 1e-3) if elapsed_ms_max > 0 else 0.0
