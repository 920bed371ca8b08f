"""Multi-GPU sharding of the render job: by building, no collective on the data path.

Mirrors the reference's process-level parallelism over buildings (scripts/render_dataset_bev.py:186-191)
with one process per GPU.  Only scalars (unit counts, elapsed time) ever cross ranks.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple


def assign_buildings(hyp_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of buildings to ranks by hypothesis count.
    Deterministic: every rank computes the same partition with no communication."""
    order = sorted(range(len(hyp_counts)), key=lambda b: (-hyp_counts[b], b))
    parts: List[List[int]] = [[] for _ in range(world_size)]
    loads = [0] * world_size
    for b in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(b)
        loads[r] += hyp_counts[b]
    return parts


def gather_totals(local_units: int, elapsed_s: float) -> Tuple[dict, List[int]]:
    """All ranks learn the job total and the slowest rank's time (the only collective in the job)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(units=int(local_units), max_elapsed_s=float(elapsed_s)), [int(local_units)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(local_units), float(elapsed_s)], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, t)
    units = [int(g[0].item()) for g in gathered]
    return dict(units=sum(units), max_elapsed_s=max(float(g[1].item()) for g in gathered)), units
