"""Multi-GPU sharding of guided sampling: one process per GPU, zero per-step traffic, ONE all-gather at the end.

(object, candidate) pairs are independent in ``guided_sample`` (generator/diffusion.py:561-570: every object
restarts from the same noise; no BatchNorm batch statistics in eval mode), so per-object mode shards OBJECTS
contiguously across ranks and nothing crosses NVLink until the final scores / designs are gathered for the
best-of-N table.  Multi-object mode averages gradients over objects per candidate
(:640-644), so it shards CANDIDATES and keeps every object on every rank -- still no per-step traffic.

The reference's own multi-GPU story is ``nn.DataParallel`` row scatter per cond_fn call (generator/train.py:86)
plus duplicated DDP validation; it is not a model to follow.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi): the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def all_gather_ragged(t: torch.Tensor, sizes: List[int], group=None) -> torch.Tensor:
    """All-gather along dim 0 of per-rank tensors whose dim-0 sizes are ``sizes`` (rank-major result).
    Pads to the largest shard so a single ``all_gather_into_tensor`` (NCCL over NVLink, or gloo) suffices."""
    world = dist.get_world_size(group)
    mx = max(sizes)
    pad = t
    if t.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.reshape((world, mx) + tuple(t.shape[1:]))
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def gather_per_object_results(local: Dict[str, torch.Tensor], n_obj_global: int, group=None,
                              gather_designs: bool = True) -> Dict[str, torch.Tensor]:
    """``local`` is ``Diffusion.guided_sample`` output for this rank's object shard.  Returns the global tables
    (rank-major == global object order, since shards are contiguous): scores (n_obj,B), best_ids (n_obj,k),
    best_scores, [designs (n_obj,B,P,1)].  best_ids index candidates within an object, so they need no offset."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    sizes = shard_sizes(n_obj_global, dist.get_world_size(group))
    out = {k: all_gather_ragged(local[k], sizes, group) for k in ("scores", "best_ids", "best_scores")}
    if gather_designs:
        out["designs"] = all_gather_ragged(local["designs"], sizes, group)
    return out


def gather_multi_object_results(local: Dict[str, torch.Tensor], batch_global: int, top_k: int, select_fn,
                                group=None) -> Dict[str, torch.Tensor]:
    """Candidate-sharded multi-object mode: gather (B,) scores and designs, then every rank runs the same
    best-of-N (lowest global index wins ties) on the full score vector."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    sizes = shard_sizes(batch_global, dist.get_world_size(group))
    scores = all_gather_ragged(local["scores"], sizes, group)
    designs = all_gather_ragged(local["designs"], sizes, group)
    idx, best = select_fn(scores[None], top_k)
    return {"scores": scores, "designs": designs, "best_ids": idx[0], "best_scores": best[0]}
