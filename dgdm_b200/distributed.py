"""Multi-GPU sharding of guided sampling: one process per GPU, zero per-step traffic, ONE all-gather at the end.

(object, candidate) pairs are independent in ``guided_sample`` (generator/diffusion.py:561-570: every object
restarts from the same noise; no BatchNorm batch statistics in eval mode), so per-object mode shards OBJECTS
contiguously across ranks when there are at least as many objects as ranks, and CANDIDATES otherwise (the stock 3D
set is 5 objects, assets/object_names_test.txt: on 8 GPUs an object split would idle three of them); nothing crosses
NVLink until the final scores / designs are gathered for the best-of-N table.  Multi-object mode
averages gradients over objects per candidate (:640-644), so it always shards CANDIDATES and keeps every object on
every rank -- still no per-step traffic.

Bit-exactness across world sizes: a design's trajectory does not depend on its neighbours in the batch, and the
guidance kernel sums a (design, object) pair's pose rows in 128-row tiles that are aligned to the START OF THE
BATCH.  An object shard therefore reproduces the 1-rank result bit for bit whenever candidates x pose rows is a
multiple of 128 (every BASELINE configuration: 256 x 900, 128 x 1125, 512 x 1125); a candidate shard changes the tile
alignment of a pair and agrees to fp32 reassociation (~1e-6 relative) instead.

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


def plan_per_object(n_obj: int, world: int) -> str:
    """'objects' when every rank can own at least one whole object, else 'candidates' (SURVEY.md §8e)."""
    return "objects" if n_obj >= world else "candidates"


def gather_per_object_candidate_shards(local: Dict[str, torch.Tensor], batch_global: int, top_k: int, select_fn,
                                       group=None, gather_designs: bool = True) -> Dict[str, torch.Tensor]:
    """Per-object mode with CANDIDATES sharded (fewer objects than ranks): ``local`` holds every object with this
    rank's candidate slice -- scores (n_obj, B_local), designs (n_obj, B_local, P, 1).  Gathers along the candidate axis
    (rank-major == global candidate order) and re-runs the selection on the full (n_obj, B) table on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    sizes = shard_sizes(batch_global, dist.get_world_size(group))
    scores = all_gather_ragged(local["scores"].transpose(0, 1).contiguous(), sizes, group).transpose(0, 1).contiguous()
    idx, best = select_fn(scores, top_k)
    out = {"scores": scores, "best_ids": idx, "best_scores": best}
    if gather_designs:
        out["designs"] = all_gather_ragged(local["designs"].transpose(0, 1).contiguous(), sizes, group).transpose(0, 1).contiguous()
    return out


def sharded_guided_sample(build_dm, objects: torch.Tensor, fps_starts: Optional[torch.Tensor], noise: torch.Tensor,
                          opt_obj: str, top_k: int = 1, gather_designs: bool = True, group=None, dm=None,
                          **sample_kwargs) -> Dict[str, torch.Tensor]:
    """Per-object guided sampling of a GLOBAL object set on every rank of ``group``; every rank returns the same global
    tables (scores (n_obj,B), best_ids (n_obj,k), best_scores [, designs (n_obj,B,P,1)]).

    ``build_dm(objects_shard, fps_shard, object_ids) -> Diffusion`` builds this rank's sampler (pass ``dm`` to reuse
    one built for the same shard).  Objects are sharded when ``n_obj >= world``, candidates otherwise."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_obj, B = objects.shape[0], noise.shape[0]
    if plan_per_object(n_obj, world) == "objects":
        lo, hi = shard_range(n_obj, world, rank)
        if dm is None:
            dm = build_dm(objects[lo:hi], None if fps_starts is None else fps_starts[lo:hi], list(range(lo, hi)))
        local = dm.guided_sample(0, B, noise, opt_obj=opt_obj, top_k=top_k, **sample_kwargs)
        return gather_per_object_results(local, n_obj, group, gather_designs=gather_designs)
    lo, hi = shard_range(B, world, rank)
    if dm is None:
        dm = build_dm(objects, fps_starts, list(range(n_obj)))
    if hi == lo:                                                        # more ranks than candidates: nothing to do here
        P = noise.shape[1]
        local = {"scores": torch.empty((n_obj, 0), device=dm.device), "designs": torch.empty((n_obj, 0, P, 1), device=dm.device)}
    else:
        local = dm.guided_sample(0, hi - lo, noise[lo:hi], opt_obj=opt_obj, top_k=min(top_k, hi - lo), **sample_kwargs)
    return gather_per_object_candidate_shards(local, B, top_k, dm.best_of_n, group, gather_designs=gather_designs)


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
