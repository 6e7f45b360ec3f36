"""Multi-GPU sharding of the scoring path: host-side logic only (no kernels here).

The path shards two ways (DESIGN.md "Multi-GPU"):

* **scene batch** (BASELINE.json configs[3]): independent scenes are block-partitioned over ranks, every
  rank scores its own scenes with ``sfw_score_batch`` and the 32-byte ``SfwBest`` records are all-gathered
  so every rank holds every scene's command.  No floats are reduced across ranks, so the result is
  independent of the rank count.
* **row slabs of one scene** (configs[1]/[4] on several GPUs): the scene is replicated, rank r scores
  linvel rows ``[rows(r), rows(r+1))`` (``sfw_set_row_slab``) and the per-rank winners are merged with the
  reference's own tie-break order (src/sfw_planner.cpp:394-414): lower cost, then higher linvel, then lower
  ``|angvel|``, then later index.

The collective is ``torch.distributed.all_gather_into_tensor`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from ._abi import BEST_DTYPE


def block_partition(n: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of ``n`` units for ``rank``: contiguous blocks, the first ``n % world`` one longer."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    q, r = divmod(n, world)
    b = rank * q + min(rank, r)
    return b, b + q + (1 if rank < r else 0)


def merge_winners(records: np.ndarray) -> np.ndarray:
    """Merge per-slab winners of ONE scene (``records``: BEST_DTYPE[k]) into the scene's winner.

    Total order of the reference's sequential best-update (src/sfw_planner.cpp:394-414):
    cost ascending, linvel descending, |angvel| ascending, index descending (a later equal sample wins).
    """
    records = np.asarray(records, dtype=BEST_DTYPE).reshape(-1)
    best = np.zeros((), dtype=BEST_DTYPE)
    for r in records:
        if not r["valid"]:
            continue
        if not best["valid"]:
            best = r.copy()
            continue
        ka = (float(r["cost"]), -float(r["v"]), abs(float(r["w"])), -int(r["index"]))
        kb = (float(best["cost"]), -float(best["v"]), abs(float(best["w"])), -int(best["index"]))
        if ka < kb:
            best = r.copy()
    return best


def all_gather_best(best: np.ndarray, device=None):
    """All-gather ``SfwBest`` records (BEST_DTYPE[n_local], same n_local on every rank).

    Returns BEST_DTYPE[world, n_local] on every rank.  Uses the default process group: NCCL when ``device``
    is a CUDA device (the record bytes are staged through a device tensor), gloo on CPU.
    """
    import torch
    import torch.distributed as dist

    best = np.ascontiguousarray(best, dtype=BEST_DTYPE).reshape(-1)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return best.reshape(1, -1).copy()
    mine = torch.from_numpy(best.view(np.uint8).copy())
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().view(BEST_DTYPE).reshape(world, -1)


def score_sharded_scenes(scorer, params, scenes, linvels, angvels, sfm=None, device=None):
    """Scene-batch sharding: score this rank's block of ``scenes`` and all-gather the winners.

    Every rank passes the same full scene list (or at least its own block); returns BEST_DTYPE[n_scenes]
    in scene order.  Requires ``n_scenes % world == 0`` (fixed-size all-gather)."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = len(scenes)
    if n % world:
        raise ValueError("score_sharded_scenes needs n_scenes divisible by the world size")
    b, e = block_partition(n, world, rank)
    _, best = scorer.score(params, scenes[b:e], linvels, angvels, sfm=sfm, want_costs=False)
    return all_gather_best(best, device=device).reshape(-1)


def score_sharded_rows(scorer, params, scene, linvels, angvels, sfm=None, device=None):
    """Row-slab sharding of one scene: returns the merged winner (BEST_DTYPE scalar) on every rank."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    b, e = block_partition(len(linvels), world, rank)
    scorer.upload(params, [scene], linvels, angvels, sfm=sfm)
    scorer.set_row_slab(b, e)
    scorer.run()
    _, best = scorer.download(want_costs=False)
    return merge_winners(all_gather_best(best, device=device).reshape(-1))
