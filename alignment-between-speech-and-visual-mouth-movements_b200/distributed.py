"""Multi-GPU plumbing for the sweep: clips are independent units, so rank r of
R takes a contiguous shard and the only exchange is one gather of scores / best offsets (and, for the
detector-training config, one all-reduce of a flat gradient bucket per step).  One process per GPU,
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank ``rank``; sizes differ by at most one; covers [0, n)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scores(scores: torch.Tensor, best: torch.Tensor, n_total: int, group=None):
    """all_gather of the per-rank [n_local, K] scores and [n_local] best offsets into the global
    [n_total, K] / [n_total] arrays (ragged shards padded to the largest shard for the collective)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scores, best
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    K = scores.shape[1]
    pad_s = scores.new_zeros((n_max, K))
    pad_b = best.new_zeros((n_max,))
    pad_s[: scores.shape[0]] = scores
    pad_b[: best.shape[0]] = best
    out_s = scores.new_empty((world * n_max, K))
    out_b = best.new_empty((world * n_max,))
    dist.all_gather_into_tensor(out_s, pad_s, group=group)
    dist.all_gather_into_tensor(out_b, pad_b, group=group)
    out_s = torch.cat([out_s[r * n_max: r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
    out_b = torch.cat([out_b[r * n_max: r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
    return out_s, out_b


def sweep_sharded(sweeper, frames: torch.Tensor, audio: torch.Tensor, n_total: Optional[int] = None, group=None):
    """Run this rank's shard through ``sweeper`` and gather everyone's results.  ``frames``/``audio``
    are this rank's LOCAL clips (shard ``shard_range(n_total, rank, world)``)."""
    scores, best = sweeper.run(frames, audio)
    if n_total is None:
        n_total = frames.shape[0] * (dist.get_world_size(group) if dist.is_initialized() else 1)
    return gather_scores(scores, best, n_total, group)


def gather_labels_probs(labels: torch.Tensor, probs: torch.Tensor, n_total: int, group=None):
    """Gather (label, prob) pairs so rank 0 can call sklearn's roc_auc_score on the full set, which
    keeps AUC parity with the reference's run_epoch (misalignment_detection_train.py:272-279)."""
    s, _ = gather_scores(torch.stack([labels, probs], dim=1), torch.zeros_like(labels, dtype=torch.int32), n_total, group)
    return s[:, 0], s[:, 1]


def auc_acc(labels: np.ndarray, probs: np.ndarray):
    """accuracy at 0.5 and ROC-AUC (NaN if a single class is present), as run_epoch reports them."""
    from sklearn.metrics import accuracy_score, roc_auc_score
    acc = accuracy_score(labels, (probs > 0.5).astype(float))
    try:
        auc = roc_auc_score(labels, probs)
    except ValueError:
        auc = float("nan")
    return acc, auc


def ddp_detector_step(model, features: torch.Tensor, labels: torch.Tensor, optimizer, criterion=None, group=None):
    """One data-parallel training step of the detector (config 5): local BCE-with-logits backward, ONE
    all-reduce of a flat gradient bucket (7 099 393 fp32 for hidden 512), identical Adam on all ranks
    (reference step: misalignment_detection_train.py:260-266).  Returns the rank-mean loss."""
    criterion = criterion or torch.nn.BCEWithLogitsLoss()
    model.train()
    logits = model(features)
    loss = criterion(logits, labels)
    optimizer.zero_grad(set_to_none=False)
    loss.backward()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= world
        off = 0
        for g in grads:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        l = loss.detach().clone()
        dist.all_reduce(l, op=dist.ReduceOp.SUM, group=group)
        loss = l / world
    optimizer.step()
    return loss.detach()
