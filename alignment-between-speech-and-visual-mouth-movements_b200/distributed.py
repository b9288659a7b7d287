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


class GraphedDetectorStep:
    """``ddp_detector_step`` captured in ONE CUDA graph (forward, BCE-with-logits, backward, the flat-bucket all-reduce,
    Adam): the step of the small detector MLP is ~30 launch-bound torch kernels, and a replay costs one launch.

    The optimizer is switched to its capturable form (step counters on the device), a few warm-up steps run on a side
    stream — the caches cuBLAS / NCCL / the optimizer build must exist before capture — and the parameters and the
    optimizer state are put back IN PLACE afterwards, so that construction does not train the model.  ``step(x, y)``
    copies the batch into the static input tensors and replays the graph; it returns the rank-mean loss (a static
    tensor, overwritten by the next step).  Batch size and feature width are fixed at construction.

    Same arithmetic as the eager step except that the capturable Adam forms its bias corrections on the device in fp32
    (eager: on the host in double): after five steps of size lr = 1e-3 the parameters differ by 2.6e-6, not bit for bit."""

    def __init__(self, model, optimizer, batch_size: int, criterion=None, group=None, warmup: int = 3):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedDetectorStep needs the model on a CUDA device")
        self.model, self.optimizer, self.group = model, optimizer, group
        self.criterion = criterion or torch.nn.BCEWithLogitsLoss()
        self.x = torch.zeros((batch_size, model.classifier[0].in_features), device=dev)
        self.y = torch.zeros((batch_size,), device=dev)
        for g in optimizer.param_groups:
            g["capturable"] = True
        for st in optimizer.state.values():            # an optimizer that has already stepped eagerly
            if "step" in st and not (torch.is_tensor(st["step"]) and st["step"].is_cuda):
                st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=dev)
        params = [p for p in model.parameters()]
        p_saved = [p.detach().clone() for p in params]
        had_state = {p: {k: v.detach().clone() for k, v in optimizer.state[p].items() if torch.is_tensor(v)}
                     for p in params if p in optimizer.state and optimizer.state[p]}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                ddp_detector_step(model, self.x, self.y, optimizer, self.criterion, group)
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():                           # undo the warm-up, keeping every tensor where it is
            for p, s in zip(params, p_saved):
                p.copy_(s)
            for p in params:
                for k, v in optimizer.state[p].items():
                    if torch.is_tensor(v):
                        v.copy_(had_state[p][k]) if p in had_state else v.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = ddp_detector_step(model, self.x, self.y, optimizer, self.criterion, group)

    def step(self, features: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.x.copy_(features, non_blocking=True)
        self.y.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.loss
