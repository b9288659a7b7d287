"""`utils.decode_prediction` drop-in (reference: utils.py:8-36) + the batched decode it is built on."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _native as N


def ctc_greedy_decode(logp: torch.Tensor, blank_index: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched greedy CTC decode (K5).  logp [B,T,V] CUDA f32 -> (ids int32 [B,T] padded with -1,
    lengths int32 [B]).  One launch and no host sync, instead of the reference's per-sample
    ``.cpu().numpy()`` + Python loop (utils.py:20-30)."""
    N.require_cuda(logp, "logp")
    if logp.dim() != 3:
        raise ValueError("logp must be [B, T, V]")
    logp = N.f32c(logp)
    B, T, V = logp.shape
    ids = torch.empty((B, T), dtype=torch.int32, device=logp.device)
    lens = torch.empty((B,), dtype=torch.int32, device=logp.device)
    if B:
        N.check(N.lib().avs_ctc_greedy(N.ptr(logp), B, T, V, int(blank_index), N.ptr(ids), N.ptr(lens),
                                       N.stream_ptr()), "ctc_greedy")
    return ids, lens


def ids_to_text(ids: List[int], dataset) -> str:
    """utils.py:33-34 — id -> char through ``dataset.idx_to_char`` (id 38 renders as '<pad>')."""
    table = dataset.idx_to_char
    return "".join(table.get(i, "") for i in ids if i in table)


def decode_prediction(outputs: torch.Tensor, dataset, blank_index: int = 0) -> str:
    """Same contract as the reference: ``outputs`` [T, V] log-probs of ONE sample -> text."""
    ids, lens = ctc_greedy_decode(outputs.unsqueeze(0), blank_index)
    n = int(lens[0].item())
    return ids_to_text(ids[0, :n].tolist(), dataset)


def decode_batch(outputs: torch.Tensor, dataset, blank_index: int = 0) -> List[str]:
    """Decode a whole batch [B,T,V] with one kernel launch and one device->host copy."""
    ids, lens = ctc_greedy_decode(outputs, blank_index)
    ids_h, lens_h = ids.cpu().tolist(), lens.cpu().tolist()
    return [ids_to_text(row[:n], dataset) for row, n in zip(ids_h, lens_h)]


def evaluate_model(model, test_loader, dataset, device, num_samples=5, verbose=True):
    """Reference utils.py:38-86 — forward, decode, naive positional character accuracy for the first
    ``num_samples`` items.  One batched decode (K5) per batch instead of one device sync per sample.
    Prints the reference's lines when ``verbose`` and returns [(true_text, predicted_text, accuracy %)]."""
    model.eval()
    model.to(device)
    results = []
    if verbose:
        print("\nModel Evaluation:")
        print("-" * 50)
    with torch.no_grad():
        for i, (videos, labels, label_lengths) in enumerate(test_loader):
            if i >= num_samples:
                break
            outputs = model(videos.to(device))
            texts = decode_batch(outputs, dataset)
            for j in range(videos.size(0)):
                n = i * test_loader.batch_size + j
                if n >= num_samples:
                    break
                true_label = labels[j][:label_lengths[j]]
                true_text = "".join(dataset.idx_to_char.get(int(t), "") for t in true_label if int(t) != 0)
                correct = sum(1 for a, b in zip(true_text, texts[j]) if a == b)
                acc = correct / max(len(true_text), 1) * 100
                results.append((true_text, texts[j], acc))
                if verbose:
                    print(f"\nSample {n + 1}:")
                    print(f"True text: {true_text}")
                    print(f"Predicted text: {texts[j]}")
                    print(f"Character accuracy: {acc:.2f}%")
    return results
