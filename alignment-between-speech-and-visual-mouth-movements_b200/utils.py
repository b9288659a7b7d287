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


def decode_metrics(pred_ids: torch.Tensor, pred_lens: torch.Tensor, target_ids: torch.Tensor, target_lens: torch.Tensor,
                   space_id: int = 37, pad_id: int = 38):
    """Batched CER / WER / positional character accuracy on DEVICE id sequences, with the text
    semantics of the reference: ``calculate_cer`` / ``calculate_wer`` (train.py:945-993) on the strings the ids
    render to (id 38 renders as the 5 characters '<pad>'), and ``evaluate_model``'s accuracy (utils.py:83-86).
    Returns a dict of CPU float tensors [B]: cer, wer, char_accuracy (percent), plus the raw int table."""
    for t, n in ((pred_ids, "pred_ids"), (pred_lens, "pred_lens"), (target_ids, "target_ids"), (target_lens, "target_lens")):
        N.require_cuda(t, n)
    pred_ids, target_ids = pred_ids.to(torch.int32).contiguous(), target_ids.to(torch.int32).contiguous()
    pred_lens, target_lens = pred_lens.to(torch.int32).contiguous(), target_lens.to(torch.int32).contiguous()
    B = pred_ids.shape[0]
    out = torch.empty((B, 6), dtype=torch.int32, device=pred_ids.device)
    if B:
        N.check(N.lib().avs_edit_metrics(N.ptr(pred_ids), N.ptr(pred_lens), pred_ids.shape[1], N.ptr(target_ids),
                                         N.ptr(target_lens), target_ids.shape[1], B, space_id, pad_id, 16, 1, 4,
                                         N.ptr(out), N.stream_ptr()), "edit_metrics")
    r = out.cpu().to(torch.float64)
    pred_n, tgt_n, tgt_w = r[:, 5], r[:, 1], r[:, 3]
    cer = torch.where(tgt_n > 0, r[:, 0] / tgt_n.clamp(min=1), (pred_n > 0).to(torch.float64))
    pred_w_nonzero = (r[:, 2] > 0) | (tgt_w > 0)          # target empty: WER = 1 if the prediction has any word
    wer = torch.where(tgt_w > 0, r[:, 2] / tgt_w.clamp(min=1), ((r[:, 2] > 0) & pred_w_nonzero).to(torch.float64))
    acc = r[:, 4] / tgt_n.clamp(min=1) * 100.0
    return {"cer": cer, "wer": wer, "char_accuracy": acc, "raw": out.cpu()}
