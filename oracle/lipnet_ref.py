"""torch-CPU restatement of the reference LipNet forward and greedy CTC decode.

TEST INFRASTRUCTURE (oracle).  Functional (weights come in as a ``state_dict``
with the reference's key names) so it can run on the GPU box where
``/root/reference`` does not exist.  Pinned against the reference's own
``model.LipNet`` / ``utils.decode_prediction`` by ``oracle/make_golden.py``
(fixtures in ``tests/golden/``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

VOCAB_CHARS = "abcdefghijklmnopqrstuvwxyz0123456789 "


def make_vocab() -> Dict[int, str]:
    """idx_to_char of ``GridDataset._create_vocab`` (dataset.py:38-46): 37 chars
    -> 1..37, '<blank>' = 0, '<pad>' = 38."""
    idx_to_char = {i + 1: c for i, c in enumerate(VOCAB_CHARS)}
    idx_to_char[0] = "<blank>"
    idx_to_char[len(VOCAB_CHARS) + 1] = "<pad>"
    return idx_to_char


def init_lipnet_state(vocab_size: int = 39, hidden: int = 256, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the shapes/keys of ``model.LipNet`` (model.py:22-48),
    PyTorch default initialisers under ``torch.manual_seed(seed)``.  Built from
    the same ``nn`` layers in the same order so the values equal the reference's
    ``LipNet(vocab_size)`` under the same seed (checked by make_golden)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    layers = [
        ("conv1", nn.Conv3d(1, 32, (3, 5, 5), padding=(1, 2, 2))),
        ("conv2", nn.Conv3d(32, 64, (3, 5, 5), padding=(1, 2, 2))),
        ("conv3", nn.Conv3d(64, 96, (3, 3, 3), padding=(1, 1, 1))),
        ("gru1", nn.GRU(96 * 6 * 12, hidden, batch_first=True, bidirectional=True)),
        ("gru2", nn.GRU(hidden * 2, hidden, batch_first=True, bidirectional=True)),
        ("fc", nn.Linear(hidden * 2, vocab_size)),
    ]
    sd = {}
    for name, mod in layers:
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def stcnn(sd: Dict[str, torch.Tensor], frames: torch.Tensor, return_intermediates: bool = False):
    """STCNN half of ``LipNet.forward`` (model.py:67-82) ==
    ``extract_visual_embeddings`` (misalignment_detection_train.py:130-144), eval
    mode (Dropout3d = identity).  frames [B,1,T,H,W] f32 -> [B,T,96*6*12]."""
    x = F.relu(F.conv3d(frames, sd["conv1.weight"], sd["conv1.bias"], padding=(1, 2, 2)))
    p1 = F.max_pool3d(x, (1, 2, 2))
    x = F.relu(F.conv3d(p1, sd["conv2.weight"], sd["conv2.bias"], padding=(1, 2, 2)))
    p2 = F.max_pool3d(x, (1, 2, 2))
    x = F.relu(F.conv3d(p2, sd["conv3.weight"], sd["conv3.bias"], padding=(1, 1, 1)))
    p3 = F.max_pool3d(x, (1, 2, 2))
    b, c, t, h, w = p3.shape
    emb = p3.permute(0, 2, 1, 3, 4).contiguous().view(b, t, -1)
    if return_intermediates:
        return emb, p1, p2, p3
    return emb


def _gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One direction of nn.GRU (gate order r,z,n; h0 = 0), batch_first."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi_all = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    out = x.new_zeros(B, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[:, t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out


def bigru(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    f = _gru_dir(x, sd[f"{prefix}.weight_ih_l0"], sd[f"{prefix}.weight_hh_l0"],
                 sd[f"{prefix}.bias_ih_l0"], sd[f"{prefix}.bias_hh_l0"], False)
    b = _gru_dir(x, sd[f"{prefix}.weight_ih_l0_reverse"], sd[f"{prefix}.weight_hh_l0_reverse"],
                 sd[f"{prefix}.bias_ih_l0_reverse"], sd[f"{prefix}.bias_hh_l0_reverse"], True)
    return torch.cat([f, b], dim=-1)


def gru_head(sd, emb: torch.Tensor) -> torch.Tensor:
    """Bi-GRU head of ``LipNet.forward`` (model.py:84-95), eval mode:
    gru1 -> gru2 -> fc -> log_softmax.  [B,T,6912] -> [B,T,vocab]."""
    x = bigru(sd, "gru1", emb)
    x = bigru(sd, "gru2", x)
    x = x @ sd["fc.weight"].t() + sd["fc.bias"]
    return F.log_softmax(x, dim=-1)


def lipnet_forward(sd, frames: torch.Tensor) -> torch.Tensor:
    """``LipNet.forward`` (model.py:57-97) in eval mode."""
    with torch.no_grad():
        return gru_head(sd, stcnn(sd, frames))


def greedy_ids(logp: np.ndarray, blank: int = 0) -> List[int]:
    """Collapse rule of ``decode_prediction`` (utils.py:20-30): argmax (first index
    on ties), emit c iff c != prev and c != blank, prev = c."""
    pred = np.asarray(torch.max(torch.as_tensor(logp), dim=-1)[1])
    out, prev = [], blank
    for c in pred.tolist():
        if c != prev and c != blank:
            out.append(c)
        prev = c
    return out


def ids_to_text(ids, idx_to_char=None) -> str:
    """utils.py:33-34."""
    idx_to_char = idx_to_char or make_vocab()
    return "".join(idx_to_char.get(i, "") for i in ids if i in idx_to_char)


def decode_prediction(logp, idx_to_char=None, blank: int = 0) -> str:
    return ids_to_text(greedy_ids(logp, blank), idx_to_char)
