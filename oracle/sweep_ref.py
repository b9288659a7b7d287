"""CPU restatement of the misalignment feature pipeline and the +-S sync sweep.

TEST INFRASTRUCTURE (oracle).  Follows ``misalignment_detection_train.py``:
``shift_audio`` :100-114, ``compute_audio_stats`` :117-127, visual stats :165,
``FeatureExtractor.build_feature`` :199-208, ``MisalignmentDetector`` :237-250,
score = sigmoid(logit) :267 / ``misalignment_detection_demo.py:249``.  The sweep
itself is the composition SURVEY.md section 3.2 describes (the reference has no
sweep function): visual stats once per clip, audio stats once per (clip, k).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from . import lipnet_ref, mfcc_ref


def shift_audio(audio: np.ndarray, shift_frames: int, fps: float, sample_rate: int) -> np.ndarray:
    """misalignment_detection_train.py:100-114 (integer delay, zero fill)."""
    if shift_frames == 0:
        return audio.copy()
    s = int(shift_frames / max(fps, 1e-5) * sample_rate)
    if s == 0:
        return audio.copy()
    out = np.zeros_like(audio)
    n = len(audio)
    if s > 0:
        if s < n:
            out[s:] = audio[:n - s]
    else:
        s = -s
        if s < n:
            out[:n - s] = audio[s:]
    return out


def shift_samples(shift_frames: int, fps: float, sample_rate: int) -> int:
    """The integer sample delay ``shift_audio`` applies (:103), sign kept."""
    if shift_frames == 0:
        return 0
    return int(shift_frames / max(fps, 1e-5) * sample_rate)


def compute_audio_stats(audio: np.ndarray, sample_rate: int, n_mfcc: int) -> torch.Tensor:
    """misalignment_detection_train.py:117-127: [mean(n_mfcc), unbiased std(n_mfcc)]."""
    if audio.size == 0:
        return torch.zeros(n_mfcc * 2, dtype=torch.float32)
    hop = max(1, int(sample_rate / 40))
    m = mfcc_ref.mfcc(audio, sr=sample_rate, n_mfcc=n_mfcc, hop_length=hop)
    if m.size == 0:
        return torch.zeros(n_mfcc * 2, dtype=torch.float32)
    mt = torch.from_numpy(np.ascontiguousarray(m.T)).float()
    return torch.cat([mt.mean(dim=0), mt.std(dim=0)], dim=0)


def visual_stats(emb: torch.Tensor) -> torch.Tensor:
    """:165 — emb [T,F] -> [2F] = cat(mean over T, unbiased std over T)."""
    return torch.cat([emb.mean(dim=0), emb.std(dim=0)], dim=0)


def init_detector_state(input_dim: int = 13864, hidden: int = 512, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init ``MisalignmentDetector`` weights (:237-247), keys ``classifier.{0,3}.*``."""
    import torch.nn as nn
    torch.manual_seed(seed)
    seq = nn.Sequential(nn.Linear(input_dim, hidden), nn.ReLU(), nn.Dropout(0.3), nn.Linear(hidden, 1))
    return {f"classifier.{k}": v.detach().clone() for k, v in seq.state_dict().items()}


def detector_logits(det: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """``MisalignmentDetector.forward`` (:249-250), eval mode."""
    h = torch.relu(x @ det["classifier.0.weight"].t() + det["classifier.0.bias"])
    return (h @ det["classifier.3.weight"].t() + det["classifier.3.bias"]).squeeze(-1)


def sweep_clip(sd, det, frames: torch.Tensor, audio: np.ndarray, shifts: Sequence[int],
               fps: float = 25.0, sr: int = 16000, n_mfcc: int = 20, batched: bool = False):
    """Reference-style sweep for ONE clip.  frames [1,T,H,W] f32; audio f32[n].

    ``batched=False`` mirrors the reference call pattern (one detector call per
    shift, B=1 STCNN); ``batched=True`` is the "best-effort CPU" variant (one
    detector call for all shifts).  Returns dict(scores[K], best, vstats, astats[K,2*n_mfcc]).
    """
    with torch.no_grad():
        emb = lipnet_ref.stcnn(sd, frames.unsqueeze(0))[0]
        v = visual_stats(emb)
        a = [compute_audio_stats(shift_audio(audio, int(k), fps, sr), sr, n_mfcc) for k in shifts]
        if batched:
            x = torch.stack([torch.cat([v, ak]) for ak in a])
            scores = torch.sigmoid(detector_logits(det, x))
        else:
            scores = torch.stack([
                torch.sigmoid(detector_logits(det, torch.cat([v, ak]).unsqueeze(0)))[0] for ak in a])
    return {"scores": scores.numpy(), "best": int(np.argmax(scores.numpy())),
            "vstats": v.numpy(), "astats": torch.stack(a).numpy()}


# ---------------------------------------------------------------- synthetic inputs
def synth_frames(n: int, seed: int = 1234, T: int = 75, H: int = 50, W: int = 100) -> torch.Tensor:
    """float32 [n,1,T,H,W] ~ U[0,1) (what /255 grayscale gives)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, 1, T, H, W), generator=g, dtype=torch.float32)


def synth_audio(n: int, seed: int = 1234, n_samples: int = 48000, kind: str = "noise") -> np.ndarray:
    """float32 [n, n_samples].  'noise' = N(0,0.1^2) clipped to [-1,1]; 'halfsilent' =
    noise with the last third zeroed; 'chirp' = 0.5*sin sweep 100..6000 Hz; 'speechlike'
    = amplitude-modulated band noise (per-clip random envelope)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / 16000.0
    out = np.empty((n, n_samples), dtype=np.float32)
    for i in range(n):
        if kind == "chirp":
            f0, f1 = 100.0 + 20 * i, 6000.0
            ph = 2 * np.pi * (f0 * t + (f1 - f0) / (2 * t[-1]) * t * t)
            x = 0.5 * np.sin(ph)
        else:
            x = np.clip(rng.normal(0.0, 0.1, n_samples), -1.0, 1.0)
            if kind == "halfsilent":
                x[(2 * n_samples) // 3:] = 0.0
            elif kind == "speechlike":
                env = np.abs(np.interp(t, np.linspace(0, t[-1], 13), rng.uniform(0, 1, 13))) ** 2
                x = x * env
        out[i] = x.astype(np.float32)
    return out
