"""CPU restatement of the sample-rate conversion in front of the MFCC stage.

TEST INFRASTRUCTURE (oracle).  The reference resamples with ``librosa.resample(audio, orig_sr=sr, target_sr=16000)``
(``misalignment_detection_train.py:202-204``); librosa's default filter is ``soxr_hq`` — the un-vendored, un-pinned C
library libsoxr (no source, wheel or test vectors in ``/root/reference``, none installable here): **parity with librosa
itself is unpinned at this boundary.**  The B200 path and this oracle implement the band-limited Kaiser-windowed-sinc
interpolator librosa shipped as ``kaiser_best`` before 0.10 (resampy: 64 zero crossings, roll-off 0.9475937167399596,
beta 14.769656459379492), in the published formulation of ``torchaudio.functional.resample`` (``_get_sinc_resample_kernel``
/ ``_apply_sinc_resample_kernel``), which ``tests/test_oracle.py`` pins this file against — plus a looser check against the
independent ``scipy.signal.resample_poly`` on band-limited signals.
"""
from __future__ import annotations

import math

import numpy as np

ZEROS = 64
ROLLOFF = 0.9475937167399596
BETA = 14.769656459379492


def polyphase_kernel(orig_sr: int, target_sr: int):
    """(h [new, 2*width + orig] float64, width, orig, new) for the reduced rates."""
    g = math.gcd(int(orig_sr), int(target_sr))
    orig, new = int(orig_sr) // g, int(target_sr) // g
    base = min(orig, new) * ROLLOFF
    width = math.ceil(ZEROS * orig / base)
    k = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (k - np.arange(new, dtype=np.float64)[:, None] / new) * base
    t = np.clip(t, -ZEROS, ZEROS)
    window = np.i0(BETA * np.sqrt(np.maximum(1.0 - (t / ZEROS) ** 2, 0.0))) / np.i0(BETA)
    a = t * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(a == 0.0, 1.0, np.sin(a) / a)
    return sinc * window * (base / orig), width, orig, new


def resample(x: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """float array [n] -> float32 [ceil(n * target / orig)], float64 accumulation."""
    x = np.asarray(x, dtype=np.float64)
    if int(orig_sr) == int(target_sr):
        return x.astype(np.float32)
    h, width, orig, new = polyphase_kernel(orig_sr, target_sr)
    n = x.shape[0]
    n_out = -(-new * n // orig)
    n_blocks = -(-n_out // new)
    padded = np.zeros(width + (n_blocks - 1) * orig + h.shape[1] + 1, dtype=np.float64)
    padded[width:width + n] = x
    idx = np.arange(n_blocks)[:, None] * orig + np.arange(h.shape[1])[None, :]     # [blocks, taps]
    y = padded[idx] @ h.T                                                           # [blocks, new]
    return y.reshape(-1)[:n_out].astype(np.float32)
