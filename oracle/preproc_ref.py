"""CPU restatement of the per-frame pre-processing of ``GridDataset.process_video``
(dataset.py:199-254): BGR->gray, crop, resize, /255, pad/truncate.  TEST INFRASTRUCTURE (oracle).

``process_frames`` calls OpenCV exactly as the reference does (cv2 is a dependency of the reference and is
installed in this image); ``gray_u8`` / ``resize_linear_u8`` restate the two OpenCV 8-bit algorithms in
numpy and are checked against cv2 in tests/test_oracle.py — they document what the GPU kernel implements.
"""
from __future__ import annotations

import numpy as np
import torch


def gray_u8(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(COLOR_BGR2GRAY) for uint8: 15-bit fixed point, round to nearest."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _coeffs(dn: int, sn: int):
    scale = sn / dn
    idx = np.zeros(dn, np.int64)
    a0 = np.zeros(dn, np.int64)
    a1 = np.zeros(dn, np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= sn - 1:
            s, f = sn - 1, np.float32(0)
        idx[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh)) (INTER_LINEAR) for a uint8 plane with at least ``dh`` rows."""
    sh, sw = src.shape
    xi, xa0, xa1 = _coeffs(dw, sw)
    yi, ya0, ya1 = _coeffs(dh, sh)
    s = src.astype(np.int64)
    rows = s[:, xi] * xa0[None, :] + s[:, np.minimum(xi + 1, sw - 1)] * xa1[None, :]
    r0, r1 = rows[yi], rows[np.minimum(yi + 1, sh - 1)]
    return ((((ya0[:, None] * (r0 >> 4)) >> 16) + ((ya1[:, None] * (r1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def process_frames(frames: np.ndarray, img_width: int = 100, img_height: int = 50, max_video_length: int = 75) -> torch.Tensor:
    """The body of the reference's frame loop (dataset.py:209-254) applied to already-decoded frames
    [n, h, w, 3] (BGR) or [n, h, w] (gray), uint8 -> FloatTensor [1, 75, 50, 100]."""
    import cv2
    out = []
    for frame in frames:
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        h, w = gray.shape
        mouth = gray[int(h * 0.6):, int(w * 0.3):int(w * 0.7)]
        if mouth.size == 0:
            mouth = gray
        out.append(cv2.resize(mouth, (img_width, img_height)) / 255.0)
        if len(out) >= max_video_length:
            break
    if len(out) == 0:
        arr = np.zeros((max_video_length, img_height, img_width))
    else:
        arr = np.array(out)
    if len(arr) < max_video_length:
        arr = np.concatenate([arr, np.zeros((max_video_length - len(arr), img_height, img_width))], axis=0)
    else:
        arr = arr[:max_video_length]
    return torch.FloatTensor(arr).unsqueeze(0)
