"""GPU pre-processing prologue for the path: the per-frame arithmetic of the
reference's ``GridDataset.process_video`` (dataset.py:176-256) — gray, mouth crop, bilinear resize to
100x50, /255, pad/truncate to 75 frames — as one kernel, bit-exact with OpenCV's 8-bit paths.
Directory discovery, ``.align`` parsing and video DECODING stay on the host (file / codec IO)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N

VOCAB_CHARS = "abcdefghijklmnopqrstuvwxyz0123456789 "


def create_vocab():
    """(char_to_idx, idx_to_char) of GridDataset._create_vocab (dataset.py:38-46): 37 characters -> 1..37,
    '<blank>' = 0, '<pad>' = 38."""
    char_to_idx = {c: i + 1 for i, c in enumerate(VOCAB_CHARS)}
    char_to_idx["<blank>"] = 0
    char_to_idx["<pad>"] = len(VOCAB_CHARS) + 1
    return char_to_idx, {i: c for c, i in char_to_idx.items()}


class GridPreprocessor:
    """Decoded uint8 frames -> float32 [B, 1, 75, 50, 100] on the GPU."""

    def __init__(self):
        self._plans: Dict[tuple, N.Handle] = {}
        self.vocab, self.idx_to_char = create_vocab()

    def _plan(self, h: int, w: int, c: int) -> N.Handle:
        key = (torch.cuda.current_device(), h, w, c)
        p = self._plans.get(key)
        if p is None:
            N.device_check()
            hnd = N.c_void_p()
            N.check(N.lib().avs_preproc_create(h, w, c, ctypes.byref(hnd)), "preproc_create")
            p = self._plans[key] = N.Handle(hnd, N.lib().avs_preproc_destroy)
        return p

    def crop_box(self, h: int, w: int, channels: int = 3):
        y0, x0, ch, cw = (ctypes.c_int() for _ in range(4))
        N.check(N.lib().avs_preproc_crop(self._plan(h, w, channels).h, *(ctypes.byref(v) for v in (y0, x0, ch, cw))))
        return y0.value, x0.value, ch.value, cw.value

    def process_batch(self, frames: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: CUDA uint8 [B, n, h, w, 3] (BGR) or [B, n, h, w] (gray); lengths: valid frames per clip."""
        N.require_cuda(frames, "frames")
        if frames.dtype != torch.uint8 or frames.dim() not in (4, 5):
            raise RuntimeError("frames must be uint8 [B, n, h, w, 3] or [B, n, h, w]")
        frames = frames.contiguous()
        B, n, h, w = frames.shape[:4]
        c = frames.shape[4] if frames.dim() == 5 else 1
        out = torch.empty((B, 1, 75, 50, 100), dtype=torch.float32, device=frames.device)
        if n == 0:
            return out.zero_()
        ln = None
        if lengths is not None:
            ln = lengths.to(device=frames.device, dtype=torch.int32).contiguous()
        N.check(N.lib().avs_preproc_run(self._plan(h, w, c).h, N.ptr(frames), B, n, N.ptr(ln), N.ptr(out),
                                        N.stream_ptr()), "preproc_run")
        return out

    def process_frames(self, frames: np.ndarray) -> torch.Tensor:
        """One clip of decoded frames (numpy uint8 [n, h, w, 3] or [n, h, w]) -> CUDA float32 [1, 75, 50, 100],
        the tensor ``GridDataset.process_video`` returns for that clip."""
        f = torch.from_numpy(np.ascontiguousarray(frames[:75])).cuda()
        return self.process_batch(f.unsqueeze(0))[0]

    def process_video(self, video_path: str) -> torch.Tensor:
        """Reference entry point (dataset.py:176): ``.npy`` clips of the right shape are normalised and padded
        on the host exactly as the reference does; video files are decoded with cv2 on the host and processed on
        the GPU.  Returns a CPU tensor like the reference."""
        if video_path.endswith(".npy"):
            frames = np.load(video_path)
            if frames.max() > 1.0:
                frames = frames / 255.0
            if frames.shape[1:] != (50, 100):
                raise RuntimeError("pre-processed .npy clips must already be 50 x 100 (float resize is host-side only)")
            frames = frames[:75]
            if len(frames) < 75:
                frames = np.concatenate([frames, np.zeros((75 - len(frames), 50, 100))], axis=0)
            return torch.FloatTensor(frames).unsqueeze(0)
        import cv2
        cap = cv2.VideoCapture(video_path)
        decoded = []
        while len(decoded) < 75:
            ret, frame = cap.read()
            if not ret:
                break
            decoded.append(frame)
        cap.release()
        if not decoded:
            print(f"Warning: No frames extracted from {video_path}")
            return torch.zeros((1, 75, 50, 100))
        return self.process_frames(np.stack(decoded)).cpu()
