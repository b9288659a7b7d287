"""Evidence for DESIGN.md section 4.1: the north star sketches the STFT as a tcgen05 GEMM; this times exactly
that formulation — the windowed-DFT GEMM [frames, 2048] x [2048, 2*1152] (cos | sin, 1025 bins padded to
1152) with the hi/lo-split tcgen05 kernel that fp32-grade log-mel needs — against the whole FFT-based K1
(log-mel for the 746 unique frames + statistics for 41 shifts).  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

N_ = A._native
L = N_.lib()
clips = int(os.environ.get("K1_CLIPS", "64"))
frames_per_clip = 746


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator().manual_seed(0)
audio = (torch.randn((clips, 48000), generator=g) * 0.1).clamp_(-1, 1).cuda()
shifts = [640 * k for k in range(-20, 21)]
import ctypes
A.audio_stats_sweep(audio, shifts)
torch.cuda.synchronize()
L.avs_prof_reset()
L.avs_prof_enable(1)
for _ in range(5):
    A.audio_stats_sweep(audio, shifts)
torch.cuda.synchronize()
L.avs_prof_enable(0)
t_fft = 0.0
for slot in (5, 6):                                     # log-mel + statistics kernels, CUDA events around each launch
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
    t_fft += t.value / max(c.value, 1)

M, K, N = clips * frames_per_clip, 2048, 2 * 1152
fr = torch.randn((M, K), generator=g).cuda()           # stands in for the windowed frames (6.1 MB per clip to materialise)
k = torch.arange(1152, dtype=torch.float64)[:, None] * torch.arange(K, dtype=torch.float64)[None, :] * (2 * torch.pi / K)
dft = torch.cat([torch.cos(k), -torch.sin(k)]).float().cuda()
bias = torch.zeros(N, device="cuda")
out = torch.empty((M, N), device="cuda")
ws = N_.workspace(L.avs_gemm_split_workspace_bytes(M, N, K), "cuda")
t_gemm = timed(lambda: N_.check(L.avs_gemm_split(N_.ptr(fr), N_.ptr(dft), N_.ptr(bias), N_.ptr(out), M, N, K, N_.ptr(ws),
                                                 ws.numel(), N_.stream_ptr())))
flop = 3 * 2.0 * M * N * K
print(f"{clips} clips: FFT-based K1 (whole stage, 41 shifts) {1e3 * t_fft / clips:.2f} us/clip | "
      f"DFT-as-GEMM, DFT step alone (tcgen05 hi/lo split, incl. operand packing) {1e3 * t_gemm / clips:.2f} us/clip "
      f"= {flop / (t_gemm * 1e-3) / 1e12:.0f} TFLOP/s of issued MMA work")
