"""K1 (MFCC statistics of all shifts) in isolation: CUDA-event times of its two kernels.  --lib NAME loads
libavsync_b200_var_NAME.so (make -C csrc VARIANT=NAME VARIANT_FLAGS=...), default the product library.  GPU box only."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

args = sys.argv[1:]
LIB = "product"
if "--lib" in args:
    LIB = args[args.index("--lib") + 1]
if LIB != "product":
    A._native.LIB_PATH = os.path.join(os.path.dirname(A._native.LIB_PATH), f"libavsync_b200_var_{LIB}.so")
L = A._native.lib()
audio = (torch.randn((64, 48000), generator=torch.Generator().manual_seed(3)) * 0.1).clamp_(-1, 1).cuda()
shifts = [640 * k for k in range(-20, 21)]
for _ in range(2):
    A.audio_stats_sweep(audio, shifts)
torch.cuda.synchronize()
L.avs_prof_reset()
L.avs_prof_enable(1)
for _ in range(3):
    A.audio_stats_sweep(audio, shifts)
torch.cuda.synchronize()
L.avs_prof_enable(0)
out = []
for slot, name in ((5, "mfcc_logmel"), (6, "mfcc_stats")):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
    out.append(f"{name} {1e3 * t.value / max(c.value, 1) / 64:.2f} us/clip")
print(f"[lib={LIB}] K1 alone, 64 clips: " + " | ".join(out), flush=True)
