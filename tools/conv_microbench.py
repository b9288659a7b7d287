"""Attribute the time of the tcgen05 conv layers: run the STCNN with the experiment switches of
avs_debug_set (1 = weights loaded once, 2 = activations loaded once, 4 = no epilogue work) and with
different pipeline depths, and print per-layer CUDA-event times.  GPU box only."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

A._native.use_experiments_build()   # libavsync_b200_exp.so: avs_debug_set + AVS_* knobs (make -C csrc EXPERIMENTS=1)
L = A._native.lib()
B = int(os.environ.get("MB_CLIPS", "32"))
frames = torch.rand((B, 1, 75, 50, 100), generator=torch.Generator().manual_seed(3)).cuda()


def run(precision, flags, label):
    torch.manual_seed(0)
    net = A.LipNet(39, precision=precision).cuda().eval()
    L.avs_debug_set(0)
    net.stcnn(frames)
    L.avs_debug_set(flags)
    for _ in range(2):
        net.stcnn(frames)
    torch.cuda.synchronize()
    L.avs_prof_reset()
    L.avs_prof_enable(1)
    for _ in range(3):
        net.stcnn(frames)
    torch.cuda.synchronize()
    L.avs_prof_enable(0)
    L.avs_debug_set(0)
    out = []
    for slot, name in ((0, "pack"), (1, "conv1"), (2, "conv2"), (3, "conv3")):
        t, c = ctypes.c_double(), ctypes.c_int()
        L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
        out.append(f"{name} {1e3 * t.value / max(c.value, 1) / B:8.2f} us/clip")
    print(f"{label:44s} " + " | ".join(out), flush=True)


FLAGS = os.environ.get("MB_FLAGS")
if FLAGS:
    for f in FLAGS.split(","):
        run("bf16", int(f), f"bf16 dbg={f}")
    sys.exit(0)
QUICK = os.environ.get("MB_QUICK") == "1"
if QUICK:
    run("bf16", 0, "bf16 dbg=0")
    run("bf16x3", 0, "bf16x3 dbg=0")
    sys.exit(0)
for prec in ("bf16",):
    for flags in (0, 7, 8, 15):
        run(prec, flags, f"{prec} dbg={flags}")
run("bf16x3", 0, "bf16x3 dbg=0")
run("bf16x3", 7, "bf16x3 dbg=7")
run("bf16x3", 15, "bf16x3 dbg=15")

# ---- K1 in isolation (nothing else on the GPU)
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
for slot, name in ((5, "mfcc_logmel"), (6, "mfcc_stats")):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
    print(f"{name} alone: {1e3 * t.value / max(c.value, 1) / 64:8.2f} us/clip", flush=True)
