"""Time the LipNet eval path (config 4: B=256 forward + greedy decode) stage by stage.  GPU box only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

if os.environ.get("K3_LIB"):   # a variant build: make -C csrc VARIANT=name VARIANT_FLAGS=...
    A._native.LIB_PATH = os.path.join(os.path.dirname(A._native.LIB_PATH), f"libavsync_b200_var_{os.environ['K3_LIB']}.so")
B = int(os.environ.get("K3_CLIPS", "256"))
prec = os.environ.get("K3_PRECISION", "bf16x3")
torch.manual_seed(0)
net = A.LipNet(39, precision=prec).cuda().eval()
frames = torch.rand((B, 1, 75, 50, 100), generator=torch.Generator().manual_seed(3)).cuda()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


t_stcnn, emb = timed(lambda: net.stcnn(frames))
t_gru, logp = timed(lambda: net.gru_head(emb))
t_dec, (ids, lens) = timed(lambda: A.ctc_greedy_decode(logp))
print(f"B={B} precision={prec}: stcnn {t_stcnn:.2f} ms ({1e3 * t_stcnn / B:.1f} us/clip) | gru_head {t_gru:.2f} ms "
      f"({1e3 * t_gru / B:.1f} us/clip) | ctc {t_dec:.3f} ms | total {B / (t_stcnn + t_gru + t_dec) * 1e3:.0f} clips/s")

import ctypes
L = A._native.lib()
L.avs_prof_reset()
L.avs_prof_enable(1)
for _ in range(3):
    net.gru_head(emb)
torch.cuda.synchronize()
L.avs_prof_enable(0)
parts = []
for slot, name in ((9, "pack"), (10, "gemm"), (11, "recurrence"), (12, "fc+softmax")):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
    parts.append(f"{name} {t.value / 3:.3f} ms ({c.value // 3} launches)")
print("gru_head breakdown per forward: " + " | ".join(parts))
