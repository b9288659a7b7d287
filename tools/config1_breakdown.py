"""Config 1 (ONE clip: LipNet forward + greedy CTC + +-15 sweep, fp32-grade) — where the ~1 ms goes: per-kernel CUDA-event
times (avs_prof_*), the CUDA-event time of the whole call sequence, and the host time spent enqueuing it.
    python tools/config1_breakdown.py [precision]"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A
import bench

PREC = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
L = A._native.lib()
torch.manual_seed(0)
net = A.LipNet(39, precision=PREC).cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
fr_u8, au = bench.synth_inputs(1, seed=1000)
fr = bench.frames_f32(fr_u8, pin=False).cuda()
au = au.cuda()
sw = A.SyncSweeper(net, det, 15, 48000, chunk_clips=1)


def cfg1():
    lp = net(fr)
    A.ctc_greedy_decode(lp)
    sw.run(fr, au)


for _ in range(5):
    cfg1()
torch.cuda.synchronize()
N = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(N):
    cfg1()
e1.record()
t_enq = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
print(f"[{PREC}] config 1: {e0.elapsed_time(e1) / N:.3f} ms per clip on the device, {t_enq:.3f} ms of host time to enqueue it")
L.avs_prof_reset()
L.avs_prof_enable(1)
for _ in range(N):
    cfg1()
torch.cuda.synchronize()
L.avs_prof_enable(0)
names = ["pack", "conv1", "conv2", "conv3", "vstats", "mfcc_logmel", "mfcc_stats", "score_gemm", "score", "gru_pack",
         "gru_gemm", "gru_recurrence", "fc_softmax"]
tot = 0.0
for i, n in enumerate(names):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(i, ctypes.byref(t), ctypes.byref(c))
    tot += t.value / N
    print(f"  {n:15s} {t.value / N * 1e3:8.1f} us  ({c.value // N} launches)")
print(f"  sum of profiled kernels {tot * 1e3:.1f} us")
