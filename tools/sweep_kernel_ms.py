"""Per-kernel CUDA-event times of the full sweep step (device-resident, 1 GPU) through the EXPERIMENTS build, so that the
AVS_* environment knobs can be A/B-ed:  AVS_K1_SCHED=0|1 python tools/sweep_kernel_ms.py [clips] [steps] [precision]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A
import bench

A._native.use_experiments_build()
L = A._native.lib()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
PREC = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
net = A.LipNet(39, precision=PREC).cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
sw = A.SyncSweeper(net, det, 20, 48000, chunk_clips=min(128, C))
fr, au = bench.synth_inputs(C, seed=1000)
fr, au = fr.cuda(), au.cuda()
for _ in range(3):
    sw.run(fr, au)
torch.cuda.synchronize()
L.avs_prof_reset()
L.avs_prof_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    sw.run(fr, au)
e1.record()
torch.cuda.synchronize()
L.avs_prof_enable(0)
ms = e0.elapsed_time(e1) / STEPS
names = ["pack", "conv1", "conv2", "conv3", "vstats", "mfcc_logmel", "mfcc_stats", "score_gemm", "score"]
out = []
for i, n in enumerate(names):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(i, ctypes.byref(t), ctypes.byref(c))
    out.append(f"{n} {t.value / STEPS:.2f}")
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("AVS_"))
print(f"[{knobs or 'defaults'}] {PREC} {C} clips: {ms:.2f} ms/step = {C / ms * 1e3:.0f} clips/s | " + " | ".join(out), flush=True)
