"""Timeline of ONE sweep step (device-resident, 1 GPU, EXPERIMENTS build): begin / end of every profiled launch in ms since
the step's first pack kernel, per chunk, and the idle gaps of the main stream.  AVS_* knobs apply.
    python tools/sweep_timeline.py [clips]"""
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
torch.manual_seed(0)
net = A.LipNet(39, precision="bf16").cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
sw = A.SyncSweeper(net, det, 20, 48000, chunk_clips=min(128, C))
fr, au = bench.synth_inputs(C, seed=1000)
fr, au = fr.cuda(), au.cuda()
for _ in range(4):
    sw.run(fr, au)
torch.cuda.synchronize()
L.avs_prof_reset()
L.avs_prof_enable(1)
sw.run(fr, au)
torch.cuda.synchronize()
L.avs_prof_enable(0)
names = ["pack", "conv1", "conv2", "conv3", "vstats", "logmel", "stats", "k4gemm", "k4score"]
main = {0, 1, 2, 3, 4, 7, 8}
spans = []
for slot, n in enumerate(names):
    t, c = ctypes.c_double(), ctypes.c_int()
    L.avs_prof_read(slot, ctypes.byref(t), ctypes.byref(c))
    for i in range(c.value):
        b, e = ctypes.c_double(), ctypes.c_double()
        A._native.check(L.avs_prof_read_span(slot, i, ctypes.byref(b), ctypes.byref(e)))
        spans.append((b.value, e.value, n, slot in main))
spans.sort()
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("AVS_"))
print(f"[{knobs or 'defaults'}] {C} clips, one step; ms since the first pack kernel")
last_main_end, gap_total = 0.0, 0.0
for b, e, n, is_main in spans:
    note = ""
    if is_main:
        gap = b - last_main_end
        if last_main_end > 0:
            gap_total += max(gap, 0.0)
            note = f"   main-stream gap {gap * 1e3:7.1f} us"
        last_main_end = e
    print(f"{'' if is_main else '        '}{n:8s} {b:8.3f} -> {e:8.3f}  ({(e - b) * 1e3:8.1f} us){note}")
print(f"step {last_main_end:.3f} ms, main-stream gaps {gap_total:.3f} ms")
