"""Measured numbers for the other BASELINE.json configs (they are parity-test cases, not bench.py lines):
config 2 = 64 clips, +-15-frame sweep, fp32-grade (bf16x3) on 1 GPU; config 4 = LipNet eval batch 256 +
greedy decode.  Prints one JSON line each.  GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator().manual_seed(3)
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
# config 1: ONE clip, +-15 frames: STCNN + Bi-GRU head (log-probs) + detector scores — the reference's CPU-runnable case
torch.manual_seed(0)
net1 = A.LipNet(39, precision="bf16x3").cuda().eval()
f1 = torch.rand((1, 1, 75, 50, 100), generator=g).cuda()
a1 = (torch.randn((1, 48000), generator=g) * 0.1).clamp_(-1, 1).cuda()
sw1 = A.SyncSweeper(net1, det, 15, chunk_clips=1)


def config1():
    sw1.run(f1, a1)
    return A.ctc_greedy_decode(net1(f1))


ms = timed(config1, n=20)
print(json.dumps({"config": "1: one clip, +-15 frames, STCNN + Bi-GRU head + detector sweep + decode", "precision": "bf16x3",
                  "ms": ms, "clips_per_s": 1e3 / ms}))
for prec in ("bf16x3", "bf16"):
    torch.manual_seed(0)
    net = A.LipNet(39, precision=prec).cuda().eval()
    frames = torch.rand((64, 1, 75, 50, 100), generator=g).cuda()
    audio = (torch.randn((64, 48000), generator=g) * 0.1).clamp_(-1, 1).cuda()
    sw = A.SyncSweeper(net, det, 15, chunk_clips=64)
    ms = timed(lambda: sw.run(frames, audio))
    print(json.dumps({"config": "2: batch 64, +-15 frames (31 offsets), 1 B200", "precision": prec, "ms": ms,
                      "clips_per_s": 64 / ms * 1e3}))
    f256 = torch.rand((256, 1, 75, 50, 100), generator=g).cuda()
    ms = timed(lambda: A.ctc_greedy_decode(net(f256)), n=3, warm=2)
    print(json.dumps({"config": "4: LipNet eval batch 256 + greedy CTC decode", "precision": prec, "ms": ms,
                      "clips_per_s": 256 / ms * 1e3}))
