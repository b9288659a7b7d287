"""Tiny driver for ncu: warm up, then ONE sweep chunk (K2 | K1 -> K4) and ONE LipNet head (K3 + K5), so that
`ncu --set full -k regex:avs -s <warm-up launches>` captures every kernel of the path once.  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

B = int(os.environ.get("PROF_CLIPS", "16"))
torch.manual_seed(0)
net = A.LipNet(39, precision="bf16").cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
g = torch.Generator().manual_seed(3)
frames = torch.rand((B, 1, 75, 50, 100), generator=g).cuda()
audio = (torch.randn((B, 48000), generator=g) * 0.1).clamp_(-1, 1).cuda()
sw = A.SyncSweeper(net, det, 20, chunk_clips=B)
L = A._native.lib()
for i in range(2):
    if i == 1:
        torch.cuda.synchronize()
        print("launches before the profiled pass:", L.avs_launch_count())
    scores, best = sw.run(frames, audio)
    logp = net.gru_head(net.stcnn(frames))
    ids, lens = A.ctc_greedy_decode(logp)
torch.cuda.synchronize()
print("ok", float(scores.sum()), int(lens.sum()), "total launches", L.avs_launch_count())
