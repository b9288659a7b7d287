"""Tiny driver for ncu: two STCNN forwards (bf16) on a few clips, so `-k 'regex:conv_umma|conv_l2_fused' -s 3 -c 3`
captures the three conv launches of the second forward.  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

B = int(os.environ.get("PROF_CLIPS", "8"))
torch.manual_seed(0)
net = A.LipNet(39, precision=os.environ.get("PROF_PRECISION", "bf16")).cuda().eval()
frames = torch.rand((B, 1, 75, 50, 100), generator=torch.Generator().manual_seed(3)).cuda()
for _ in range(2):
    emb = net.stcnn(frames)
torch.cuda.synchronize()
print("ok", float(emb.sum()))
