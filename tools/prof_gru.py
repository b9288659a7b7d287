"""ncu driver: two GRU-head forwards on K3_CLIPS clips (second one is profiled with -k regex:gru_cluster -s 2 -c 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

B = int(os.environ.get("K3_CLIPS", "256"))
torch.manual_seed(0)
net = A.LipNet(39, precision="bf16x3").cuda().eval()
emb = torch.rand((B, 75, 6912), generator=torch.Generator().manual_seed(1)).cuda() * 0.2
for _ in range(2):
    out = net.gru_head(emb)
torch.cuda.synchronize()
print("ok", float(out.sum()))
