"""Where does the MMA-issuing warp of the tcgen05 conv kernels spend its time?  (avs_debug_set bit 16: per-CTA
clock64 split of the issuer loop printed by blocks 0 and 77.)  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

A._native.use_experiments_build()   # libavsync_b200_exp.so: avs_debug_set + AVS_* knobs (make -C csrc EXPERIMENTS=1)
L = A._native.lib()
B = int(os.environ.get("MB_CLIPS", "64"))
frames = torch.rand((B, 1, 75, 50, 100), generator=torch.Generator().manual_seed(3)).cuda()
FLAGS = [int(f) for f in os.environ.get("SPLIT_FLAGS", "16,23").split(",")]
for prec in os.environ.get("SPLIT_PREC", "bf16,bf16x3").split(","):
    torch.manual_seed(0)
    net = A.LipNet(39, precision=prec).cuda().eval()
    net.stcnn(frames)
    torch.cuda.synchronize()
    for flags in FLAGS:
        print(f"== {prec} dbg={flags}", flush=True)
        L.avs_debug_set(flags)
        net.stcnn(frames)
        torch.cuda.synchronize()
        L.avs_debug_set(0)
