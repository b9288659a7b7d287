"""Per-kernel CUDA-event times of the full sweep step (device-resident, 1 GPU) through the EXPERIMENTS build, so that the
AVS_* environment knobs can be A/B-ed:  AVS_K1_SCHED=0|1 python tools/sweep_kernel_ms.py [clips] [steps] [precision]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A
import bench

# --lib NAME: load libavsync_b200_var_NAME.so (make -C csrc VARIANT=NAME VARIANT_FLAGS=...), or "product"; default: the
# experiments build (AVS_* environment knobs)
args = [a for a in sys.argv[1:]]
LIB = "exp"
if "--lib" in args:
    i = args.index("--lib")
    LIB = args[i + 1]
    del args[i:i + 2]
if LIB == "exp":
    A._native.use_experiments_build()
elif LIB != "product":
    A._native.LIB_PATH = os.path.join(os.path.dirname(A._native.LIB_PATH), f"libavsync_b200_var_{LIB}.so")
L = A._native.lib()
C = int(args[0]) if len(args) > 0 else 1024
STEPS = int(args[1]) if len(args) > 1 else 10
PREC = args[2] if len(args) > 2 else "bf16"
torch.manual_seed(0)
net = A.LipNet(39, precision=PREC).cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
sw = A.SyncSweeper(net, det, 20, 48000, chunk_clips=min(128, C))
fr, au = bench.synth_inputs(C, seed=1000)
fr, au = fr.cuda(), au.cuda()
for _ in range(3):
    res = sw.run(fr, au)
torch.cuda.synchronize()
# checksum of the results, so that variants can be compared for bit identity inside one session
CHK = f"scores {res[0].double().sum().item():.12f} best {int(res[1].long().sum().item())}"
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
print(f"[lib={LIB} {knobs or 'defaults'}] {PREC} {C} clips: {ms:.2f} ms/step = {C / ms * 1e3:.0f} clips/s | " + " | ".join(out) + " | " + CHK, flush=True)
