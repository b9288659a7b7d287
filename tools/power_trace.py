"""Power / clock trace of the sweep step under sustained load (product library): runs the 1024-clip step back to back for
SECONDS seconds while nvidia-smi samples every 20 ms; prints the enforced power limit, the median / max instantaneous
draw, the SM clock distribution and the throttle reasons seen, plus the step time over the whole run.
    python tools/power_trace.py [seconds]"""
import os
import statistics
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A
import bench

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
C = 1024
torch.manual_seed(0)
net = A.LipNet(39, precision="bf16").cuda().eval()
torch.manual_seed(1)
det = A.MisalignmentDetector(13864, 512).cuda().eval()
sw = A.SyncSweeper(net, det, 20, 48000, chunk_clips=128)
fr, au = bench.synth_inputs(C, seed=1000)
fr, au = fr.cuda(), au.cuda()
for _ in range(3):
    sw.run(fr, au)
torch.cuda.synchronize()
q = "power.draw.instant,power.draw.average,clocks.sm,enforced.power.limit,power.max_limit,clocks_event_reasons.sw_power_cap," \
    "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu"
smi = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                       stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 0
while time.perf_counter() - t0 < SECONDS:
    for _ in range(4):
        sw.run(fr, au)
        n += 1
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
smi.terminate()
rows = []
for line in smi.stdout.read().splitlines():
    p = [x.strip() for x in line.split(",")]
    try:
        rows.append((float(p[0]), float(p[1]), float(p[2]), float(p[3]), float(p[4]), p[5], p[6], p[7], float(p[8])))
    except (ValueError, IndexError):
        continue
rows = rows[len(rows) // 5:]   # drop the ramp
inst = [r[0] for r in rows]
clk = [r[2] for r in rows]
print(f"{n} steps of {C} clips in {SECONDS:.0f} s: {ms:.2f} ms/step = {C / ms * 1e3:.0f} clips/s")
print(f"samples {len(rows)} (every 20 ms, first fifth dropped); enforced power limit {rows[0][3]:.0f} W (max {rows[0][4]:.0f} W)")
print(f"instantaneous draw: median {statistics.median(inst):.0f} W, p90 {sorted(inst)[int(0.9 * len(inst))]:.0f} W, max {max(inst):.0f} W; "
      f"1 s average at the end {rows[-1][1]:.0f} W; temperature {rows[-1][8]:.0f} C")
print(f"SM clock: median {statistics.median(clk):.0f} MHz, min {min(clk):.0f}, max {max(clk):.0f} (max boost 1965)")
print("sw_power_cap active in %d %% of the samples, hw_slowdown %d %%, sw_thermal_slowdown %d %%" % (
    100 * sum(r[5] == "Active" for r in rows) // len(rows), 100 * sum(r[6] == "Active" for r in rows) // len(rows),
    100 * sum(r[7] == "Active" for r in rows) // len(rows)))
