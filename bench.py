#!/usr/bin/env python
"""bench.py — clips/s of the 41-offset (+-20 frame) AV sync sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--precision bf16|bf16x3|fp32] [--clips-per-gpu C] [--chunk M]

A "step" is one pass of the hot path (K2 STCNN + visual stats | K1 MFCC stats for 41 shifts -> K4
scores + arg-max, plus the cross-rank score gather when N > 1) over one batch of synthetic clips:
C clips per GPU, fixed as N grows (weak scaling; N = 8, C = 1024 is BASELINE config 3's 8192 clips).
Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
host-buffer entry point (pinned host inputs, H2D/D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_FRAMES = 20
N_SHIFTS = 2 * S_FRAMES + 1
N_SAMPLES = 48000
FRAME_ELEMS = 75 * 50 * 100
METRIC = "clips/sec for 41-offset AV sync sweep"
# algorithmic work per clip: MACs of the three conv layers
CONV_FLOP = {1: 2 * 0.900e9, 2: 2 * 14.400e9, 3: 2 * 3.73248e9}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def conv2_traffic(clips_per_launch: int, precision: str):
    """DRAM bytes per launch of the layer-2 conv kernel from the committed `ncu --set full` capture
    (profiles/r01_conv2_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum = 864.9 MB for a
    64-clip bf16 launch = 13.51 MB per clip; algorithmic: 8.6 MB padded input + 2.9 MB pooled output),
    scaled to the clips one bench launch processes."""
    if precision != "bf16":
        return None
    return 13.51e6 * clips_per_launch


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any("Active" == r[5 + j].strip() for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": reasons}


def synth_inputs(n: int, seed: int):
    """Pinned host tensors: frames [n,1,75,50,100] ~ U[0,1), audio [n,48000] ~ N(0,0.1^2) clipped, with a
    per-clip random amplitude envelope so shifted versions differ."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.empty((n, 1, 75, 50, 100), dtype=torch.float32).pin_memory()
    audio = torch.empty((n, N_SAMPLES), dtype=torch.float32).pin_memory()
    for i in range(0, n, 64):
        m = min(64, n - i)
        frames[i:i + m] = torch.rand((m, 1, 75, 50, 100), generator=g)
        env = torch.nn.functional.interpolate(torch.rand((m, 1, 13), generator=g), size=N_SAMPLES, mode="linear",
                                              align_corners=True)[:, 0]
        audio[i:i + m] = (torch.randn((m, N_SAMPLES), generator=g) * 0.1).clamp_(-1, 1) * env * env
    return frames, audio


def cpu_sweep_clips_per_s(n_clips: int, warm: int = 1):
    """The reference's CPU path (oracle port: reference-style loop, B=1 STCNN, one MFCC + one detector call
    per shift) on this box's host cores."""
    from oracle import lipnet_ref, sweep_ref
    torch.set_num_threads(os.cpu_count() or 1)
    sd = lipnet_ref.init_lipnet_state(39, 256, seed=0)
    det = sweep_ref.init_detector_state(13864, 512, seed=1)
    frames = sweep_ref.synth_frames(n_clips + warm, seed=4321)
    audio = sweep_ref.synth_audio(n_clips + warm, seed=4321, kind="speechlike")
    shifts = list(range(-S_FRAMES, S_FRAMES + 1))
    for i in range(warm):
        sweep_ref.sweep_clip(sd, det, frames[i], audio[i], shifts)
    t0 = time.perf_counter()
    for i in range(warm, warm + n_clips):
        sweep_ref.sweep_clip(sd, det, frames[i], audio[i], shifts)
    dt = time.perf_counter() - t0
    return n_clips / dt, dt


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation (oracle port; /root/reference is Python and does
    not exist on the GPU box) timed on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    per_step = args.ref_clips
    for _ in range(args.warmup):
        cpu_sweep_clips_per_s(1, warm=0)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        cpu_sweep_clips_per_s(per_step, warm=0)
        tot += per_step
    dt = time.perf_counter() - t0
    v = tot / dt
    cores = os.cpu_count() or 1
    line = {
        "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"+-{S_FRAMES}-frame (41-offset) sync sweep, reference-style CPU loop, "
                               f"{per_step} clips per step (bounded sample)", "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": f"{tot} clips, torch {torch.get_num_threads()} threads"},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--clips-per-gpu", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--ref-clips", type=int, default=4, help="clips per step of the reference arm")
    ap.add_argument("--cpu-clips", type=int, default=12, help="clips of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.steps < 1 or args.warmup < 3:
        args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import avsync_b200 as A

    torch.cuda.set_device(local_rank)
    A._native.device_check()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    L = A._native.lib()

    # random-init weights of the reference architecture (PyTorch default initialisers, seeded); the
    # product package only — oracle/ is imported by the cpu_baseline leg alone
    torch.manual_seed(0)
    net = A.LipNet(39, precision=args.precision).to(dev).eval()
    torch.manual_seed(1)
    det = A.MisalignmentDetector(13864, 512).to(dev).eval()
    C = args.clips_per_gpu
    n_total = C * world
    sw = A.SyncSweeper(net, det, S_FRAMES, N_SAMPLES, chunk_clips=min(args.chunk, C))
    frames_h, audio_h = synth_inputs(C, seed=1000 + rank)
    frames_d, audio_d = frames_h.to(dev), audio_h.to(dev)

    def step_device():
        s, b = sw.run(frames_d, audio_d)
        return A.distributed.gather_scores(s, b, n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    for _ in range(args.warmup):
        step_device()
    barrier()
    L.avs_prof_reset()
    L.avs_prof_enable(1)
    launches0 = L.avs_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        scores, best = step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    L.avs_prof_enable(0)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    launches = torch.tensor([L.avs_launch_count() - launches0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    value = n_total * args.steps / (ms_total / 1e3)

    # per-kernel times of rank 0 (CUDA events on the launching streams, inside the timed region)
    import ctypes
    prof = {}
    names = ["pack", "conv1", "conv2", "conv3", "vstats", "mfcc_logmel", "mfcc_stats", "score_gemm", "score"]
    for i, nme in enumerate(names):
        t, c = ctypes.c_double(), ctypes.c_int()
        L.avs_prof_read(i, ctypes.byref(t), ctypes.byref(c))
        prof[nme] = {"ms_total": t.value, "launches": c.value}

    # ---------------- end-to-end through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        fh, ah = frames_h.numpy(), audio_h.numpy()

        def step_host():
            s, b = sw.run_host(fh, ah)
            if world > 1:
                return A.distributed.gather_scores(torch.from_numpy(s).to(dev), torch.from_numpy(b).to(dev), n_total)
            return s, b
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s_h, b_h = step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * args.steps / float(dt.item()), "unit": "clips/s",
               "h2d_bytes_per_step": n_total * (FRAME_ELEMS + N_SAMPLES) * 4,
               "d2h_bytes_per_step": n_total * (N_SHIFTS + 1) * 4,
               "timer": "host wall clock around the synchronous host-buffer call, barrier + cuda sync both sides, max over ranks"}
        same = np.array_equal(np.asarray(s_h if world == 1 else s_h.cpu().numpy()), scores.cpu().numpy())
        e2e["matches_device_path"] = bool(same)

    if rank == 0:
        pk, pk_src = peaks()
        conv2 = prof["conv2"]
        clips_per_launch = min(args.chunk, C)
        nl = max(conv2["launches"], 1)
        avg_ms = conv2["ms_total"] / nl
        mult = 3 if args.precision == "bf16x3" else 1
        achieved = (CONV_FLOP[2] * clips_per_launch / (avg_ms / 1e3) / 1e12) if avg_ms > 0 else 0.0
        tensor_peak = pk["bf16_tflops_sustained"]
        roof = {"kernel": "conv_umma_kernel[layer 2: Conv3d 32->64, 3x5x5 + bias + ReLU + pool]" if args.precision != "fp32"
                else "conv_pool_ffma_kernel[layer 2]",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "traffic": conv2_traffic(clips_per_launch, args.precision),
                "peak_source": f"{pk_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "algorithmic_flop_per_launch": CONV_FLOP[2] * clips_per_launch,
                "issued_mma_multiplier": mult, "avg_launch_ms": avg_ms, "launches_timed": conv2["launches"],
                "share_of_step": conv2["ms_total"] / ms_total if ms_total else None,
                # the sustained peak is a measured cuBLAS bf16 GEMM rate, not a hardware bound: a frac near (or a
                # little above) 1 says "as fast as cuBLAS keeps this GPU busy under its power cap"
                "frac_of_burst_peak": achieved / pk["bf16_tflops"] if pk.get("bf16_tflops") else None}
        cpu = None
        if args.cpu_clips > 0:
            v, dt = cpu_sweep_clips_per_s(args.cpu_clips, warm=1)
            cpu = {"value": v, "unit": "clips/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{args.cpu_clips} clips of the same 41-offset sweep, reference-style loop "
                             f"(oracle port, torch {torch.get_num_threads()} threads), {dt:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"+-{S_FRAMES}-frame (41-offset) sync sweep, {C} clips per GPU per step "
                                   f"({n_total} clips/step), GRID-shaped 1x75x50x100 f32 frames + 3 s 16 kHz audio, "
                                   f"random-init LipNet STCNN + detector(hidden 512), chunks of {min(args.chunk, C)} clips",
                       "clips_per_gpu": C, "n_shifts": N_SHIFTS, "precision": args.precision,
                       "parallelism": f"clip-sharded x{world}, NCCL all_gather of [{n_total},41] scores" if world > 1 else "1 GPU",
                       "l2": f"inputs per step {C * (FRAME_ELEMS + N_SAMPLES) * 4 / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)"},
            "e2e": e2e, "gpu_launches": int(launches.item()), "clocks": clocks, "roofline": roof,
            "cpu_baseline": cpu, "kernel_ms": prof,
        }
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at communicator
    creation): point fd 1 at stderr for the life of the process and keep the real stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    _REAL_STDOUT = _json_only_stdout()
    main()
