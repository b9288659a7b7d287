"""Config 5 alone (eager and CUDA-graph detector training step) on 1..N GPUs:
    python tools/ddp_graph_check.py      |     torchrun --nproc-per-node N tools/ddp_graph_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A
import bench

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
out, ctx = bench.ddp_config5(A, dev, world, rank)
out["cuda_graph"] = bench.ddp_config5_graph(A, dev, world, rank, ctx)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
