"""Config 5 on real GPUs (run under torchrun, NCCL): one data-parallel detector training step
(hidden 512, batch 64 per rank, BCE-with-logits + Adam(lr 1e-3, L2 1e-5), flat 28.4 MB gradient
all-reduce) must equal the single-process step on the concatenated batch; prints the step time."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsync_b200 as A

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
torch.manual_seed(7)
model = A.MisalignmentDetector(13864, 512, dropout=0.0).to(dev)
ref = A.MisalignmentDetector(13864, 512, dropout=0.0).to(dev)
ref.load_state_dict(model.state_dict())
g = torch.Generator().manual_seed(11)
x = torch.randn((64 * world, 13864), generator=g).to(dev)
y = (torch.rand((64 * world,), generator=g) > 0.5).float().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-5)
sl = slice(rank * 64, rank * 64 + 64)
loss = A.distributed.ddp_detector_step(model, x[sl], y[sl], opt)
ref.train()
l_ref = torch.nn.BCEWithLogitsLoss()(ref(x), y)
opt_ref.zero_grad()
l_ref.backward()
opt_ref.step()
# gradients must agree to fp32 summation-order noise; parameters only to ~lr * that noise amplified by
# Adam's g / sqrt(g^2) normalisation of near-zero gradients (first step), hence the looser bound
gerr = max((p.grad - q.grad).abs().max().item() for p, q in zip(model.parameters(), ref.parameters()))
err = max((p - q).abs().max().item() for p, q in zip(model.parameters(), ref.parameters()))
assert abs(loss.item() - l_ref.item()) < 1e-5 and gerr < 1e-6 and err < 2e-4, (loss.item(), l_ref.item(), gerr, err)
for _ in range(5):
    A.distributed.ddp_detector_step(model, x[sl], y[sl], opt)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    A.distributed.ddp_detector_step(model, x[sl], y[sl], opt)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    n = sum(p.numel() for p in model.parameters())
    print(f"ddp step ok: world={world} params={n} max|grad - single-process|={gerr:.2e} max|param diff|={err:.2e} "
          f"loss {loss.item():.6f} vs {l_ref.item():.6f}; {t.item():.3f} ms/step (batch {64 * world}, max over ranks)")
dist.destroy_process_group()
