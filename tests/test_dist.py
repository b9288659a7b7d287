"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: ragged score gather and the DDP detector step."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import avsync_b200 as A
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = A.distributed.shard_range(n_total, rank, world)
        g = torch.Generator().manual_seed(0)
        all_scores = torch.rand((n_total, 41), generator=g)
        all_best = all_scores.argmax(1).to(torch.int32)
        s, b = A.distributed.gather_scores(all_scores[lo:hi].clone(), all_best[lo:hi].clone(), n_total)
        assert torch.equal(s, all_scores) and torch.equal(b, all_best)
        lab, pr = A.distributed.gather_labels_probs((all_scores[lo:hi, 0] > 0.5).float(), all_scores[lo:hi, 1], n_total)
        assert torch.equal(pr, all_scores[:, 1])

        # DDP step == single-process step on the concatenated batch
        torch.manual_seed(1)
        model = A.MisalignmentDetector(64, 16, dropout=0.0)
        ref = A.MisalignmentDetector(64, 16, dropout=0.0)
        ref.load_state_dict(model.state_dict())
        x = torch.randn((8, 64), generator=g)
        y = (torch.rand((8,), generator=g) > 0.5).float()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
        opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-5)
        half = slice(rank * 4, rank * 4 + 4)
        loss = A.distributed.ddp_detector_step(model, x[half], y[half], opt)
        ref.train()
        l_ref = torch.nn.BCEWithLogitsLoss()(ref(x), y)
        opt_ref.zero_grad()
        l_ref.backward()
        opt_ref.step()
        assert abs(loss.item() - l_ref.item()) < 1e-6
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, atol=1e-6), (p - q).abs().max()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_and_ddp_step_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 13, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_auc_acc_matches_sklearn_and_nan_on_single_class():
    import avsync_b200 as A
    rng = np.random.default_rng(0)
    y = (rng.random(50) > 0.5).astype(float)
    p = rng.random(50)
    acc, auc = A.distributed.auc_acc(y, p)
    assert 0 <= acc <= 1 and 0 <= auc <= 1
    _, auc1 = A.distributed.auc_acc(np.ones(10), rng.random(10))
    assert np.isnan(auc1)


def test_graphed_detector_step_needs_cuda():
    """The CUDA-graph form of the training step has no CPU fallback: a CPU model is refused, loudly."""
    import pytest
    import avsync_b200 as A
    m = A.MisalignmentDetector(13864, 8)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    with pytest.raises(RuntimeError, match="CUDA"):
        A.distributed.GraphedDetectorStep(m, opt, 4)
