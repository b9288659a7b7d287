"""Host-side logic of the drop-in package (CPU only): reference-compatible names, shapes, state_dict keys,
checkpoint formats, shift arithmetic, shard arithmetic."""
import os

import numpy as np
import pytest
import torch

import avsync_b200 as A
from oracle import lipnet_ref, sweep_ref


def test_shift_audio_matches_oracle_everywhere():
    rng = np.random.default_rng(0)
    a = rng.normal(size=997).astype(np.float32)
    for fps, sr in ((25.0, 16000), (29.97, 16000), (25.0, 8000), (0.0, 16000)):
        for k in (-40, -20, -3, -1, 0, 1, 2, 7, 20, 40):
            got = A.shift_audio(a, k, fps, sr)
            want = sweep_ref.shift_audio(a, k, fps, sr)
            assert np.array_equal(got, want), (fps, sr, k)
            assert got is not a
            assert A.shift_samples(k, fps, sr) == sweep_ref.shift_samples(k, fps, sr)
    assert A.shift_audio(np.zeros(0, np.float32), 3, 25.0, 16000).size == 0


def test_compute_audio_stats_empty_input_guard():
    # reference returns zeros before touching librosa (misalignment_detection_train.py:118-119)
    assert A.compute_audio_stats(np.zeros(0, np.float32), 16000, 20).tolist() == [0.0] * 40


def test_lipnet_state_dict_is_reference_compatible(lipnet_sd):
    net = A.LipNet(vocab_size=39)
    assert list(net.state_dict().keys()) == list(lipnet_sd.keys())
    for k, v in net.state_dict().items():
        assert v.shape == lipnet_sd[k].shape, k
    net.load_state_dict(lipnet_sd)                     # bare form
    assert net.conv_output_dim == 6912 and sum(p.numel() for p in net.parameters()) == 12537927
    torch.manual_seed(0)
    fresh = A.LipNet(vocab_size=39)                    # same default init as the reference under the same seed
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, lipnet_sd[k]), k
    for attr in ("conv1", "conv2", "conv3", "pool1", "pool2", "pool3", "dropout1", "dropout2", "dropout3",
                 "gru1", "gru2", "fc"):
        assert hasattr(net, attr)


def test_load_lipnet_accepts_both_checkpoint_forms(tmp_path, lipnet_sd):
    p1, p2 = tmp_path / "bare.pth", tmp_path / "wrapped.pth"
    torch.save(lipnet_sd, p1)
    torch.save({"epoch": 3, "model_state_dict": lipnet_sd, "train_loss": 0.0}, p2)
    for p in (p1, p2):
        net = A.load_lipnet(str(p), 39, torch.device("cpu"))
        assert not net.training and all(not q.requires_grad for q in net.parameters())
        assert torch.equal(net.fc.weight, lipnet_sd["fc.weight"])


def test_detector_module_and_checkpoint_format(tmp_path, det_sd):
    det = A.MisalignmentDetector(13864, 512)
    assert list(det.state_dict().keys()) == list(det_sd.keys())
    det.load_state_dict(det_sd)
    det.eval()
    x = torch.randn(5, 13864)
    with torch.no_grad():
        np.testing.assert_allclose(det(x).numpy(), sweep_ref.detector_logits(det_sd, x).numpy(), rtol=1e-5, atol=1e-6)
    path = str(tmp_path / "det.pth")
    A.save_detector(det, path, A.DetectorConfig(max_shift_frames=20))
    ck = torch.load(path)
    assert set(ck) == {"model_state_dict", "input_dim", "hidden_dim", "config"}
    assert ck["config"] == {"sample_rate": 16000, "n_mfcc": 20, "max_shift_frames": 20}
    back = A.load_detector(path, torch.device("cpu"))
    assert back.hidden_dim == 512 and torch.equal(back.classifier[0].weight, det.classifier[0].weight)


def test_detector_config_defaults():
    c = A.DetectorConfig()
    assert (c.img_width, c.img_height, c.max_video_length, c.sample_rate, c.n_mfcc, c.max_shift_frames,
            c.num_negative_samples, c.default_fps) == (100, 50, 75, 16000, 20, 10, 1, 25.0)


def test_ids_to_text_matches_reference_table():
    class DS:
        idx_to_char = lipnet_ref.make_vocab()
    assert A.utils.ids_to_text([1, 1, 37, 38, 27, 99], DS) == "aa <pad>0"


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 8192, 8191):
        for w in (1, 2, 3, 8):
            spans = [A.distributed.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_feature_extractor_errors_like_reference():
    class Grid:
        def process_video(self, p):
            return torch.zeros(1, 75, 50, 100)
    fx = A.FeatureExtractor(Grid(), A.LipNet(39).eval(), torch.device("cpu"), A.DetectorConfig())
    with pytest.raises(RuntimeError, match="Failed to load audio"):
        fx._load_audio("nope.mpg")


def _describe_plan(n, sr, shifts):
    import ctypes
    N = A._native
    arr = (ctypes.c_int32 * len(shifts))(*shifts)
    nf, nu = ctypes.c_int(), ctypes.c_int()
    N.check(N.lib().avs_mfcc_plan_describe(n, sr, arr, len(shifts), ctypes.byref(nf), ctypes.byref(nu), None, None))
    frames = (ctypes.c_int32 * (3 * nu.value))()
    fmap = (ctypes.c_int32 * (len(shifts) * nf.value))()
    N.check(N.lib().avs_mfcc_plan_describe(n, sr, arr, len(shifts), ctypes.byref(nf), ctypes.byref(nu), frames, fmap))
    return nf.value, np.array(frames).reshape(-1, 3), np.array(fmap).reshape(len(shifts), nf.value)


def test_mfcc_frame_plan_dedup_is_exact_on_host():
    """The K1 host plan (no GPU needed): frames rebuilt from the unique (start, lo, hi) table through the
    map must equal, sample for sample, the centre-padded frames of the np-shifted signal."""
    rng = np.random.default_rng(0)
    for (n, sr, fps, S) in ((48000, 16000, 25.0, 20), (48000, 16000, 29.97, 6), (5000, 8000, 25.0, 9), (900, 16000, 25.0, 2)):
        x = rng.normal(size=n).astype(np.float32)
        ks = list(range(-S, S + 1))
        shifts = [sweep_ref.shift_samples(k, fps, sr) for k in ks]
        nf, frames, fmap = _describe_plan(n, sr, shifts)
        hop = max(1, sr // 40)
        assert nf == 1 + n // hop
        xz = np.concatenate([np.zeros(4096 + n, np.float32), x, np.zeros(4096 + n, np.float32)])   # x[i] at xz[4096+n+i]
        for j, k in enumerate(ks):
            y = np.pad(sweep_ref.shift_audio(x, k, fps, sr), (1024, 1024))
            for f in (0, 1, 2, 3, nf // 2, nf - 4, nf - 3, nf - 2, nf - 1):
                if not 0 <= f < nf:
                    continue
                start, lo, hi = frames[fmap[j, f]]
                idx = np.arange(start, start + 2048)
                got = np.where((idx >= lo) & (idx < hi), xz[4096 + n + idx], 0.0)
                assert np.array_equal(got, y[hop * f: hop * f + 2048]), (n, sr, fps, k, f)
    nf, frames, fmap = _describe_plan(48000, 16000, [640 * k for k in range(-20, 21)])
    assert nf == 121 and len(frames) == 746            # 41 x 121 = 4961 (shift, frame) pairs -> 746 unique frames
    nf, frames, _ = _describe_plan(48000, 16000, [640 * k for k in range(-15, 16)])
    assert len(frames) < 31 * 121 / 5
    _, frames, fmap = _describe_plan(1000, 16000, [5000, -5000, 0])      # |shift| >= len: all-zero frames
    assert (fmap[0] == fmap[1]).all() and len(set(fmap[0])) == 1


class _FakeExtractor:
    """Stands in for FeatureExtractor on the CPU: records (path, shift) and returns a feature that encodes them."""

    def __init__(self):
        self.calls = []

    def build_feature(self, path, shift):
        self.calls.append((path, shift))
        return torch.full((4,), float(shift)), {"video_path": path, "shift_frames": shift, "fps": 25.0}

    def build_features_sweep(self, path, S):
        self.calls.append((path, "sweep", S))
        return torch.arange(-S, S + 1, dtype=torch.float32)[:, None].expand(-1, 4).clone()


def test_misalignment_dataset_matches_reference_rng_and_items():
    from oracle import reference_import
    cfg = A.DetectorConfig(max_shift_frames=20, num_negative_samples=2)
    paths = [f"clip{i}.npy" for i in range(7)]
    fx = _FakeExtractor()
    ds = A.MisalignmentDataset(paths, fx, cfg, seed=42)
    assert len(ds) == 21
    items = [ds[i] for i in range(len(ds))]
    labels = [float(l) for _, l in items]
    assert labels == [1.0, 0.0, 0.0] * 7
    shifts = [s for _, s in fx.calls]
    assert all(s == 0 for s in shifts[0::3]) and all(1 <= abs(s) <= 20 for i, s in enumerate(shifts) if i % 3)
    # sweep-aware variant: same RNG stream -> same shifts, features served from one table per clip
    fx2 = _FakeExtractor()
    ds2 = A.MisalignmentDataset(paths, fx2, cfg, seed=42, precompute=True)
    got = [float(ds2[i][0][0]) for i in range(len(ds2))]
    assert got == [float(s) for s in shifts]
    assert len(fx2.calls) == len(paths) and all(c[1] == "sweep" for c in fx2.calls)
    if reference_import.available():
        _, _, _, mdt = reference_import.load()
        fx3 = _FakeExtractor()
        ref_cfg = mdt.DetectorConfig(max_shift_frames=20, num_negative_samples=2)
        rds = mdt.MisalignmentDataset(paths, fx3, ref_cfg, seed=42)
        ref_items = [rds[i] for i in range(len(rds))]
        assert fx3.calls == fx.calls
        assert [float(l) for _, l in ref_items] == labels


def test_run_epoch_matches_reference_semantics():
    from oracle import reference_import
    torch.manual_seed(0)
    x = torch.randn(40, 16)
    y = (torch.rand(40) > 0.5).float()
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=8)
    model = A.MisalignmentDetector(16, 8, dropout=0.0)
    crit = torch.nn.BCEWithLogitsLoss()
    ev = A.run_epoch(model, loader, crit, torch.device("cpu"))
    assert set(ev) == {"loss", "acc", "auc", "labels", "probs"} and ev["probs"].shape == (40,)
    if reference_import.available():
        _, _, _, mdt = reference_import.load()
        ref_model = mdt.MisalignmentDetector(16, 8, dropout=0.0)
        ref_model.load_state_dict(model.state_dict())
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
        ropt = torch.optim.Adam(ref_model.parameters(), lr=1e-3, weight_decay=1e-5)
        a = A.run_epoch(model, loader, crit, torch.device("cpu"), opt)
        b = mdt.run_epoch(ref_model, loader, crit, torch.device("cpu"), ropt)
        assert abs(a["loss"] - b["loss"]) < 1e-6 and a["acc"] == b["acc"] and abs(a["auc"] - b["auc"]) < 1e-9
        for p, q in zip(model.parameters(), ref_model.parameters()):
            assert torch.allclose(p, q, atol=1e-7)
    one = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, torch.ones(40)), batch_size=8)
    assert np.isnan(A.run_epoch(model, one, crit, torch.device("cpu"))["auc"])


def test_conv_work_partition_covers_every_item_once_and_is_balanced():
    """The persistent conv kernels cut their work items into contiguous, cost-balanced spans (one per CTA).  The host
    mirror of that device code needs no GPU: every item in exactly one span, spans in order, cost within one item."""
    import ctypes
    L = A._native.lib()
    # (tiles per plane, tiles per item) of the six layer kinds, plus awkward sizes
    for n_tiles, nt in ((20, 4), (5, 2), (2, 2), (20, 2), (5, 1), (2, 1), (7, 3)):
        n_ts = -(-n_tiles // nt)
        for n_clips, steps, ctas in ((1, 75, 148), (3, 75, 148), (64, 75, 148), (5, 75, 7), (2, 4, 148), (1, 1, 1)):
            n_items = n_clips * steps * n_ts
            grid = min(n_items, ctas)
            tiles = lambda it: min(nt, n_tiles - ((it // steps) % n_ts) * nt)
            cost = lambda it: tiles(it) * (5 if tiles(it) == nt else 7)   # a partial tile set costs 1.4x per tile (measured)
            prev_last, costs = 0, []
            for c in range(grid):
                f, l = ctypes.c_int(), ctypes.c_int()
                assert L.avs_conv_item_span(n_clips, steps, n_tiles, nt, grid, c, ctypes.byref(f), ctypes.byref(l)) == 0
                assert f.value == prev_last and l.value >= f.value, (n_tiles, nt, n_clips, steps, c)
                prev_last = l.value
                costs.append(sum(cost(i) for i in range(f.value, l.value)))
            assert prev_last == n_items
            assert sum(costs) == sum(cost(i) for i in range(n_items))
            assert max(costs) - min(costs) <= 2 * 5 * nt, (n_tiles, nt, n_clips, steps, ctas, max(costs), min(costs))
    f, l = ctypes.c_int(), ctypes.c_int()
    assert L.avs_conv_item_span(1, 75, 5, 2, 4, 4, ctypes.byref(f), ctypes.byref(l)) != 0      # cta out of range


def test_feature_extractor_keeps_the_reference_constructor_and_audio_loading(tmp_path):
    """FeatureExtractor(grid, lipnet, device, cfg) — the reference's four arguments (:148) — must construct and load
    audio by itself: librosa -> moviepy like the reference, scipy for .wav when neither is installed; failures raise
    the reference's RuntimeError (:191)."""
    from scipy.io import wavfile

    class Grid:
        def process_video(self, path):
            return torch.zeros((1, 75, 50, 100))
    fx = A.FeatureExtractor(Grid(), A.LipNet(39).eval(), torch.device("cpu"), A.DetectorConfig())
    assert fx.audio_loader is A.default_audio_loader and fx.visual_cache == {} and fx.audio_cache == {} and fx.fps_cache == {}
    rng = np.random.default_rng(1)
    pcm = (rng.normal(0, 0.1, (4410, 2)).clip(-1, 1) * 32767).astype(np.int16)
    wav = str(tmp_path / "clip.wav")
    wavfile.write(wav, 44100, pcm)
    audio, sr = fx._load_audio(wav)
    assert sr == 44100 and audio.dtype == np.float32 and audio.shape == (4410,)
    np.testing.assert_allclose(audio, pcm.astype(np.float32).mean(axis=1) / 32768.0, atol=1e-6)
    assert fx._load_audio(wav)[0] is audio                                   # cached per path
    with pytest.raises(RuntimeError, match="Failed to load audio from"):
        fx._load_audio(str(tmp_path / "missing.mpg"))
    assert A.get_video_fps("clip.npy", 30.0) == 30.0 and A.get_video_fps(str(tmp_path / "missing.mpg")) == 25.0
