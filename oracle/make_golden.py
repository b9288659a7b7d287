"""Mint tests/golden/*.npz by running the UNMODIFIED reference (this container only).

    python -m oracle.make_golden

Inputs and weights are regenerated from seeds (``sweep_ref.synth_*``,
``lipnet_ref.init_lipnet_state``, ``sweep_ref.init_detector_state``); each
fixture stores a CRC of the regenerated inputs so a silent RNG change is caught.
Outputs stored are those of the reference's own functions:
  stcnn.npz   - extract_visual_embeddings (strided sample of emb + full vstats)
  lipnet.npz  - LipNet.forward log-probs + utils.decode_prediction texts
  decode.npz  - decode_prediction on hand-built edge cases
  astats.npz  - compute_audio_stats(shift_audio(.)) for k in [-20,20], 4 audio kinds
                (librosa.feature.mfcc is the stubbed restatement: parity unpinned there)
  sweep.npz   - full reference-style sweep scores / argmax / top-2 margin
"""
from __future__ import annotations

import os
import zlib

import numpy as np
import torch

from . import lipnet_ref, reference_import, sweep_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EMB_STRIDE = 61          # coprime with 6912 and 75*6912: samples every feature/time combination class
AUDIO_KINDS = ("noise", "halfsilent", "chirp", "speechlike")


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(np.asarray(a)).tobytes())


class _Vocab:
    idx_to_char = lipnet_ref.make_vocab()


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    model, utils, dataset, mdt = reference_import.load()
    torch.set_num_threads(os.cpu_count() or 1)

    # reference modules with the seeded weights
    torch.manual_seed(0)
    ref_lipnet = model.LipNet(vocab_size=39).eval()
    sd = lipnet_ref.init_lipnet_state(39, 256, seed=0)
    for k, v in ref_lipnet.state_dict().items():
        assert torch.equal(v, sd[k]), f"seeded init differs from reference LipNet at {k}"
    det_sd = sweep_ref.init_detector_state(13864, 512, seed=1)
    ref_det = mdt.MisalignmentDetector(13864, 512).eval()
    ref_det.load_state_dict(det_sd)
    ref_vocab = dataset.GridDataset._create_vocab(_Vocab())        # checks our table too
    assert _Vocab.idx_to_char == lipnet_ref.make_vocab() and len(ref_vocab) == 39

    frames = sweep_ref.synth_frames(2, seed=1234)
    with torch.no_grad():
        emb = mdt.extract_visual_embeddings(ref_lipnet, frames)           # [2,75,6912]
        vstats = torch.stack([torch.cat([e.mean(dim=0), e.std(dim=0)]) for e in emb])
        logp = ref_lipnet(frames)
    texts = [utils.decode_prediction(logp[i], _Vocab) for i in range(2)]
    flat = emb.reshape(2, -1).numpy()
    np.savez_compressed(os.path.join(OUT, "stcnn.npz"), frames_crc=crc(frames.numpy()),
                        emb_stride=EMB_STRIDE, emb_sample=flat[:, ::EMB_STRIDE].copy(),
                        emb_sum=flat.astype(np.float64).sum(1), emb_abs_sum=np.abs(flat).astype(np.float64).sum(1),
                        vstats=vstats.numpy())
    np.savez_compressed(os.path.join(OUT, "lipnet.npz"), frames_crc=crc(frames.numpy()),
                        logp=logp.numpy(), texts=np.array(texts))

    # decode edge cases (utils.py:8-36): rows are one-hot-ish log-prob tables
    def table(ids, V=39, T=None):
        t = np.full((len(ids), V), -5.0, dtype=np.float32)
        for i, c in enumerate(ids):
            t[i, c] = -0.1
        return t
    cases = {
        "all_blank": table([0] * 75),
        "repeat_blank_repeat": table([1, 1, 0, 1, 0, 0, 2, 2, 3]),       # a,blank,a -> "aab c"...
        "pad_and_space": table([38, 38, 37, 0, 37, 36, 1]),
        "no_blank_runs": table([5, 5, 5, 6, 6, 5, 5]),
        "single": table([7]),
        "ties_first_index": np.zeros((6, 39), dtype=np.float32),          # all equal -> argmax 0 (blank)
    }
    tie = np.full((5, 39), -3.0, dtype=np.float32)
    tie[:, 4] = -1.0
    tie[:, 9] = -1.0                                                       # tie between 4 and 9 -> 4
    tie[2, 0] = -0.5
    cases["ties_two_way"] = tie
    rng = np.random.default_rng(7)
    cases["random_uniformish"] = rng.normal(-3.66, 0.02, (75, 39)).astype(np.float32)
    dec = {}
    for name, t in cases.items():
        dec[f"{name}__in"] = t
        dec[f"{name}__text"] = np.array(utils.decode_prediction(torch.from_numpy(t), _Vocab))
        dec[f"{name}__ids"] = np.array(lipnet_ref.greedy_ids(t), dtype=np.int32)
        assert lipnet_ref.decode_prediction(t) == str(dec[f"{name}__text"])
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **dec)

    # audio stats for every shift, reference functions (librosa stubbed)
    shifts = np.arange(-20, 21)
    ast = {"shifts": shifts}
    for kind in AUDIO_KINDS:
        a = sweep_ref.synth_audio(1, seed=1234, kind=kind)[0]
        ast[f"{kind}__crc"] = crc(a)
        ast[kind] = np.stack([
            mdt.compute_audio_stats(mdt.shift_audio(a, int(k), 25.0, 16000), 16000, 20).numpy() for k in shifts])
    np.savez_compressed(os.path.join(OUT, "astats.npz"), **ast)

    # full sweep, reference call pattern (FeatureExtractor.build_feature order, :199-208)
    audio = np.stack([sweep_ref.synth_audio(1, seed=1234, kind="noise")[0],
                      sweep_ref.synth_audio(1, seed=1235, kind="speechlike")[0]])
    scores = np.zeros((2, 41), dtype=np.float32)
    with torch.no_grad():
        for i in range(2):
            for j, k in enumerate(shifts):
                a_k = mdt.compute_audio_stats(mdt.shift_audio(audio[i], int(k), 25.0, 16000), 16000, 20)
                feat = torch.cat([vstats[i], a_k], dim=0)
                scores[i, j] = torch.sigmoid(ref_det(feat.unsqueeze(0))).item()
    srt = np.sort(scores, axis=1)
    np.savez_compressed(os.path.join(OUT, "sweep.npz"), shifts=shifts, audio_crc=crc(audio),
                        frames_crc=crc(frames.numpy()), scores=scores, best=scores.argmax(1).astype(np.int32),
                        margin=srt[:, -1] - srt[:, -2])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("texts", texts, "best", scores.argmax(1), "margin", srt[:, -1] - srt[:, -2])


if __name__ == "__main__":
    main()
