"""CPU restatement of the text metrics around the decode: CER / WER (train.py:945-993, the reference's
``calculate_cer`` / ``calculate_wer``) and evaluate_model's positional character accuracy (utils.py:83-86).
TEST INFRASTRUCTURE (oracle).  ``train.py`` imports TensorFlow and cannot be imported here, so these follow
its definitions: Levenshtein distance with unit costs over characters (resp. whitespace-split words),
divided by the target length; an empty target scores 1.0 against a non-empty prediction and 0.0 otherwise."""
from __future__ import annotations

from typing import Sequence


def levenshtein(a: Sequence, b: Sequence) -> int:
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(prev[j - 1] if x == y else 1 + min(prev[j], cur[j - 1], prev[j - 1]))
        prev = cur
    return prev[-1]


def cer(prediction: str, target: str) -> float:
    if not target:
        return 1.0 if prediction else 0.0
    return levenshtein(prediction, target) / len(target)


def wer(prediction: str, target: str) -> float:
    p, t = prediction.split(), target.split()
    if not t:
        return 1.0 if p else 0.0
    return levenshtein(p, t) / len(t)


def char_accuracy(true_text: str, predicted_text: str) -> float:
    return sum(1 for a, b in zip(true_text, predicted_text) if a == b) / max(len(true_text), 1) * 100
