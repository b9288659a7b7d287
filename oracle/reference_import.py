"""Import the UNMODIFIED reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE (oracle).  Used by ``oracle/make_golden.py`` to mint the
fixtures in ``tests/golden/`` and by ``tests/test_oracle.py`` (skipped when the
mount is absent, e.g. on the GPU box).  ``misalignment_detection_train.py``
imports ``librosa`` (:16) and ``matplotlib.pyplot`` (:17), neither installed and
no network: stub modules are injected into ``sys.modules`` first, with
``librosa.feature.mfcc`` bound to the restatement in ``oracle/mfcc_ref.py``.
Nothing here is copied into the product.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AVS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _install_stubs() -> None:
    from . import mfcc_ref
    if "librosa" not in sys.modules:
        librosa = types.ModuleType("librosa")
        feature = types.ModuleType("librosa.feature")
        feature.mfcc = lambda y=None, sr=22050, n_mfcc=20, hop_length=512, **kw: mfcc_ref.mfcc(
            y, sr=sr, n_mfcc=n_mfcc, hop_length=hop_length)
        librosa.feature = feature

        def _no(*a, **k):
            raise RuntimeError("librosa stub: only feature.mfcc is restated")
        librosa.load = _no
        librosa.resample = _no
        sys.modules["librosa"] = librosa
        sys.modules["librosa.feature"] = feature
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load():
    """Returns (model, utils, dataset, misalignment_detection_train) reference modules."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    saved = {k: sys.modules.get(k) for k in ("model", "utils", "dataset", "misalignment_detection_train")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        mods = tuple(importlib.import_module(k) for k in saved)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():           # do not leave 'model'/'utils' shadowing anything
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return mods
