"""CPU oracle for the AV sync-scoring hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy / torch-CPU) of the
reference algorithm, used as the *checker* by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  The product package never imports it and has no CPU
fallback: it raises if the CUDA extension is missing.

Parity pinning (see DESIGN.md section "Oracle"):
  * ``lipnet_ref`` / ``sweep_ref`` restate ``model.py``, ``utils.py`` and
    ``misalignment_detection_train.py``; they are pinned against outputs of the
    reference code itself, imported unmodified from ``/root/reference`` by
    ``oracle/make_golden.py`` and committed as ``tests/golden/*.npz``.
  * ``mfcc_ref`` restates ``librosa.feature.mfcc`` (librosa >= 0.10 semantics).
    librosa is NOT vendored, pinned or installed: that one boundary is
    **parity unpinned** against librosa itself; it is cross-checked against
    ``torchaudio.transforms.MFCC`` (independent implementation, same options).
"""
