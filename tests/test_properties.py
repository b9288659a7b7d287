"""Property-based tests (hypothesis) of the host-side logic and of the oracle's small algorithms.  CPU only."""
import ctypes
import itertools

import numpy as np
from hypothesis import given, settings, strategies as st

import avsync_b200 as A
from oracle import lipnet_ref, metrics_ref, sweep_ref


@settings(max_examples=150, deadline=None)
@given(n=st.integers(0, 400), k=st.integers(-60, 60), fps=st.sampled_from([25.0, 29.97, 30.0, 12.5, 0.0]),
       sr=st.sampled_from([8000, 16000, 22050, 100]))
def test_shift_audio_is_a_zero_filled_delay(n, k, fps, sr):
    a = np.arange(1, n + 1, dtype=np.float32)
    got = A.shift_audio(a, k, fps, sr)
    s = A.shift_samples(k, fps, sr)
    assert got.shape == a.shape and np.array_equal(got, sweep_ref.shift_audio(a, k, fps, sr))
    want = np.zeros_like(a)                              # independent statement of "delay by s, zero fill"
    for i in range(n):
        if 0 <= i - s < n:
            want[i] = a[i - s]
    if s == 0 or abs(s) < n or n == 0:
        assert np.array_equal(got, want)
    else:
        assert not got.any()


@settings(max_examples=100, deadline=None)
@given(n=st.integers(0, 10000), w=st.integers(1, 16))
def test_shard_range_is_a_balanced_partition(n, w):
    spans = [A.distributed.shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=40, deadline=None)
@given(n=st.integers(300, 6000), shifts=st.lists(st.integers(-7000, 7000), min_size=1, max_size=9), seed=st.integers(0, 10))
def test_mfcc_frame_plan_reproduces_shifted_frames(n, shifts, seed):
    """Any integer delays: every (shift, frame) of the plan, rebuilt from its (start, lo, hi) key, equals the
    centre-padded frame of the delayed signal."""
    N = A._native
    arr = (ctypes.c_int32 * len(shifts))(*shifts)
    nf, nu = ctypes.c_int(), ctypes.c_int()
    N.check(N.lib().avs_mfcc_plan_describe(n, 16000, arr, len(shifts), ctypes.byref(nf), ctypes.byref(nu), None, None))
    frames = (ctypes.c_int32 * (3 * nu.value))()
    fmap = (ctypes.c_int32 * (len(shifts) * nf.value))()
    N.check(N.lib().avs_mfcc_plan_describe(n, 16000, arr, len(shifts), ctypes.byref(nf), ctypes.byref(nu), frames, fmap))
    frames = np.array(frames).reshape(-1, 3)
    fmap = np.array(fmap).reshape(len(shifts), nf.value)
    assert nf.value == 1 + n // 400 and nu.value <= len(shifts) * nf.value
    x = np.random.default_rng(seed).normal(size=n).astype(np.float32)
    pad = 8192 + n
    xz = np.concatenate([np.zeros(pad, np.float32), x, np.zeros(pad + 2048, np.float32)])
    for j, s in enumerate(shifts):
        y = np.zeros_like(x)
        if 0 <= s < n:
            y[s:] = x[:n - s]
        elif -n < s < 0:
            y[:n + s] = x[-s:]
        y = np.pad(y, (1024, 1024))
        for f in {0, nf.value // 2, nf.value - 1}:
            start, lo, hi = frames[fmap[j, f]]
            idx = np.arange(start, start + 2048)
            got = np.where((idx >= lo) & (idx < hi), xz[np.clip(pad + idx, 0, len(xz) - 1)], 0.0)
            assert np.array_equal(got, y[400 * f: 400 * f + 2048])


@settings(max_examples=200, deadline=None)
@given(ids=st.lists(st.integers(0, 38), min_size=0, max_size=80))
def test_greedy_collapse_equals_groupby_formulation(ids):
    """utils.py:24-30 == 'collapse runs, then drop blanks'."""
    logp = np.full((max(len(ids), 1), 39), -9.0, dtype=np.float32)
    for t, c in enumerate(ids):
        logp[t, c] = -0.1
    if not ids:
        logp[0, 0] = -0.1
    want = [c for c, _ in itertools.groupby(ids) if c != 0]
    assert lipnet_ref.greedy_ids(logp) == want


@settings(max_examples=200, deadline=None)
@given(a=st.text(alphabet="abc ", max_size=12), b=st.text(alphabet="abc ", max_size=12))
def test_levenshtein_properties(a, b):
    d = metrics_ref.levenshtein(a, b)
    assert d == metrics_ref.levenshtein(b, a) and abs(len(a) - len(b)) <= d <= max(len(a), len(b))
    assert (d == 0) == (a == b)
    assert metrics_ref.levenshtein(a + "c", b + "c") == d
