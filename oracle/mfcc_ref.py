"""numpy/scipy restatement of ``librosa.feature.mfcc`` as called by the reference.

TEST INFRASTRUCTURE (oracle).  PARITY UNPINNED against librosa itself: librosa
is an un-vendored, un-pinned dependency of the reference (only named in
``README.md:66``; call site ``misalignment_detection_train.py:120-121``) and is
not installed here.  This file restates the published librosa >= 0.10 call
chain ``feature.mfcc -> feature.melspectrogram -> core.spectrum._spectrogram ->
core.stft`` + ``filters.mel`` + ``core.power_to_db`` + ``scipy.fftpack.dct``
with the defaults the reference relies on, and is cross-checked against
``torchaudio.transforms.MFCC`` in ``tests/test_oracle.py``.

Defaults used by the call site (``y`` float32, ``sr=16000``, ``n_mfcc=20``,
``hop_length=400``): n_fft=2048, win_length=2048, periodic Hann, center=True,
pad_mode="constant", power=2.0, n_mels=128, fmin=0, fmax=sr/2, Slaney mel scale
(htk=False) with Slaney area normalisation, power_to_db(ref=1.0, amin=1e-10,
top_db=80.0) with the max over the whole [n_mels, n_frames] array, DCT-II
``norm="ortho"`` along the mel axis, first ``n_mfcc`` rows, no liftering.
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.fft
import scipy.fftpack

N_FFT = 2048
N_MELS = 128
AMIN = 1e-10
TOP_DB = 80.0


def _hz_to_mel(f):
    """librosa.core.convert.hz_to_mel(htk=False) (Slaney Auditory Toolbox scale)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        big = f >= min_log_hz
        mels[big] = min_log_mel + np.log(f[big] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    """librosa.core.convert.mel_to_hz(htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    big = m >= min_log_mel
    freqs[big] = min_log_hz * np.exp(logstep * (m[big] - min_log_mel))
    return freqs


@functools.lru_cache(maxsize=8)
def mel_filterbank(sr: int = 16000, n_fft: int = N_FFT, n_mels: int = N_MELS,
                   fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney',
    dtype=float32) -> float32 [n_mels, 1 + n_fft//2]."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    min_mel = _hz_to_mel(fmin)
    max_mel = _hz_to_mel(fmax)
    mel_f = _mel_to_hz(np.linspace(min_mel, max_mel, n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    weights.setflags(write=False)
    return weights


@functools.lru_cache(maxsize=4)
def hann_window(n_fft: int = N_FFT) -> np.ndarray:
    """scipy.signal.get_window('hann', n_fft, fftbins=True): periodic Hann, float64."""
    n = np.arange(n_fft, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)
    w.setflags(write=False)
    return w


def stft_power(y: np.ndarray, hop_length: int, n_fft: int = N_FFT) -> np.ndarray:
    """|librosa.stft(y, n_fft, hop_length, center=True, pad_mode='constant')|**2.

    librosa multiplies the float32 frames by the float64 window, runs the real
    FFT in double precision and stores the result as complex64; ``np.abs`` then
    ``**2`` are taken in float32.  Returns float32 [1 + n_fft//2, n_frames].
    """
    y = np.asarray(y, dtype=np.float32)
    pad = n_fft // 2
    yp = np.pad(y, (pad, pad), mode="constant")
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = yp[idx]                                     # [n_fft, n_frames] float32
    spec = scipy.fft.rfft(hann_window(n_fft)[:, None] * frames, axis=0)
    spec = spec.astype(np.complex64)
    return np.abs(spec) ** 2.0                           # float32


def power_to_db(S: np.ndarray) -> np.ndarray:
    """librosa.power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0)."""
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(np.float32(AMIN), S))
    log_spec = log_spec - np.float32(10.0 * np.log10(max(AMIN, 1.0)))
    return np.maximum(log_spec, log_spec.max() - np.float32(TOP_DB)).astype(np.float32)


def melspectrogram(y: np.ndarray, sr: int, hop_length: int) -> np.ndarray:
    S = stft_power(y, hop_length)
    return mel_filterbank(sr).dot(S).astype(np.float32)   # einsum('ft,mf->mt')


def mfcc(y: np.ndarray, sr: int = 16000, n_mfcc: int = 20, hop_length: int = 512) -> np.ndarray:
    """librosa.feature.mfcc(y=y, sr=sr, n_mfcc=n_mfcc, hop_length=hop_length)
    -> float32 [n_mfcc, n_frames]."""
    S = power_to_db(melspectrogram(y, sr, hop_length))
    M = scipy.fftpack.dct(S, axis=-2, type=2, norm="ortho")[..., :n_mfcc, :]
    return M.astype(np.float32)


@functools.lru_cache(maxsize=4)
def dct_matrix(n_mfcc: int = 20, n_mels: int = N_MELS) -> np.ndarray:
    """Explicit DCT-II ortho basis, float64 [n_mfcc, n_mels] (what fftpack.dct applies)."""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    D = np.cos(np.pi * k * (2 * n + 1) / (2 * n_mels)) * np.sqrt(2.0 / n_mels)
    D[0] *= np.sqrt(0.5)
    return D
