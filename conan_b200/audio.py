"""Host-side log-mel front-end (SURVEY.md 8f "next" row f1, host part): the arithmetic of the
reference's `utils/audio/__init__.py::librosa_wav2spec` (:36-80) restated with torch.stft and an
own Slaney mel filterbank -- librosa is not in the image, so this is checked against torchaudio's
`melscale_fbanks(norm='slaney', mel_scale='slaney')` in tests, not against librosa itself."""
from __future__ import annotations

import math

import numpy as np
import torch


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sample_rate: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """[n_mels, n_fft//2+1] triangular filters on the Slaney mel scale, area-normalised (librosa default)."""
    freqs = np.linspace(0, sample_rate / 2, n_fft // 2 + 1)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (pts[2:n_mels + 2] - pts[:n_mels]))[:, None]
    return w.astype(np.float32)


def wav2mel(wav: np.ndarray, fft_size=1024, hop_size=320, win_length=1024, num_mels=80, fmin=80, fmax=7600,
            sample_rate=16000, eps=1e-6) -> np.ndarray:
    """wav float [-1,1] -> log10 mel [T, num_mels] (centre-padded STFT, magnitude, Slaney mel, log10 clamp)."""
    x = torch.as_tensor(np.asarray(wav, dtype=np.float32))
    spec = torch.stft(x, n_fft=fft_size, hop_length=hop_size, win_length=win_length,
                      window=torch.hann_window(win_length, periodic=True), center=True, pad_mode="constant",
                      return_complex=True).abs()                                   # [bins, T]
    fmin = 0 if fmin == -1 else fmin
    fmax = sample_rate / 2 if fmax == -1 else fmax
    basis = torch.from_numpy(slaney_mel_basis(sample_rate, fft_size, num_mels, fmin, fmax))
    mel = torch.log10(torch.clamp(basis @ spec, min=eps))
    return mel.t().contiguous().numpy()


def load_wav(path: str, sample_rate: int) -> np.ndarray:
    """16-bit / float wav -> mono float32 at `sample_rate` (polyphase resampling when needed)."""
    from scipy.io import wavfile
    from scipy.signal import resample_poly
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float32) / 2147483648.0
    else:
        data = data.astype(np.float32)
    if data.ndim > 1:
        data = data.mean(axis=1)
    if sr != sample_rate:
        g = math.gcd(sr, sample_rate)
        data = resample_poly(data, sample_rate // g, sr // g).astype(np.float32)
    return data


def integrated_loudness(wav: np.ndarray, rate: int) -> float:
    """ITU-R BS.1770-4 integrated loudness (LUFS) of a mono signal, as `pyloudnorm.Meter(rate).integrated_loudness`
    computes it (K-weighting = high-shelf 1500 Hz +4 dB, Q 1/sqrt(2), then high-pass 38 Hz, Q 0.5, both as RBJ biquads;
    400 ms blocks with 75 % overlap; absolute gate -70 LUFS, relative gate -10 LU).  pyloudnorm is not in the image:
    restated from the published algorithm, parity against the package itself unpinned."""
    from scipy.signal import lfilter
    x = np.asarray(wav, dtype=np.float64)

    def biquad(kind, G, Q, fc):
        A = 10 ** (G / 40.0)
        w0 = 2.0 * np.pi * (fc / rate)
        alpha = np.sin(w0) / (2.0 * Q)
        c = np.cos(w0)
        if kind == "high_shelf":
            b = np.array([A * ((A + 1) + (A - 1) * c + 2 * np.sqrt(A) * alpha), -2 * A * ((A - 1) + (A + 1) * c),
                          A * ((A + 1) + (A - 1) * c - 2 * np.sqrt(A) * alpha)])
            a = np.array([(A + 1) - (A - 1) * c + 2 * np.sqrt(A) * alpha, 2 * ((A - 1) - (A + 1) * c),
                          (A + 1) - (A - 1) * c - 2 * np.sqrt(A) * alpha])
        else:
            b = np.array([(1 + c) / 2, -(1 + c), (1 + c) / 2])
            a = np.array([1 + alpha, -2 * c, 1 - alpha])
        return b / a[0], a / a[0]

    for kind, G, Q, fc in (("high_shelf", 4.0, 1 / np.sqrt(2), 1500.0), ("high_pass", 0.0, 0.5, 38.0)):
        b, a = biquad(kind, G, Q, fc)
        x = lfilter(b, a, x)
    T_g, overlap = 0.400, 0.75
    step = 1.0 - overlap
    T = x.shape[0] / rate
    n_blocks = int(np.round((T - T_g) / (T_g * step)) + 1)
    if n_blocks < 1:
        raise ValueError("audio must be longer than the 400 ms gating block")
    z = np.empty(n_blocks)
    for j in range(n_blocks):
        lo, hi = int(T_g * (j * step) * rate), int(T_g * (j * step + 1) * rate)
        z[j] = np.sum(np.square(x[lo:hi])) / (T_g * rate)
    with np.errstate(divide="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
    keep = l >= -70.0
    if not keep.any():
        return -np.inf
    gamma_r = -0.691 + 10.0 * np.log10(np.mean(z[keep])) - 10.0
    keep = (l > gamma_r) & (l > -70.0)
    if not keep.any():
        return -np.inf
    return float(-0.691 + 10.0 * np.log10(np.mean(z[keep])))


def loudness_normalize(wav: np.ndarray, rate: int, target_lufs: float = -22.0) -> np.ndarray:
    """`loud_norm=True` branch of librosa_wav2spec (utils/audio/__init__.py:57-62): normalise to -22 LUFS, then peak-limit."""
    loud = integrated_loudness(wav, rate)
    out = np.asarray(wav, dtype=np.float32) * np.float32(10.0 ** ((target_lufs - loud) / 20.0))
    peak = np.abs(out).max()
    return out / peak if peak > 1 else out


def save_wav(wav: np.ndarray, path: str, sr: int, norm: bool = False):
    from scipy.io import wavfile
    if norm:
        wav = wav / np.abs(wav).max()
    wavfile.write(path[:-4] + ".wav", sr, (wav * 32767).astype(np.int16))
