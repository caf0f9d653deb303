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


def save_wav(wav: np.ndarray, path: str, sr: int, norm: bool = False):
    from scipy.io import wavfile
    if norm:
        wav = wav / np.abs(wav).max()
    wavfile.write(path[:-4] + ".wav", sr, (wav * 32767).astype(np.int16))
