"""Test infrastructure: an independent numpy restatement of the two librosa calls the reference's mel front-end makes
(utils/audio/__init__.py:62-72), written from librosa's documented algorithm and NOT sharing code with conan_b200/audio.py:

    librosa.stft(wav, n_fft, hop_length, win_length, window="hann", pad_mode="constant")     (center=True default)
    librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)                                        (htk=False, norm="slaney" defaults)

librosa itself is not installed in this image and cannot be fetched (no network), so parity against the package is UNPINNED;
what this file pins is the build's front-end (host torch.stft path and the GPU `conan_logmel` path) against a second,
separately written float64 implementation of the same published definition.  Definitions restated:
  * stft: zero-pad n_fft // 2 samples on both sides; frame t = padded[t*hop : t*hop + n_fft] * w, w = scipy-style periodic
    Hann  0.5 - 0.5 cos(2 pi n / N); real FFT; 1 + len // hop frames;
  * Slaney mel scale: linear below 1 kHz (200/3 Hz per mel), logarithmic above (log(6.4) / 27 per mel);
  * filters.mel: n_mels + 2 band edges equally spaced in mel between fmin and fmax; triangle i rises over [f_i, f_i+1] and
    falls over [f_i+1, f_i+2], evaluated at the FFT bin centres k * sr / n_fft; Slaney normalisation 2 / (f_i+2 - f_i).
"""
import numpy as np


def _hz_to_mel_slaney(hz):
    hz = np.atleast_1d(np.asarray(hz, dtype=np.float64))
    out = hz / (200.0 / 3.0)
    big = hz >= 1000.0
    out[big] = 15.0 + np.log(hz[big] / 1000.0) / (np.log(6.4) / 27.0)
    return out


def _mel_to_hz_slaney(mel):
    mel = np.atleast_1d(np.asarray(mel, dtype=np.float64))
    out = mel * (200.0 / 3.0)
    big = mel >= 15.0
    out[big] = 1000.0 * np.exp((np.log(6.4) / 27.0) * (mel[big] - 15.0))
    return out


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    edges = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin)[0], _hz_to_mel_slaney(fmax)[0], n_mels + 2))
    bins = np.arange(n_fft // 2 + 1, dtype=np.float64) * sr / n_fft
    fb = np.zeros((n_mels, bins.shape[0]))
    for i in range(n_mels):
        lo, mid, hi = edges[i], edges[i + 1], edges[i + 2]
        up = (bins - lo) / (mid - lo)
        down = (hi - bins) / (hi - mid)
        fb[i] = np.clip(np.minimum(up, down), 0.0, None) * (2.0 / (hi - lo))
    return fb


def stft_magnitude(wav, n_fft, hop):
    x = np.concatenate([np.zeros(n_fft // 2), np.asarray(wav, dtype=np.float64), np.zeros(n_fft // 2)])
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)
    n_frames = 1 + len(wav) // hop
    frames = np.stack([x[t * hop:t * hop + n_fft] * w for t in range(n_frames)])
    return np.abs(np.fft.rfft(frames, axis=1))                        # [T, bins]


def log_mel(wav, sr=16000, n_fft=1024, hop=320, n_mels=80, fmin=80, fmax=7600, eps=1e-6):
    """librosa_wav2spec's mel (utils/audio/__init__.py:62-76): log10(max(eps, mel_basis @ |stft|)) -> [T, n_mels] float64."""
    mag = stft_magnitude(wav, n_fft, hop)
    return np.log10(np.maximum(eps, mag @ mel_filterbank(sr, n_fft, n_mels, fmin, fmax).T))


def test_signal(seed=7, n=9000, sr=16000):
    """Seeded harmonic sweep + noise in [-1, 1] (SURVEY.md 8d synthetic audio)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    f0 = 120.0 + 80.0 * t
    x = sum(np.sin(2 * np.pi * h * np.cumsum(f0) / sr) / h for h in range(1, 9))
    x = 0.25 * x / np.abs(x).max() + 0.02 * rng.standard_normal(n)
    return x.astype(np.float32)
