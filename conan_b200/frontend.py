"""Log-mel front-end on the GPU: 16 kHz PCM -> the clipped log10 Slaney mel frames the chunk loop
consumes, through `conan_logmel` of the C ABI (windowed DFT on the conv-GEMM engine + one mel/log kernel).

Same arithmetic as the reference's offline `librosa_wav2spec` (utils/audio/__init__.py:36-80: centre-padded 1024-point STFT,
hop 320, periodic Hann, magnitude, Slaney mel 80..7600 Hz, log10(max(., 1e-6))) followed by the clip of
inference/Conan.py:58-70 -- but frame by frame, so PCM can be streamed: frame f needs samples up to 320 f + 511, i.e. 32 ms
of look-ahead, less than the two look-ahead frames the Emformer chunk already waits for."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .audio import slaney_mel_basis


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


class GpuLogMel:
    def __init__(self, hp: Dict, device: str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("conan_b200 has no CPU path: a CUDA (sm_100a) device is required")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.n_fft, self.hop, self.win = hp["fft_size"], hp["hop_size"], hp["win_size"]
        self.n_mels, self.sr = hp["audio_num_mel_bins"], hp["audio_sample_rate"]
        self.vmin, self.vmax = float(hp["mel_vmin"]), float(hp["mel_vmax"])
        if self.win != self.n_fft or self.hop % 16 != 0:
            raise ValueError("front-end expects win_size == fft_size and hop_size a multiple of 16")
        fmin = 0 if hp["fmin"] == -1 else hp["fmin"]
        fmax = self.sr / 2 if hp["fmax"] == -1 else hp["fmax"]
        self.bins = self.n_fft // 2 + 1
        self.taps = math.ceil(self.n_fft / self.hop)
        # window-weighted DFT basis, tap-major K = taps*hop (zero beyond n_fft): rows [cos | -sin]
        s = np.arange(self.taps * self.hop, dtype=np.float64)
        win = np.zeros_like(s)
        win[:self.n_fft] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(self.n_fft) / self.n_fft)      # periodic Hann
        ang = 2 * np.pi * np.outer(np.arange(self.bins, dtype=np.float64), s) / self.n_fft
        w = np.concatenate([np.cos(ang) * win, -np.sin(ang) * win], axis=0)
        self.dft_w = torch.from_numpy(w.astype(np.float32)).to(self.device).contiguous()
        basis = slaney_mel_basis(self.sr, self.n_fft, self.n_mels, fmin, fmax)                        # [n_mels, bins]
        self.basis_t = torch.from_numpy(np.ascontiguousarray(basis.T)).to(self.device)
        self.pad = self.n_fft // 2

    # ------------------------------------------------------------------ core call
    def frames(self, rows: torch.Tensor, row0: int, n_frames: int) -> torch.Tensor:
        """rows [n, R, hop] fp32 on the device (centre-padded signal) -> mel [n, n_frames, n_mels] for frames row0.."""
        n, R, hop = rows.shape
        assert hop == self.hop and rows.is_contiguous() and rows.dtype == torch.float32
        mel = torch.empty(n, n_frames, self.n_mels, device=self.device)
        spec = torch.empty(n * n_frames * 2 * self.bins, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.conan_logmel(_ptr(rows), n, R, self.hop, self.taps, row0, n_frames, _ptr(self.dft_w), self.bins,
                                         _ptr(self.basis_t), self.n_mels, 1e-6, self.vmin, self.vmax, _ptr(spec), _ptr(mel), st),
                   "logmel")
        return mel

    def n_frames_for(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop                       # centre-padded STFT

    def offline(self, wav) -> torch.Tensor:
        """wav [n, samples] (or [samples]) float -> mel [n, T, n_mels] with T = 1 + samples // hop (whole utterances)."""
        x = torch.as_tensor(np.asarray(wav, dtype=np.float32) if not torch.is_tensor(wav) else wav, dtype=torch.float32)
        if x.dim() == 1:
            x = x[None]
        n, m = x.shape
        T = self.n_frames_for(m)
        R = T + self.taps - 1
        sig = torch.zeros(n, R * self.hop, device=self.device)
        sig[:, self.pad:self.pad + m] = x.to(self.device)
        return self.frames(sig.view(n, R, self.hop), 0, T)


class StreamingLogMel:
    """PCM in arbitrary-sized pieces -> mel frames as soon as their samples exist (lock-step streams of one batch)."""

    def __init__(self, fe: GpuLogMel, n_streams: int, max_seconds: float = 60.0):
        self.fe, self.n = fe, n_streams
        self.cap_rows = int(max_seconds * fe.sr) // fe.hop + fe.taps + 2
        self.sig = torch.zeros(n_streams, self.cap_rows * fe.hop, device=fe.device)
        self.received = 0           # samples pushed so far
        self.done = 0               # frames emitted so far

    def push(self, wav: torch.Tensor, final: bool = False) -> Optional[torch.Tensor]:
        """wav [n, m] new samples (m may be 0 with final=True).  Returns the newly complete frames [n, f, n_mels] or None."""
        fe = self.fe
        m = wav.shape[1]
        if fe.pad + self.received + m > self.sig.shape[1]:
            raise RuntimeError("StreamingLogMel: stream longer than max_seconds")
        if m:
            self.sig[:, fe.pad + self.received: fe.pad + self.received + m] = wav.to(fe.device, torch.float32)
            self.received += m
        if final:
            avail = fe.n_frames_for(self.received)                                   # zero tail = centre padding
        else:
            # frame f reads padded samples [f*hop, f*hop + n_fft): complete once received >= f*hop + n_fft - pad
            avail = (self.received + fe.pad - fe.n_fft) // fe.hop + 1 if self.received + fe.pad >= fe.n_fft else 0
        if avail <= self.done:
            return None
        mel = fe.frames(self.sig.view(self.n, self.cap_rows, fe.hop), self.done, avail - self.done)
        self.done = avail
        return mel
