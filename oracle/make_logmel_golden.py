"""Writes tests/golden/logmel_f1.npz: the log-mel of a seeded test signal computed by oracle/librosa_restatement.py (float64),
clipped like inference/Conan.py:58-70.  Re-run: `python -m oracle.make_logmel_golden`."""
import os

import numpy as np

from oracle import librosa_restatement as lr

if __name__ == "__main__":
    wav = lr.test_signal()
    mel = np.clip(lr.log_mel(wav), -6.0, 1.5).astype(np.float32)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "logmel_f1.npz")
    np.savez_compressed(out, mel=mel, seed=7, n=9000, wav_checksum=np.float64(np.abs(wav.astype(np.float64)).sum()))
    print("wrote", out, mel.shape)
