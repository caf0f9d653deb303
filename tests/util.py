import numpy as np


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return 10.0 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-300))


def snr_ac_db(ref, x):
    """SNR after removing the reference's DC offset from both signals (the synthetic vocoder
    output has a DC component; this is the stricter figure)."""
    m = np.asarray(ref, dtype=np.float64).mean()
    return snr_db(np.asarray(ref, dtype=np.float64) - m, np.asarray(x, dtype=np.float64) - m)
