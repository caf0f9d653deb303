"""GPU: the log-mel front-end against the host restatement of the reference's offline STFT path
(conan_b200/audio.py::wav2mel = torch.stft + Slaney basis, itself checked against torchaudio's filterbank in test_host)."""
import numpy as np
import pytest
import torch

from conan_b200 import audio, synth

pytestmark = pytest.mark.gpu

LOGMEL_TOL = 1e-4          # log10 units, after the [-6, 1.5] clip; fp32 dot-product DFT vs pocketfft (measured 3e-6)


def _audio(n, seconds, seed):
    g = np.random.default_rng(seed)
    t = np.arange(int(seconds * 16000)) / 16000.0
    out = []
    for i in range(n):
        f0 = 110.0 * (1 + i) + 40 * np.sin(2 * np.pi * 0.7 * t)
        x = sum(np.sin(2 * np.pi * np.cumsum(f0 * h) / 16000.0) / h for h in range(1, 12))
        x = 0.08 * x * (0.6 + 0.4 * np.sin(2 * np.pi * 1.3 * t)) + 0.01 * g.standard_normal(t.shape)
        out.append(x.astype(np.float32))
    return np.stack(out)


def _hp():
    import os
    from conan_b200.hparams import set_hparams
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cwd = os.getcwd()
    os.chdir(root)
    try:
        return set_hparams(config="egs/conan_emformer.yaml", print_hparams=False, global_hparams=False)
    finally:
        os.chdir(cwd)


HP = _hp()


@pytest.fixture(scope="module")
def fe():
    from conan_b200.frontend import GpuLogMel
    return GpuLogMel(HP)


def _host_mel(x):
    hp = HP
    m = audio.wav2mel(x, fft_size=hp["fft_size"], hop_size=hp["hop_size"], win_length=hp["win_size"], num_mels=hp["audio_num_mel_bins"],
                      fmin=hp["fmin"], fmax=hp["fmax"], sample_rate=hp["audio_sample_rate"])
    return np.clip(m, hp["mel_vmin"], hp["mel_vmax"])


@pytest.mark.parametrize("seconds", [0.5, 2.013])
def test_offline_logmel_matches_host(fe, seconds):
    wav = _audio(3, seconds, 1)
    mel = fe.offline(wav).cpu().numpy()
    for i in range(3):
        ref = _host_mel(wav[i])
        assert mel[i].shape == ref.shape
        err = np.abs(mel[i] - ref).max()
        print("logmel max-abs", err, "frames", ref.shape[0])
        assert err < LOGMEL_TOL


def test_silence_and_clip(fe):
    wav = np.zeros((1, 4000), np.float32)
    mel = fe.offline(wav).cpu().numpy()
    assert (mel == HP["mel_vmin"]).all()              # log10(1e-6) = -6 = the clip floor
    loud = _audio(1, 0.3, 2) * 50.0
    assert fe.offline(loud).max().item() <= HP["mel_vmax"]


def test_streamed_pcm_equals_offline_bitwise(fe):
    """Frames come out as soon as their samples exist, identical to the whole-utterance result, whatever the piece sizes."""
    from conan_b200.frontend import StreamingLogMel
    wav = torch.from_numpy(_audio(2, 1.7, 3))
    off = fe.offline(wav)
    s = StreamingLogMel(fe, 2, max_seconds=4)
    got, pos, g = [], 0, np.random.default_rng(0)
    n_before_final = 0
    while pos < wav.shape[1]:
        m = int(g.integers(1, 3000))
        piece = wav[:, pos:pos + m]
        pos += piece.shape[1]
        out = s.push(piece)
        if out is not None:
            got.append(out)
            n_before_final += out.shape[1]
    out = s.push(wav[:, :0], final=True)
    if out is not None:
        got.append(out)
    got = torch.cat(got, 1)
    assert got.shape == off.shape and torch.equal(got, off)
    assert 0 < off.shape[1] - n_before_final <= 3                     # only the look-ahead frames wait for the end of the stream


def test_gpu_logmel_against_independent_librosa_restatement(fe):
    """f1 pinning: `conan_logmel` against the committed golden of oracle/librosa_restatement.py (float64 numpy, written
    independently of conan_b200/audio.py from librosa's documented stft / filters.mel definitions)."""
    import os
    from oracle import librosa_restatement as lr
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "logmel_f1.npz"))
    wav = lr.test_signal(int(d["seed"]), int(d["n"]))
    mel = fe.offline(wav)[0].cpu().numpy()
    err = np.abs(mel - d["mel"]).max()
    print("GPU log-mel vs independent float64 restatement: max-abs", err)
    assert mel.shape == d["mel"].shape and err < LOGMEL_TOL


def test_stream_server_pcm_feed_matches_offline_path(fe, state_dicts):
    """f3 serving shell on the real engine: PCM fed in uneven pieces (per-stream framing on the GPU front-end, input ring,
    packed steps, output jitter buffer) == the offline path (whole-utterance mel, one stream through infer-style stepping)."""
    from conan_b200.engine import Engine, make_config
    from conan_b200.scheduler import ChunkScheduler
    from conan_b200.serving import StreamServer
    eng = Engine(*state_dicts, make_config(max_slots=4, max_ref_frames=64))
    try:
        wavs = _audio(2, 1.21, 9)
        refs = [synth.synth_mel(40, 70 + i).numpy() for i in range(2)]
        # offline: whole-utterance mel, scheduler one stream at a time
        expect = []
        sch = ChunkScheduler(eng, 4)
        for i in range(2):
            mel = fe.offline(wavs[i])[0].cpu().numpy()
            sid = sch.open(refs[i])
            sch.push(sid, mel)
            sch.end(sid)
            out = []
            while not sch.finished(sid):
                out.append(sch.step()[sid][0])
            sch.close(sid)
            expect.append((np.concatenate(out), mel))
        # served: two concurrent sessions, PCM in pieces of different sizes, pumped as frames become available
        srv = StreamServer(eng, 4, frontend=fe, keep_mel=True)
        sids = [srv.session_of(srv.admit(refs[i])) for i in range(2)]
        pos, piece = [0, 0], [997, 1603]
        while any(p < wavs.shape[1] for p in pos):
            for i in range(2):
                if pos[i] < wavs.shape[1]:
                    n = min(piece[i], wavs.shape[1] - pos[i])
                    took = srv.feed_pcm(sids[i], wavs[i, pos[i]:pos[i] + n], final=pos[i] + n >= wavs.shape[1])
                    assert took == n
                    pos[i] += n
            srv.pump()
        while not all(srv.done(s) for s in sids):
            srv.pump()
        for i, sid in enumerate(sids):
            wav = srv.read(sid)
            mel = srv.release(sid)
            assert wav.shape == expect[i][0].shape
            assert np.array_equal(wav, expect[i][0])          # same frames, same chunking, same kernels: bit-identical
        assert srv.stats["closed"] == 2 and srv.stats["admitted"] == 2
    finally:
        eng.close()
