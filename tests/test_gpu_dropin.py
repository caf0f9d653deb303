"""GPU: the reference-facing surface (StreamingVoiceConversion built from checkpoint directories in the
reference's layout, the vocoder registry, the module-level views) against the golden vectors."""
import os

import numpy as np
import pytest
import torch

from conan_b200 import ckpt, synth
from util import snr_ac_db

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def svc(tmp_path_factory, state_dicts):
    from conan_b200.hparams import set_hparams
    from conan_b200.streaming import StreamingVoiceConversion
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tmp = tmp_path_factory.mktemp("ckpts")
    sd_c, sd_e, sd_v = state_dicts
    ckpt.save_checkpoint(sd_c, str(tmp / "conan"), "model", steps=160000)
    ckpt.save_checkpoint(sd_e, str(tmp / "emformer"), "model", steps=3)
    ckpt.save_checkpoint(sd_v, str(tmp / "hifigan_vc"), "model_gen", steps=1, config=dict(synth.DEFAULT_VOC_HP, upsample="shuffle", resblock="1"))
    cwd = os.getcwd()
    os.chdir(root)
    hp = set_hparams(config="egs/conan_emformer.yaml", print_hparams=False)
    os.chdir(cwd)
    hp["work_dir"], hp["emformer_ckpt"], hp["vocoder_ckpt"] = str(tmp / "conan"), str(tmp / "emformer"), str(tmp / "hifigan_vc")
    s = StreamingVoiceConversion(hp, max_streams=4, max_ref_frames=256)
    yield s
    s.engine.close()


def test_infer_mels_matches_reference_golden(svc, golden_dir):
    d = np.load(os.path.join(golden_dir, "e2e_short.npz"))
    ref = synth.synth_mel(int(d["ref_frames"]), int(d["ref_seed"])).numpy()
    src = synth.synth_mel(int(d["src_frames"]), int(d["src_seed"])).numpy()
    wav, mel = svc.infer_mels(ref, src)
    assert wav.dtype == np.float32 and wav.shape == d["wav"].shape and mel.shape == d["mel"].shape
    assert np.abs(mel - d["mel"]).max() <= 1e-3
    assert snr_ac_db(d["wav"], wav) >= 40.0
    wav2, _ = svc.infer_mels(ref, src)                     # slot reuse: a second utterance starts from clean state
    assert np.array_equal(wav, wav2)


def test_module_level_views(svc, golden_dir, state_dicts):
    from oracle.incremental import EmformerOracle, assemble_chunk
    d = np.load(os.path.join(golden_dir, "vocoder_24f.npz"))
    wav = svc.vocoder.spec2wav(d["mel"])
    assert wav.shape == d["wav"].shape and snr_ac_db(d["wav"], wav) >= 40.0
    wav = svc.vocoder.spec2wav(d["mel"][:22])              # T not a multiple of the chunk: padded and cut
    assert wav.shape == (22 * 320,) and snr_ac_db(d["wav"][:22 * 320], wav) >= 40.0
    # emformer.emformer.infer(chunk, lengths, state) + proj, as the reference loop calls them
    src = synth.synth_mel(14, 3)[None]
    o = EmformerOracle(state_dicts[1])
    o.reset(1)
    state = None
    for pos in (0, 4, 8):
        chunk, _ = assemble_chunk(src, pos)
        out, lengths, state = svc.emformer.emformer.infer(chunk, torch.full((1,), 6), state)
        with torch.no_grad():
            ref = o.step(chunk)
        assert int(lengths[0]) == 4 and (out.cpu() - ref).abs().max() < 1e-4
        logits = svc.emformer.proj(out)
        assert (logits.cpu() - o.logits(ref)).abs().max() < 1e-4
    with pytest.raises(ValueError):
        svc.emformer.emformer.infer(torch.zeros(1, 5, 80), torch.full((1,), 5), None)
    # model(content=..., ref=..., infer=True)["mel_out"] over full history
    g = np.load(os.path.join(golden_dir, "e2e_short.npz"))
    ref_mel = synth.synth_mel(int(g["ref_frames"]), int(g["ref_seed"]))[None]
    out = svc.model(content=torch.from_numpy(g["tokens"].astype(np.int64))[None], spk_embed=None, target=None, ref=ref_mel,
                    f0=None, uv=None, infer=True, global_steps=200000)
    assert np.abs(out["mel_out"][0].cpu().numpy() - g["mel"]).max() <= 1e-3


def test_infer_once_from_wav_files(svc, tmp_path):
    from scipy.io import wavfile
    sr = 16000
    t = np.arange(sr) / sr
    for name, f0 in (("ref", 180.0), ("src", 120.0)):
        x = 0.3 * np.sin(2 * np.pi * f0 * t) + 0.05 * np.random.default_rng(1).standard_normal(sr)
        wavfile.write(str(tmp_path / f"{name}.wav"), sr, (x * 32767).astype(np.int16))
    wav, mel = svc.infer_once({"ref_wav": str(tmp_path / "ref.wav"), "src_wav": str(tmp_path / "src.wav")})
    T = sr // 320 + 1
    assert mel.shape == (T, 80) and wav.shape == (T * 320,) and np.isfinite(wav).all() and np.abs(wav).max() <= 1.0


def test_unknown_vocoder_is_rejected(svc):
    from conan_b200.streaming import StreamingVoiceConversion
    hp = dict(svc.hparams, vocoder="NoSuchVocoder")
    with pytest.raises(ValueError):
        StreamingVoiceConversion(hp)


def test_batch_runner_converts_pairs_concurrently_and_matches_infer_once(svc, tmp_path, monkeypatch):
    """inference/run_voice_conversion_nvae.py on the real engine: five pairs over four slots (admission as slots free up), every
    saved wav equal to the one-at-a-time `infer_once` result of the same pair; a missing file is reported, not raised."""
    import json
    from scipy.io import wavfile
    from conan_b200.serving import VoiceConversionRunner
    sr = 16000
    rng = np.random.default_rng(4)
    names = []
    for i, secs in enumerate((0.7, 1.0, 0.55, 0.9, 0.8, 0.6)):
        t = np.arange(int(sr * secs)) / sr
        x = 0.3 * np.sin(2 * np.pi * (110.0 + 30 * i) * t) + 0.04 * rng.standard_normal(t.shape)
        p = str(tmp_path / f"utt{i}.wav")
        wavfile.write(p, sr, (x * 32767).astype(np.int16))
        names.append(p)
    pairs = [{"ref_wav": names[(i + 1) % 6], "src_wav": names[i], "src_corpus": "syn", "src_utt_id": f"u{i}", "output_name": f"o{i}"}
             for i in range(5)]
    pairs.append({"ref_wav": names[0], "src_wav": str(tmp_path / "nope.wav"), "src_corpus": "syn", "src_utt_id": "missing", "output_name": "x"})
    cfg = tmp_path / "voice_conversion_config.json"
    cfg.write_text(json.dumps({"total_pairs": len(pairs), "conversion_pairs": pairs}))
    out = tmp_path / "out"
    runner = VoiceConversionRunner(str(cfg), hparams=svc.hparams, engine=svc, output_dir=str(out))
    rep = runner.run_all_conversions(batch_size=2)
    assert rep["successful"] == 5 and rep["failed"] == 1 and "Pair 5" in rep["errors"][0]
    for i in range(5):
        _, got = wavfile.read(str(out / f"syn_u{i}.wav"))
        wav, _ = svc.infer_once({"ref_wav": pairs[i]["ref_wav"], "src_wav": pairs[i]["src_wav"]})
        assert got.shape == wav.shape and np.array_equal(got, (wav * 32767).astype(np.int16))
    assert not svc.scheduler.streams and len(svc.scheduler.free) == svc.scheduler.S
