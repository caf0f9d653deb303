"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(`inference/Conan.py::StreamingVoiceConversion`, its own loop, its own modules,
its own checkpoint loader) on the synthetic checkpoints of conan_b200.synth.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
The only intervention is `_wav_to_mel`, whose librosa front-end is absent from
the image; it is replaced by a lookup that returns seeded synthetic log-mel for a
"path" of the form synth:<seed>:<frames>.  Everything after the mel (chunk
assembly, Emformer state handling, full-history recompute, vocoder slicing) is
the reference's code as shipped.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from conan_b200 import ckpt, synth  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SEED = 1234

# (name, ref seed, ref frames, src seed, src frames)
CASES = [
    ("e2e_short", 11, 150, 21, 22),     # 5 full chunks + a 2-frame tail; T_ref % 4 == 2
    ("e2e_long", 12, 64, 22, 240),      # 60 chunks: left context saturates (50) and the K/V ring wraps
]


def _path(seed, frames):
    return f"synth:{seed}:{frames}"


def _mel_of(path: str) -> np.ndarray:
    _, seed, frames = path.split(":")
    return synth.synth_mel(int(frames), int(seed)).numpy()


def build_reference_engine(tmp):
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    from utils.commons.hparams import set_hparams, hparams
    hp = set_hparams(config="egs/conan_emformer.yaml", print_hparams=False)
    voc_hp = set_hparams(config="egs/hifi_16k320_shuffle.yaml", print_hparams=False, global_hparams=False)
    sd_c, sd_e, sd_v = synth.make_all_state_dicts(SEED)
    ckpt.save_checkpoint(sd_c, f"{tmp}/conan", "model")
    ckpt.save_checkpoint(sd_e, f"{tmp}/emformer", "model")
    ckpt.save_checkpoint(sd_v, f"{tmp}/hifigan_vc", "model_gen",
                         config={k: voc_hp[k] for k in ("upsample_rates", "upsample_kernel_sizes", "upsample",
                                                        "upsample_initial_channel", "resblock",
                                                        "resblock_kernel_sizes", "resblock_dilation_sizes")})
    for d in (hp, hparams):
        d["work_dir"] = f"{tmp}/conan"
        d["emformer_ckpt"] = f"{tmp}/emformer"
        d["vocoder_ckpt"] = f"{tmp}/hifigan_vc"
    from inference.Conan import StreamingVoiceConversion
    StreamingVoiceConversion._wav_to_mel = staticmethod(_mel_of)
    torch.set_num_threads(os.cpu_count())
    eng = StreamingVoiceConversion(hparams)
    return eng, (sd_c, sd_e, sd_v)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        eng, (sd_c, sd_e, sd_v) = build_reference_engine(tmp)
        sums = np.array([synth.state_dict_checksum(s) for s in (sd_c, sd_e, sd_v)])
        for name, rs, rf, ss, sf in CASES:
            t0 = time.time()
            wav, mel = eng.infer_once({"ref_wav": _path(rs, rf), "src_wav": _path(ss, sf)})
            # stage tensors of the reference, from one more full-history pass of its own modules
            src = torch.from_numpy(_mel_of(_path(ss, sf)))[None]
            with torch.no_grad():
                logits = eng.emformer.inference(src)                      # modules/Emformer/emformer.py:49-98
                tokens = logits.argmax(-1)
                srt = logits.sort(-1).values
                ret = eng.model(content=tokens, spk_embed=None, target=None,
                                ref=torch.from_numpy(_mel_of(_path(rs, rf)))[None], f0=None, uv=None,
                                infer=True, global_steps=200000)
            assert np.allclose(ret["mel_out"][0].numpy(), mel, atol=1e-4), "reference prefix consistency"
            np.savez_compressed(
                os.path.join(GOLDEN_DIR, f"{name}.npz"),
                seed=SEED, weight_checksums=sums, ref_seed=rs, ref_frames=rf, src_seed=ss, src_frames=sf,
                wav=wav.astype(np.float32), mel=mel.astype(np.float32),
                tokens=tokens[0].numpy().astype(np.int16),
                logits=logits[0].numpy().astype(np.float32),
                argmax_margin=(srt[0, :, -1] - srt[0, :, -2]).numpy().astype(np.float32),
                style_embed=ret["style_embed"][0, 0].numpy(),
                uv_pred=ret["uv_pred"][0].numpy(),
                f0_denorm=ret["f0_denorm_pred"][0].numpy(),
                decoder_inp=ret["decoder_inp"][0].numpy(),
                pitch_inp=ret["pitch_embed"][0].numpy(),
            )
            print(f"{name}: wav {wav.shape} mel {mel.shape} rms {np.sqrt((wav ** 2).mean()):.4f} "
                  f"min argmax margin {float((srt[0, :, -1] - srt[0, :, -2]).min()):.2e} "
                  f"voiced {(ret['f0_denorm_pred'][0] > 0).float().mean():.2f} in {time.time() - t0:.1f}s")
        # vocoder-only golden through the reference's spec2wav
        mel_in = (torch.randn(24, 80, generator=torch.Generator().manual_seed(5)) * 0.6).numpy()
        wav = eng.vocoder.spec2wav(mel_in)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "vocoder_24f.npz"), seed=SEED, weight_checksums=sums,
                            mel=mel_in.astype(np.float32), wav=wav.astype(np.float32))
        print("vocoder_24f: rms", float(np.sqrt((wav ** 2).mean())))


if __name__ == "__main__":
    main()
