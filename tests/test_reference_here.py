"""Build-container only (needs /root/reference): the unmodified reference vs the oracle and the host
mirror.  Skipped on the GPU box, where the reference tree does not exist."""
import os

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_hparams_loader_agrees_with_reference_on_reference_yamls(monkeypatch):
    ref_import.install()
    monkeypatch.chdir(ref_import.REF_ROOT)
    from utils.commons.hparams import set_hparams as ref_set
    from conan_b200.hparams import set_hparams as my_set
    for cfgfile, ov in (("egs/conan_emformer.yaml", "hidden_size=256,dec_dilations=[1 1 1 1]"), ("egs/hifi_16k320_shuffle.yaml", "max_updates=7")):
        a = ref_set(config=cfgfile, print_hparams=False, global_hparams=False, hparams_str=ov)
        b = my_set(config=cfgfile, print_hparams=False, global_hparams=False, hparams_str=ov)
        for k in a:
            if k in ("infer", "debug", "validate"):
                continue
            assert a[k] == b[k], (cfgfile, k)


def test_reference_modules_accept_synthetic_checkpoints_and_match_oracle(state_dicts):
    """One Conan.forward of the real reference class vs the oracle's open()+step() (stage tensors)."""
    from oracle.incremental import ConanOracle
    from conan_b200 import synth
    hp, conan, emf, voc, voc_hp = ref_import.build_reference_models()
    conan.load_state_dict(state_dicts[0], strict=True)
    ref = synth.synth_mel(70, 5)[None]
    tokens = torch.randint(0, 100, (1, 16), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ret = conan(content=tokens, ref=ref, infer=True, global_steps=200000)
        o = ConanOracle(state_dicts[0])
        o.open(ref)
        mel = torch.cat([o.step(tokens[:, i:i + 4]) for i in range(0, 16, 4)], 1)
    assert (o.style - ret["style_embed"][:, 0]).abs().max() < 1e-5
    assert (mel - ret["mel_out"]).abs().max() < 2e-5


def test_reference_constructed_checkpoints_load_key_for_key(tmp_path, capsys):
    """SURVEY 8f row f2.  The three REFERENCE modules are constructed by the reference's own constructors, their state_dicts
    are written by the reference's own `Trainer.dump_checkpoint` / `_atomic_save` (utils/commons/trainer.py:451-470, the layout a
    released checkpoint has), and read back through conan_b200.ckpt + weights.py:
      * every key the hot path needs exists in the reference module with the same shape;
      * every reference key is either consumed or on the explicit not-on-the-inference-path list;
      * packing (tap-major convs, weight-norm folding, pixel-shuffle row order, split fp16) accepts them;
      * strict=False drops a shape-mismatched key with the reference's message and keeps the initial value
        (utils/commons/ckpt_utils.py:48-58)."""
    import types
    import torch.nn as nn
    from conan_b200 import ckpt, synth
    from conan_b200.engine import make_config
    from conan_b200.weights import pack_engine_weights
    hp, conan, emf, voc, voc_hp = ref_import.build_reference_models()
    from utils.commons.trainer import Trainer

    def dump(children, work_dir, step):
        task = nn.Module()
        for k, m in children.items():
            setattr(task, k, m)
        fake = types.SimpleNamespace(current_epoch=3, global_step=step, best_val_results=0.0, optimizers=[],
                                     get_task_ref=lambda: task)
        fake.dump_checkpoint = lambda: Trainer.dump_checkpoint(fake)
        os.makedirs(work_dir, exist_ok=True)
        Trainer._atomic_save(fake, f"{work_dir}/model_ckpt_steps_{step}.ckpt")

    dump({"model": conan}, str(tmp_path / "conan"), 160000)
    dump({"model": emf}, str(tmp_path / "emformer"), 50000)
    disc = nn.Linear(2, 2)                                                     # the vocoder task also holds discriminators
    dump({"model_gen": voc, "model_disc": disc}, str(tmp_path / "hifigan_vc"), 400000)

    sd_c = ckpt.load_state_dict(str(tmp_path / "conan"), "model")
    sd_e = ckpt.load_state_dict(str(tmp_path / "emformer"), "model")
    sd_v = ckpt.load_state_dict(str(tmp_path / "hifigan_vc"), "model_gen")
    assert set(sd_c) == set(conan.state_dict()) and set(sd_e) == set(emf.state_dict()) and set(sd_v) == set(voc.state_dict())
    unused_ok = ("pitch_predictor.", "prosody_extractor.vqvae.ema_", "prosody_extractor.vqvae.data_initialized", "embed_positions.")
    for sd, spec, strict in ((sd_c, synth.conan_spec(hp), False), (sd_e, synth.emformer_spec(hp), False), (sd_v, synth.hifigan_spec(voc_hp), True)):
        keys = {k for k, *_ in spec}
        assert not keys - set(sd), sorted(keys - set(sd))[:5]
        for k, shape, *_ in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
        extra = [k for k in set(sd) - keys if not k.startswith(unused_ok)]
        assert not extra, extra[:5]
        got = ckpt.filter_to_spec(sd, spec, strict=strict and not (set(sd) - keys))
        assert all(torch.equal(got[k], sd[k]) for k in keys)
    cfg = make_config(hp, voc_hp, max_slots=2, max_ref_frames=64)
    packed = pack_engine_weights(ckpt.filter_to_spec(sd_c, synth.conan_spec(hp), False), ckpt.filter_to_spec(sd_e, synth.emformer_spec(hp), False),
                                 ckpt.filter_to_spec(sd_v, synth.hifigan_spec(voc_hp), False), cfg)
    assert all(torch.isfinite(t.float()).all() for t in packed.values())
    w = voc.ups[1].conv.conv                                                   # folded + shuffle-permuted rows of a reference module
    folded = (w.weight_v * (w.weight_g / w.weight_v.flatten(1).norm(dim=1).view(-1, 1, 1))).detach()
    r, co = 5, 128
    expect = folded.view(co, r, 256, 10).permute(1, 0, 2, 3).reshape(r * co, 256, 10).permute(0, 2, 1).reshape(r * co, -1)
    assert torch.allclose(packed["voc.up.1.w"].float(), expect.half().float())
    # strict=False drop semantics on a shape-mismatched tensor
    bad = dict(sd_c)
    bad["mel_out.weight"] = torch.zeros(81, 256)
    init = synth.make_state_dict(synth.conan_spec(hp), 1234)
    capsys.readouterr()
    got = ckpt.filter_to_spec(bad, synth.conan_spec(hp), strict=False, defaults=init)
    assert "| Unmatched keys:  mel_out.weight" in capsys.readouterr().out
    assert torch.equal(got["mel_out.weight"], init["mel_out.weight"])
