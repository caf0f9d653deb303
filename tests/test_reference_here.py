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
