"""Checkpoint layout of the reference, read and written drop-in.

Layout (utils/commons/trainer.py:457-470 of the reference): a file
`model_ckpt_steps_{N}.ckpt` holding `{'state_dict': {child_name: child_state_dict}, ...}`
where child_name is `model` for the Conan and Emformer tasks and `model_gen` for
the vocoder task; the newest N wins (utils/commons/ckpt_utils.py:17-23).  A flat
`"child.key"` form is also accepted (ckpt_utils.py:35-47).  The vocoder directory
additionally holds `config.yaml` (tasks/tts/vocoder_infer/hifigan.py:15-16).
Weight-norm convolutions are stored un-folded as weight_g / weight_v and are
folded by `fold_weight_norm` at load time.
"""
from __future__ import annotations

import glob
import os
import re
from typing import Dict, Optional

import torch
import yaml


def get_all_ckpts(work_dir: str, steps: Optional[int] = None):
    pat = f"{work_dir}/model_ckpt_steps_{'*' if steps is None else steps}.ckpt"
    return sorted(glob.glob(pat), key=lambda x: -int(re.findall(r".*steps_(\d+)\.ckpt", x)[0]))


def load_state_dict(ckpt_base_dir: str, model_name: str = "model", force: bool = True) -> Optional[Dict[str, torch.Tensor]]:
    """Same selection rules as the reference's load_ckpt, but returns the sub state_dict
    instead of loading it into an nn.Module (there is no nn.Module in this build)."""
    if os.path.isfile(ckpt_base_dir):
        base_dir, ckpt_path = os.path.dirname(ckpt_base_dir), ckpt_base_dir
    else:
        base_dir = ckpt_base_dir
        paths = get_all_ckpts(ckpt_base_dir)
        ckpt_path = paths[0] if paths else None
    if ckpt_path is None:
        msg = f"| ckpt not found in {base_dir}."
        if force:
            raise AssertionError(msg)      # the reference asserts here (ckpt_utils.py:61-64)
        print(msg)
        return None
    checkpoint = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    state_dict = checkpoint["state_dict"]
    if any("." in k for k in state_dict.keys()):
        state_dict = {k[len(model_name) + 1:]: v for k, v in state_dict.items() if k.startswith(f"{model_name}.")}
    elif "." not in model_name:
        state_dict = state_dict[model_name]
    else:
        base, rest = model_name.split(".")[0], model_name[len(model_name.split(".")[0]) + 1:]
        state_dict = {k[len(rest) + 1:]: v for k, v in state_dict[base].items() if k.startswith(f"{rest}.")}
    print(f"| load '{model_name}' from '{ckpt_path}'.")
    return {k: v.detach().float().contiguous() for k, v in state_dict.items()}


def filter_to_spec(state_dict: Dict[str, torch.Tensor], spec, strict: bool, defaults: Optional[Dict[str, torch.Tensor]] = None):
    """The `cur_model.load_state_dict(state_dict, strict=strict)` step of the reference's load_ckpt
    (utils/commons/ckpt_utils.py:48-58), with `spec` = [(key, shape, ...)] standing in for `cur_model.state_dict()`.

    strict=True : a missing, unexpected or shape-mismatched key raises RuntimeError, as nn.Module.load_state_dict does.
    strict=False: a shape-mismatched key is DROPPED with the reference's own message (`| Unmatched keys: ...`), unexpected
                  keys are ignored, and a dropped / missing key keeps the module's initial value -- here `defaults[key]`
                  (the reference's nn.Module would keep its constructor init; this build has no nn.Module, so the caller
                  supplies the initial tensors, see streaming.build_engine).  Without `defaults` such a key is an error:
                  the engine cannot run on an unbound weight."""
    out, missing, mismatched = {}, [], []
    for key, shape, *_ in spec:
        if key not in state_dict:
            missing.append(key)
            continue
        t = state_dict[key]
        if tuple(t.shape) != tuple(shape):
            mismatched.append((key, tuple(shape), tuple(t.shape)))
            continue
        out[key] = t
    if strict:
        extra = sorted(set(state_dict) - {k for k, *_ in spec})
        if missing or extra or mismatched:
            raise RuntimeError("Error(s) in loading state_dict: "
                               + (f"Missing key(s): {missing[:5]}... " if missing else "")
                               + (f"Unexpected key(s): {extra[:5]}... " if extra else "")
                               + "".join(f"size mismatch for {k}: checkpoint {got}, model {want}. " for k, want, got in mismatched[:5]))
        return out
    for k, want, got in mismatched:
        print("| Unmatched keys: ", k, want, got)          # ckpt_utils.py:55
    holes = missing + [k for k, _, _ in mismatched]
    if holes:
        if defaults is None:
            raise KeyError(f"checkpoint leaves {len(holes)} tensors of the hot path without a value (e.g. '{holes[0]}') and no "
                           "initial values were given (strict=False keeps the module's init in the reference)")
        for k in holes:
            out[k] = defaults[k]
        print(f"| {len(holes)} keys left at their initial value (strict=False): {holes[:5]}{' ...' if len(holes) > 5 else ''}")
    return out


def fold_weight_norm(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """w = v * g / ||v||, norm over all dims but the out-channel one."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / nrm)


def save_checkpoint(state_dict: Dict[str, torch.Tensor], ckpt_dir: str, model_name: str, steps: int = 1,
                    config: Optional[dict] = None) -> str:
    """Writes `ckpt_dir/model_ckpt_steps_{steps}.ckpt` (and config.yaml) in the reference layout."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"model_ckpt_steps_{steps}.ckpt")
    tmp = path + ".part"
    torch.save({"state_dict": {model_name: dict(state_dict)}, "global_step": steps, "epoch": 0}, tmp)
    os.replace(tmp, path)
    if config is not None:
        with open(os.path.join(ckpt_dir, "config.yaml"), "w") as f:
            yaml.safe_dump(config, f)
    return path
