"""YAML hparams with `base_config` inheritance and `-hp k=v` overrides, drop-in for the
reference's `utils/commons/hparams.py::set_hparams` (:25-131): same signature, same global
`hparams` dict that the hot path's builders read, same `work_dir = checkpoints/{exp_name}` rule.
Only the behaviours the inference path uses are kept (no interactive --remove prompt)."""
from __future__ import annotations

import argparse
import os
from typing import Dict

import yaml

hparams: Dict = {}


def _merge(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def _load_chain(path: str, seen: set, chain: list) -> dict:
    """Depth-first inheritance: bases first (in list order), the including file last; a file is
    visited once; a base starting with '.' is relative to the including file, anything else to cwd."""
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    seen.add(path)
    out: dict = {}
    bases = cfg.get("base_config", [])
    if not isinstance(bases, list):
        bases = [bases]
        cfg["base_config"] = bases          # the reference normalises this key to a list in place
    for b in bases:
        if b.startswith("."):
            b = os.path.normpath(os.path.join(os.path.dirname(path), b))
        if b not in seen:
            _merge(out, _load_chain(b, seen, chain))
    _merge(out, cfg)
    chain.append(path)
    return out


def _apply_overrides(cfg: dict, spec: str):
    """-hp "a=1,b.c=2,d=[1 1 1]": typed by the existing value, lists written with spaces."""
    for item in spec.split(","):
        if not item.strip():
            continue
        key, val = item.split("=", 1)
        val = val.strip("'\" ")
        node = cfg
        parts = key.strip().split(".")
        for p in parts[:-1]:
            node = node[p]
        leaf = parts[-1]
        old = node.get(leaf)
        if val in ("True", "False") or isinstance(old, (bool, list, dict)):
            if isinstance(old, list):
                val = val.replace(" ", ",")
            node[leaf] = eval(val)   # noqa: S307  (same contract as the reference loader)
        elif old is None:
            node[leaf] = yaml.safe_load(val)
        else:
            node[leaf] = type(old)(val)


def set_hparams(config: str = "", exp_name: str = "", hparams_str: str = "", print_hparams: bool = True,
                global_hparams: bool = True) -> dict:
    reset = infer = debug = validate = False
    if config == "" and exp_name == "":
        ap = argparse.ArgumentParser(description="")
        ap.add_argument("--config", type=str, default="")
        ap.add_argument("--exp_name", type=str, default="")
        ap.add_argument("-hp", "--hparams", type=str, default="")
        ap.add_argument("--infer", action="store_true")
        ap.add_argument("--validate", action="store_true")
        ap.add_argument("--reset", action="store_true")
        ap.add_argument("--remove", action="store_true")
        ap.add_argument("--debug", action="store_true")
        args, unknown = ap.parse_known_args()
        print("| Unknown hparams: ", unknown)
        config, exp_name, hparams_str, reset, infer = args.config, args.exp_name, args.hparams, args.reset, args.infer
        debug, validate = args.debug, args.validate
    assert config != "" or exp_name != ""
    if config != "":
        assert os.path.exists(config), config
    work_dir = f"checkpoints/{exp_name}" if exp_name else ""
    saved = {}
    if work_dir and os.path.exists(f"{work_dir}/config.yaml"):
        with open(f"{work_dir}/config.yaml") as f:
            saved = yaml.safe_load(f) or {}
    chain: list = []
    cfg: dict = {}
    if config:
        cfg.update(_load_chain(config, set(), chain))
    if not reset:
        cfg.update(saved)
    cfg["work_dir"] = work_dir
    if hparams_str:
        _apply_overrides(cfg, hparams_str)
    # assigned unconditionally, as the reference does (utils/commons/hparams.py:113-116): a saved config.yaml cannot mask --infer
    cfg["infer"], cfg["debug"], cfg["validate"] = infer, debug, validate
    cfg["exp_name"] = exp_name
    if global_hparams:
        hparams.clear()
        hparams.update(cfg)
        if print_hparams:
            print("| Hparams chains: ", chain)
    return cfg
