"""Synthetic weights and inputs for the Conan hot path.

No real checkpoints or audio are reachable (no network), so parity and
throughput are measured on seeded synthetic weights that use the *reference's
state_dict key names and shapes* (`modules/Conan/Conan.py:46-113`,
`modules/Emformer/emformer.py:14-30`, `modules/vocoder/hifigan/hifigan_causal.py:272-312`)
and init distributions of the same family as the reference constructors
(xavier / kaiming / torch default conv init, weight-norm stored un-folded as
weight_g / weight_v).  Every tensor is drawn from its own generator seeded by
crc32(key) ^ seed, so a tensor's values do not depend on construction order and
are bit-identical wherever the same torch build runs.

Deliberate departures from a literal random init, so that the discrete branches
of the path are exercised (SURVEY.md section 7 "hard parts"):
  * LayerNorm affine is 1 + 0.1 N / 0.05 N instead of 1 / 0,
  * the VQ codebook is N(0, 0.5) instead of U(+-1/512) (argmin margins are not ties),
  * `uv_predictor.linear` is scaled/biased so that log2-f0 lands in ~[6, 9] and
    about half the frames are voiced (f0 bucket + pitch-embed gather are used).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

# --------------------------------------------------------------------------
# default hyper-parameters of the hot path (egs/conan_emformer.yaml chain and
# egs/hifi_16k320_shuffle.yaml chain of the reference)
# --------------------------------------------------------------------------
DEFAULT_HP = dict(
    hidden_size=256, kernel_size=3, dec_dilations=[1, 1, 1, 1], dec_kernel_size=5,
    layers_in_block=2, dec_post_net_kernel=3, predictor_kernel=5, nVQ=512,
    audio_num_mel_bins=80, emformer_layers=6, chunk_size=80, right_context=2,
    emformer_output_dim=100, silent_token=57,
)
DEFAULT_VOC_HP = dict(
    upsample_rates=[8, 5, 4, 2], upsample_kernel_sizes=[16, 10, 8, 4],
    upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80,
)

Spec = List[Tuple[str, Tuple[int, ...], str]]


def _conv_ln_blocks(spec: Spec, prefix: str, ch: int, k: int, n_blocks: int, n_sub: int,
                    conv_idx: int, pw_idx: int, out_dims: int, post_key: str, post_k: int):
    for b in range(n_blocks):
        for s in range(n_sub):
            p = f"{prefix}.res_blocks.{b}.blocks.{s}"
            spec += [(f"{p}.0.weight", (ch,), "ln_w"), (f"{p}.0.bias", (ch,), "ln_b"),
                     (f"{p}.{conv_idx}.weight", (2 * ch, ch, k), "xavier"),
                     (f"{p}.{conv_idx}.bias", (2 * ch,), "bias"),
                     (f"{p}.{pw_idx}.weight", (ch, 2 * ch, 1), "xavier"),
                     (f"{p}.{pw_idx}.bias", (ch,), "bias")]
    spec += [(f"{prefix}.last_norm.weight", (ch,), "ln_w"), (f"{prefix}.last_norm.bias", (ch,), "ln_b"),
             (f"{post_key}.weight", (out_dims, ch, post_k), "xavier"), (f"{post_key}.bias", (out_dims,), "bias")]


def conan_spec(hp: Dict = None) -> Spec:
    hp = {**DEFAULT_HP, **(hp or {})}
    H = hp["hidden_size"]
    spec: Spec = []
    # decoder: CausalConvBlocks (modules/commons/conv.py:181-264) -> indices 0 LN, 2 conv, 5 1x1
    _conv_ln_blocks(spec, "decoder", H, hp["dec_kernel_size"], len(hp["dec_dilations"]), hp["layers_in_block"],
                    2, 5, H, "decoder.post_net1.1", hp["dec_post_net_kernel"])
    spec += [("mel_out.weight", (hp["audio_num_mel_bins"], H), "linear"), ("mel_out.bias", (hp["audio_num_mel_bins"],), "bias"),
             ("pitch_embed.weight", (300, H), "embed_pad0")]
    # FastSpeech.pitch_predictor: built by the ctor, never used at inference (modules/tts/fs.py:73)
    for i in range(5):
        spec += [(f"pitch_predictor.conv.{i}.0.conv.weight", (H, H, hp["predictor_kernel"]), "kaiming"),
                 (f"pitch_predictor.conv.{i}.0.conv.bias", (H,), "zeros")]
    spec += [("pitch_predictor.post_ln.weight", (H,), "ln_w"), ("pitch_predictor.post_ln.bias", (H,), "ln_b"),
             ("pitch_predictor.linear.weight", (2, H), "linear"), ("pitch_predictor.linear.bias", (2,), "bias")]
    spec += [("content_embedding.weight", (102, H), "normal1"),
             ("content_proj.0.conv.weight", (H, H, hp["kernel_size"]), "kaiming"),
             ("content_proj.0.conv.bias", (H,), "bias"),
             ("global_conv_in.weight", (H, 80, 1), "default"), ("global_conv_in.bias", (H,), "bias")]
    # global_encoder: ConvBlocks k31, 5 blocks x 2 (modules/Conan/Conan.py:62-70) -> 0 LN, 1 conv, 4 1x1
    _conv_ln_blocks(spec, "global_encoder", H, 31, 5, 2, 1, 4, H, "global_encoder.post_net1", 3)
    # prosody_extractor.encoder: local ConvBlocks(80, H, [1]*5, 5) (modules/Conan/prosody_util.py:176)
    _conv_ln_blocks(spec, "prosody_extractor.encoder", 80, 5, 5, 2, 1, 4, H, "prosody_extractor.encoder.post_net1", 3)
    spec += [("prosody_extractor.vqvae.data_initialized", (1,), "ones"),
             ("prosody_extractor.vqvae.embedding", (hp["nVQ"], H), "codebook"),
             ("prosody_extractor.vqvae.ema_count", (hp["nVQ"],), "zeros"),
             ("prosody_extractor.vqvae.ema_weight", (hp["nVQ"], H), "codebook")]
    for i in range(4):
        spec += [(f"prosody_extractor.wavenet.in_layers.{i}.bias", (160,), "bias"),
                 (f"prosody_extractor.wavenet.in_layers.{i}.weight_g", (160, 1, 1), f"wn_g:prosody_extractor.wavenet.in_layers.{i}.weight_v"),
                 (f"prosody_extractor.wavenet.in_layers.{i}.weight_v", (160, 80, 3), "default")]
    for i in range(4):
        co = 160 if i < 3 else 80
        spec += [(f"prosody_extractor.wavenet.res_skip_layers.{i}.bias", (co,), "bias"),
                 (f"prosody_extractor.wavenet.res_skip_layers.{i}.weight_g", (co, 1, 1), f"wn_g:prosody_extractor.wavenet.res_skip_layers.{i}.weight_v"),
                 (f"prosody_extractor.wavenet.res_skip_layers.{i}.weight_v", (co, 80, 1), "default")]
    spec += [("l1.weight", (H, 2 * H), "linear"), ("l1.bias", (H,), "bias")]
    for i in range(2):
        p = f"align.layers.{i}"
        spec += [(f"{p}.multihead_attn.in_proj_weight", (3 * H, H), "xavier"),
                 (f"{p}.multihead_attn.in_proj_bias", (3 * H,), "bias"),
                 (f"{p}.multihead_attn.out_proj.weight", (H, H), "linear"),
                 (f"{p}.multihead_attn.out_proj.bias", (H,), "bias"),
                 (f"{p}.linear1.weight", (2048, H), "linear"), (f"{p}.linear1.bias", (2048,), "bias"),
                 (f"{p}.norm1.weight", (H,), "ln_w"), (f"{p}.norm1.bias", (H,), "ln_b"),
                 (f"{p}.linear2.weight", (H, 2048), "linear"), (f"{p}.linear2.bias", (H,), "bias"),
                 (f"{p}.norm2.weight", (H,), "ln_w"), (f"{p}.norm2.bias", (H,), "ln_b")]
    spec += [("embed_positions._float_tensor", (1,), "zeros")]
    for i in range(5):
        ci = H if i == 0 else 128
        spec += [(f"uv_predictor.conv.{i}.0.conv.weight", (128, ci, hp["predictor_kernel"]), "kaiming"),
                 (f"uv_predictor.conv.{i}.0.conv.bias", (128,), "bias")]
    spec += [("uv_predictor.post_ln.weight", (128,), "ln_w"), ("uv_predictor.post_ln.bias", (128,), "ln_b"),
             ("uv_predictor.linear.weight", (2, 128), "uv_linear_w"), ("uv_predictor.linear.bias", (2,), "uv_linear_b")]
    return spec


def emformer_spec(hp: Dict = None) -> Spec:
    hp = {**DEFAULT_HP, **(hp or {})}
    D, F = 80, 2048
    spec: Spec = []
    for i in range(hp["emformer_layers"]):
        p = f"emformer.emformer_layers.{i}"
        spec += [(f"{p}.attention.emb_to_key_value.weight", (2 * D, D), "xavier"),
                 (f"{p}.attention.emb_to_key_value.bias", (2 * D,), "bias"),
                 (f"{p}.attention.emb_to_query.weight", (D, D), "xavier"),
                 (f"{p}.attention.emb_to_query.bias", (D,), "bias"),
                 (f"{p}.attention.out_proj.weight", (D, D), "linear"),
                 (f"{p}.attention.out_proj.bias", (D,), "bias"),
                 (f"{p}.pos_ff.0.weight", (D,), "ln_w"), (f"{p}.pos_ff.0.bias", (D,), "ln_b"),
                 (f"{p}.pos_ff.1.weight", (F, D), "linear"), (f"{p}.pos_ff.1.bias", (F,), "bias"),
                 (f"{p}.pos_ff.4.weight", (D, F), "linear"), (f"{p}.pos_ff.4.bias", (D,), "bias"),
                 (f"{p}.layer_norm_input.weight", (D,), "ln_w"), (f"{p}.layer_norm_input.bias", (D,), "ln_b"),
                 (f"{p}.layer_norm_output.weight", (D,), "ln_w"), (f"{p}.layer_norm_output.bias", (D,), "ln_b")]
    spec += [("proj.weight", (hp["emformer_output_dim"], D), "linear"), ("proj.bias", (hp["emformer_output_dim"],), "bias")]
    return spec


def hifigan_spec(voc_hp: Dict = None) -> Spec:
    vh = {**DEFAULT_VOC_HP, **{k: v for k, v in (voc_hp or {}).items() if v is not None}}
    spec: Spec = []

    def wn(prefix, co, ci, k):
        spec.extend([(f"{prefix}.bias", (co,), "bias"),
                     (f"{prefix}.weight_g", (co, 1, 1), f"wn_g2:{prefix}.weight_v"),
                     (f"{prefix}.weight_v", (co, ci, k), "default")])

    ch = vh["upsample_initial_channel"]
    wn("conv_pre.conv", ch, vh["num_mels"], 7)
    for i, (u, k) in enumerate(zip(vh["upsample_rates"], vh["upsample_kernel_sizes"])):
        wn(f"ups.{i}.conv.conv", (ch // 2) * u, ch, k)
        ch //= 2
    ch = vh["upsample_initial_channel"]
    rb = 0
    for i in range(len(vh["upsample_rates"])):
        ch //= 2
        for k, dils in zip(vh["resblock_kernel_sizes"], vh["resblock_dilation_sizes"]):
            for j in range(len(dils)):
                wn(f"resblocks.{rb}.convs1.{j}.conv", ch, ch, k)
            for j in range(len(dils)):
                wn(f"resblocks.{rb}.convs2.{j}.conv", ch, ch, k)
            rb += 1
    wn("conv_post.conv", 1, ch, 7)
    return spec


# --------------------------------------------------------------------------
def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _fans(shape):
    rf = 1
    for s in shape[2:]:
        rf *= s
    return shape[1] * rf, shape[0] * rf


def make_state_dict(spec: Spec, seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    deferred = []
    for key, shape, kind in spec:
        g = _gen(key, seed)
        if kind.startswith("wn_g:") or kind.startswith("wn_g2:"):
            # wn_g2 = vocoder: gain 2 keeps activations O(0.4) through all four scales so the
            # synthetic waveform is not a DC offset (measured: wav std 0.18 instead of 0.007)
            deferred.append((key, shape, kind.split(":", 1)[1], 2.0 if kind.startswith("wn_g2:") else 1.0))
            sd[key] = None
            continue
        if kind == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "xavier":
            fi, fo = _fans(shape)
            b = math.sqrt(6.0 / (fi + fo))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "kaiming":
            fi, _ = _fans(shape)
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fi)
        elif kind in ("default", "linear"):
            fi, _ = _fans(shape)
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fi)
        elif kind == "normal1":
            t = torch.randn(shape, generator=g)
        elif kind == "embed_pad0":
            t = torch.randn(shape, generator=g) * shape[1] ** -0.5
            t[0] = 0
        elif kind == "codebook":
            # ema_weight mirrors embedding in the reference ctor (prosody_util.py:28-31)
            t = torch.randn(shape, generator=_gen("prosody_extractor.vqvae.embedding", seed)) * 0.5
        elif kind == "uv_linear_w":
            t = torch.randn(shape, generator=g) * torch.tensor([[0.09], [0.06]])
        elif kind == "uv_linear_b":
            t = torch.tensor([0.0, 7.5])
        else:
            raise ValueError(kind)
        sd[key] = t.float().contiguous()
    for key, shape, vkey, gain in deferred:
        v = sd[vkey]
        g = _gen(key, seed)
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        sd[key] = (nrm * gain * (0.9 + 0.2 * torch.rand(shape, generator=g))).float().contiguous()
    return sd


def make_all_state_dicts(seed: int = 1234, hp: Dict = None, voc_hp: Dict = None):
    """-> (conan_sd, emformer_sd, hifigan_sd) with the reference's key names / shapes."""
    return (make_state_dict(conan_spec(hp), seed),
            make_state_dict(emformer_spec(hp), seed),
            make_state_dict(hifigan_spec(voc_hp), seed))


def state_dict_checksum(sd) -> float:
    """Order-independent fingerprint used to pin fixtures to the generator output."""
    tot = 0.0
    for k in sorted(sd.keys()):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float(v.abs().sum()) + 1e-3 * (zlib.crc32(k.encode()) % 997)
    return tot


# --------------------------------------------------------------------------
# synthetic inputs: log-mel ~ clip(N(-3, 1.5^2), -6, 1.5)
# --------------------------------------------------------------------------
def synth_mel(n_frames: int, seed: int, smooth: bool = True) -> torch.Tensor:
    """[n_frames, 80] float32, values in [mel_vmin, mel_vmax] = [-6, 1.5]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1) & 0x7FFFFFFF)
    x = torch.randn(n_frames, 80, generator=g)
    if smooth and n_frames > 2:
        # mild temporal correlation so consecutive frames resemble speech envelopes
        x[1:] = 0.6 * x[1:] + 0.4 * x[:-1]
    return (x * 1.5 - 3.0).clamp_(-6.0, 1.5).contiguous()
