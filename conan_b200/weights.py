"""Host-side weight packing: reference state_dicts -> the tensors the engine binds by name.

Layouts the kernels expect:
  * every Conv1d weight [Cout, Cin, k] is packed K-major and tap-major as [Cout, k*Cin]
    (W[n, j*Cin + c] = w[n, c, j]) so one K-slice of the implicit GEMM is a contiguous run of
    input channels of one tap;
  * weight-norm convolutions are folded (w = v * g / ||v||) at load time;
  * the pixel-shuffle of the vocoder's upsampling blocks is folded into the row order of the
    conv weight: new row j*C + c <- old row c*r + j, so the conv output [t, r*C] *is*
    out[t*r + j, c] of CausalPixelShuffle1d (hifigan_causal.py:186-188) without a shuffle kernel;
  * vocoder conv weights are rounded to fp16 (RN) when voc_precision = 1.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .ckpt import fold_weight_norm


def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, k] -> [Cout, k*Cin] (tap-major K)."""
    return w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()


SPLIT_WEIGHT_SCALE = 1024.0      # must match kSplitWeightScale in csrc/engine.cu


def split_linear(w_packed: torch.Tensor, bias: torch.Tensor, k: int = 1):
    """fp32 packed weight [N, k*K] -> the operand of the fp32-grade tensor-core GEMM:
    fp16 [Npad, 3*k*Kpad] = [W_hi | W_lo | W_hi] of 2^10 * W (W_hi = fp16(W'), W_lo = fp16(W' - W_hi)), N and K
    zero-padded to multiples of 32, plus the bias zero-padded to Npad.  The kernel accumulates
    x_hi*W_hi + x_hi*W_lo + x_lo*W_hi in fp32 and multiplies by 2^-10."""
    N, KK = w_packed.shape
    K = KK // k
    Np, Kp = (N + 31) // 32 * 32, (K + 31) // 32 * 32
    w = torch.zeros(Np, k, Kp, dtype=torch.float32)
    w[:N, :, :K] = w_packed.view(N, k, K) * SPLIT_WEIGHT_SCALE
    w = w.view(Np, k * Kp)
    hi = w.half()
    lo = (w - hi.float()).half()
    b = torch.zeros(Np, dtype=torch.float32)
    b[:N] = bias
    return torch.cat([hi, lo, hi], dim=1).contiguous(), b


def sinusoid_table(n_pos: int, dim: int, padding_idx: int = 0) -> torch.Tensor:
    """Sinusoidal position table laid out as the reference builds it
    (modules/commons/transformer.py:31-47): [sin | cos] halves, log(10000)/(half-1) spacing,
    row `padding_idx` zero."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
    ang = torch.arange(n_pos, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
    tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(n_pos, -1)
    tab[padding_idx, :] = 0
    return tab


def pack_engine_weights(sd_conan: Dict[str, torch.Tensor], sd_emf: Dict[str, torch.Tensor],
                        sd_voc: Dict[str, torch.Tensor], cfg) -> Dict[str, torch.Tensor]:
    """cfg: a _lib.ConanConfig.  Returns {engine weight name: CPU tensor (fp32 or fp16)}."""
    out: Dict[str, torch.Tensor] = {}
    H = cfg.hidden_size
    # ---- Emformer (TA:105-115, 357-374)
    for l in range(cfg.emformer_layers):
        s = f"emformer.emformer_layers.{l}."
        d = f"emf.{l}."
        out[d + "ln_in.g"], out[d + "ln_in.b"] = sd_emf[s + "layer_norm_input.weight"], sd_emf[s + "layer_norm_input.bias"]
        out[d + "qkv.w"] = torch.cat([sd_emf[s + "attention.emb_to_query.weight"], sd_emf[s + "attention.emb_to_key_value.weight"]], 0)
        out[d + "qkv.b"] = torch.cat([sd_emf[s + "attention.emb_to_query.bias"], sd_emf[s + "attention.emb_to_key_value.bias"]], 0)
        out[d + "out.w"], out[d + "out.b"] = sd_emf[s + "attention.out_proj.weight"], sd_emf[s + "attention.out_proj.bias"]
        out[d + "ffn_ln.g"], out[d + "ffn_ln.b"] = sd_emf[s + "pos_ff.0.weight"], sd_emf[s + "pos_ff.0.bias"]
        out[d + "ffn1.w"], out[d + "ffn1.b"] = sd_emf[s + "pos_ff.1.weight"], sd_emf[s + "pos_ff.1.bias"]
        out[d + "ffn2.w"], out[d + "ffn2.b"] = sd_emf[s + "pos_ff.4.weight"], sd_emf[s + "pos_ff.4.bias"]
        out[d + "ln_out.g"], out[d + "ln_out.b"] = sd_emf[s + "layer_norm_output.weight"], sd_emf[s + "layer_norm_output.bias"]
    out["emf.proj.w"], out["emf.proj.b"] = sd_emf["proj.weight"], sd_emf["proj.bias"]
    # ---- Conan chunk path
    c = sd_conan
    out["conan.content_embedding"] = c["content_embedding.weight"]
    out["conan.content_proj.w"], out["conan.content_proj.b"] = pack_conv(c["content_proj.0.conv.weight"]), c["content_proj.0.conv.bias"]
    for l in range(2):
        s, d = f"align.layers.{l}.", f"conan.align.{l}."
        W, b = c[s + "multihead_attn.in_proj_weight"], c[s + "multihead_attn.in_proj_bias"]
        out[d + "q.w"], out[d + "q.b"] = W[:H], b[:H]
        out[d + "kv.w"], out[d + "kv.b"] = W[H:], b[H:]
        out[d + "out.w"], out[d + "out.b"] = c[s + "multihead_attn.out_proj.weight"], c[s + "multihead_attn.out_proj.bias"]
        out[d + "norm1.g"], out[d + "norm1.b"] = c[s + "norm1.weight"], c[s + "norm1.bias"]
        out[d + "ffn1.w"], out[d + "ffn1.b"] = c[s + "linear1.weight"], c[s + "linear1.bias"]
        out[d + "ffn2.w"], out[d + "ffn2.b"] = c[s + "linear2.weight"], c[s + "linear2.bias"]
        out[d + "norm2.g"], out[d + "norm2.b"] = c[s + "norm2.weight"], c[s + "norm2.bias"]
    for i in range(5):
        out[f"conan.uv.{i}.w"] = pack_conv(c[f"uv_predictor.conv.{i}.0.conv.weight"])
        out[f"conan.uv.{i}.b"] = c[f"uv_predictor.conv.{i}.0.conv.bias"]
    out["conan.uv.ln.g"], out["conan.uv.ln.b"] = c["uv_predictor.post_ln.weight"], c["uv_predictor.post_ln.bias"]
    out["conan.uv.lin.w"], out["conan.uv.lin.b"] = c["uv_predictor.linear.weight"], c["uv_predictor.linear.bias"]
    out["conan.pitch_embed"] = c["pitch_embed.weight"]
    for b_ in range(cfg.dec_blocks):
        for s_ in range(2):
            s, d = f"decoder.res_blocks.{b_}.blocks.{s_}.", f"conan.dec.{b_}.{s_}."
            out[d + "ln.g"], out[d + "ln.b"] = c[s + "0.weight"], c[s + "0.bias"]
            out[d + "conv.w"], out[d + "conv.b"] = pack_conv(c[s + "2.weight"]), c[s + "2.bias"]
            out[d + "pw.w"], out[d + "pw.b"] = pack_conv(c[s + "5.weight"]), c[s + "5.bias"]
    out["conan.dec.last_norm.g"], out["conan.dec.last_norm.b"] = c["decoder.last_norm.weight"], c["decoder.last_norm.bias"]
    out["conan.dec.post.w"], out["conan.dec.post.b"] = pack_conv(c["decoder.post_net1.1.weight"]), c["decoder.post_net1.1.bias"]
    out["conan.mel_out.w"], out["conan.mel_out.b"] = c["mel_out.weight"], c["mel_out.bias"]
    # ---- session-setup branch
    out["conan.global_in.w"], out["conan.global_in.b"] = pack_conv(c["global_conv_in.weight"]), c["global_conv_in.bias"]
    for b_ in range(5):
        for s_ in range(2):
            for src, dst in (("global_encoder", "conan.genc"), ("prosody_extractor.encoder", "conan.penc")):
                s, d = f"{src}.res_blocks.{b_}.blocks.{s_}.", f"{dst}.{b_}.{s_}."
                out[d + "ln.g"], out[d + "ln.b"] = c[s + "0.weight"], c[s + "0.bias"]
                out[d + "conv.w"], out[d + "conv.b"] = pack_conv(c[s + "1.weight"]), c[s + "1.bias"]
                out[d + "pw.w"], out[d + "pw.b"] = pack_conv(c[s + "4.weight"]), c[s + "4.bias"]
    for src, dst in (("global_encoder", "conan.genc"), ("prosody_extractor.encoder", "conan.penc")):
        out[dst + ".last_norm.g"], out[dst + ".last_norm.b"] = c[src + ".last_norm.weight"], c[src + ".last_norm.bias"]
        out[dst + ".post.w"], out[dst + ".post.b"] = pack_conv(c[src + ".post_net1.weight"]), c[src + ".post_net1.bias"]
    for i in range(4):
        pi, pr = f"prosody_extractor.wavenet.in_layers.{i}", f"prosody_extractor.wavenet.res_skip_layers.{i}"
        out[f"conan.wn.{i}.in.w"], out[f"conan.wn.{i}.in.b"] = pack_conv(fold_weight_norm(c, pi)), c[pi + ".bias"]
        out[f"conan.wn.{i}.rs.w"], out[f"conan.wn.{i}.rs.b"] = pack_conv(fold_weight_norm(c, pr)), c[pr + ".bias"]
    E = c["prosody_extractor.vqvae.embedding"]
    out["conan.vq.embedding"] = E
    out["conan.vq.e2"] = torch.sum(E ** 2, dim=1)
    tp_max = (cfg.max_ref_frames - 1) // 4 + 1
    out["conan.pos_table"] = sinusoid_table(tp_max + 1, H, 0)
    out["conan.l1.w"], out["conan.l1.b"] = c["l1.weight"], c["l1.bias"]
    # ---- vocoder
    v = sd_voc
    wdt = torch.float16 if cfg.voc_precision == 1 else torch.float32      # 2 (split): packed in fp32 here, split below
    w_pre = fold_weight_norm(v, "conv_pre.conv")
    if cfg.voc_use_tensor_cores:      # mel channels padded to a multiple of 32 (zero weights): conv_pre runs on the tcgen05 ring kernel
        pad = (-w_pre.shape[1]) % 32
        w_pre = torch.nn.functional.pad(w_pre, (0, 0, 0, pad))
    out["voc.pre.w"], out["voc.pre.b"] = pack_conv(w_pre).to(wdt), v["conv_pre.conv.bias"]
    ch = cfg.voc_initial_channel
    rb = 0
    for i in range(cfg.voc_n_ups):
        r, co = cfg.voc_rates[i], ch // 2
        w = fold_weight_norm(v, f"ups.{i}.conv.conv")                           # [co*r, ch, k], row index c*r + j
        w = w.view(co, r, w.shape[1], w.shape[2]).permute(1, 0, 2, 3).reshape(r * co, w.shape[1], w.shape[2])
        out[f"voc.up.{i}.w"] = pack_conv(w).to(wdt)
        out[f"voc.up.{i}.b"] = v[f"ups.{i}.conv.conv.bias"].view(co, r).t().reshape(-1)
        for rr in range(cfg.voc_n_res):
            for j in range(cfg.voc_n_dil):
                for which in ("c1", "c2"):
                    src = f"resblocks.{rb}.convs{which[1]}.{j}.conv"
                    out[f"voc.res.{i}.{rr}.{which}.{j}.w"] = pack_conv(fold_weight_norm(v, src)).to(wdt)
                    out[f"voc.res.{i}.{rr}.{which}.{j}.b"] = v[src + ".bias"]
            rb += 1
        ch = co
    wpost = fold_weight_norm(v, "conv_post.conv")                                # [1, ch, 7]
    out["voc.post.w"] = wpost[0].t().contiguous().reshape(-1)                    # [7, ch]
    out["voc.post.b"] = v["conv_post.conv.bias"]
    if cfg.voc_precision == 2:
        # fp32-grade tensor-core vocoder: every conv weight becomes the split-fp16 operand [W_hi | W_lo | W_hi] of 2^10 W
        taps = {"voc.pre": 7}
        for i in range(cfg.voc_n_ups):
            taps[f"voc.up.{i}"] = cfg.voc_up_kernels[i]
            for rr in range(cfg.voc_n_res):
                for j in range(cfg.voc_n_dil):
                    taps[f"voc.res.{i}.{rr}.c1.{j}"] = taps[f"voc.res.{i}.{rr}.c2.{j}"] = cfg.voc_res_kernels[rr]
        for n, k in taps.items():
            out[n + ".w"], out[n + ".b"] = split_linear(out[n + ".w"].float(), out[n + ".b"].float(), k)
    # ---- fp32-grade tensor-core mode: the per-chunk linear / conv contractions take split-fp16 weights
    if cfg.lin_use_tensor_cores and cfg.lin_fuse_ffn:
        # fused Emformer path: the attention kernel applies out_proj itself (fp32, 80 x 80): transposed weight [k][c] + bias
        for l in range(cfg.emformer_layers):
            out[f"emf.{l}.out.wt"] = out[f"emf.{l}.out.w"].float().t().contiguous()
            out[f"emf.{l}.out.bf"] = out[f"emf.{l}.out.b"].float().contiguous()
    if cfg.lin_use_tensor_cores:
        names = [f"emf.{l}.{n}" for l in range(cfg.emformer_layers) for n in ("qkv", "out", "ffn1", "ffn2")] + ["emf.proj"]
        names += [f"conan.align.{l}.{n}" for l in range(2) for n in ("q", "out", "ffn1", "ffn2")]
        names += [f"conan.dec.{b_}.{s_}.pw" for b_ in range(cfg.dec_blocks) for s_ in range(2)]
        taps = {n: 1 for n in names}
        taps["conan.content_proj"] = cfg.content_kernel
        taps["conan.dec.post"] = cfg.dec_post_kernel
        taps["conan.mel_out"] = 1
        taps.update({f"conan.uv.{i}": cfg.predictor_kernel for i in range(5)})
        taps.update({f"conan.dec.{b_}.{s_}.conv": cfg.dec_kernel for b_ in range(cfg.dec_blocks) for s_ in range(2)})
        for n, k in taps.items():
            out[n + ".w"], out[n + ".b"] = split_linear(out[n + ".w"].float(), out[n + ".b"].float(), k)
    if cfg.ses_use_tensor_cores:
        ses = {f"conan.genc.{b_}.{s_}.conv": 31 for b_ in range(5) for s_ in range(2)}
        ses.update({f"conan.genc.{b_}.{s_}.pw": 1 for b_ in range(5) for s_ in range(2)})
        ses["conan.genc.post"] = 3
        for n, k in ses.items():
            out[n + ".w"], out[n + ".b"] = split_linear(out[n + ".w"].float(), out[n + ".b"].float(), k)
    return {k: t.detach().contiguous() for k, t in out.items()}
