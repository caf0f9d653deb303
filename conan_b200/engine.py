"""Python handle on the native engine (libconan_b200.so).

PyTorch is used for device memory, streams and host<->device copies only; every op on
the path is a hand-written CUDA kernel behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .synth import DEFAULT_HP, DEFAULT_VOC_HP
from .weights import pack_engine_weights

PARTS_EMFORMER, PARTS_CONAN, PARTS_VOCODER, PARTS_ALL = 1, 2, 4, 7


def make_config(hp: Optional[Dict] = None, voc_hp: Optional[Dict] = None, *, max_slots: int = 64,
                max_ref_frames: int = 512, device: int = 0, voc_precision: str = "fp16",
                voc_tensor_cores: bool = True, voc_group: int = 0, lin_tensor_cores: Optional[bool] = None,
                voc_residual_from_ctx: Optional[bool] = None, voc_fuse_resblocks: Optional[bool] = None,
                lin_fuse_ffn: Optional[bool] = None, ses_tensor_cores: Optional[bool] = None,
                emformer_memory_size: Optional[int] = None, step_graphs: bool = True,
                lin_fuse_blocks: Optional[bool] = None) -> _lib.ConanConfig:
    """Builds the native config from reference-style hparams dicts (the keys the hot path reads,
    SURVEY.md section 5)."""
    hp = {**DEFAULT_HP, **{k: v for k, v in (hp or {}).items() if v is not None}}
    vh = {**DEFAULT_VOC_HP, **{k: v for k, v in (voc_hp or {}).items() if v is not None}}
    if vh.get("upsample", "shuffle") != "shuffle":
        raise ValueError("only upsample: 'shuffle' (CausalUpsampleBlock3) is on the hot path")
    if str(vh.get("resblock", "1")) != "1":
        raise ValueError("only resblock: '1' is on the hot path")
    if hp.get("decoder_type", "conv") != "conv" or hp.get("f0_gen", "orig") != "orig" or not hp.get("style", True):
        raise ValueError("config outside the hot path (decoder_type conv, f0_gen orig, style true)")
    dil = vh["resblock_dilation_sizes"]
    if any(list(d) != list(dil[0]) for d in dil):
        raise ValueError("resblock_dilation_sizes must be the same for every kernel size")
    cfg = _lib.ConanConfig()
    cfg.abi_version = _lib.ABI_VERSION
    cfg.device, cfg.max_slots, cfg.max_ref_frames = device, max_slots, max_ref_frames
    cfg.emformer_layers, cfg.emformer_dim, cfg.emformer_heads, cfg.emformer_ffn = hp["emformer_layers"], 80, 8, 2048
    cfg.segment, cfg.right_context, cfg.left_context = hp["chunk_size"] // 20, hp["right_context"], 50
    cfg.emformer_output_dim = hp["emformer_output_dim"]
    cfg.hidden_size, cfg.content_kernel = hp["hidden_size"], hp["kernel_size"]
    cfg.dec_blocks, cfg.dec_kernel = len(hp["dec_dilations"]), hp["dec_kernel_size"]
    if any(d != 1 for d in hp["dec_dilations"]) or hp["layers_in_block"] != 2:
        raise ValueError("decoder dilations must be 1 and layers_in_block 2")
    cfg.dec_post_kernel, cfg.predictor_kernel = hp["dec_post_net_kernel"], hp["predictor_kernel"]
    cfg.n_vq, cfg.silent_token, cfg.n_mels = hp["nVQ"], hp["silent_token"], hp["audio_num_mel_bins"]
    cfg.voc_initial_channel = vh["upsample_initial_channel"]
    cfg.voc_n_ups = len(vh["upsample_rates"])
    for i, (r, k) in enumerate(zip(vh["upsample_rates"], vh["upsample_kernel_sizes"])):
        cfg.voc_rates[i], cfg.voc_up_kernels[i] = r, k
    cfg.voc_n_res = len(vh["resblock_kernel_sizes"])
    for i, k in enumerate(vh["resblock_kernel_sizes"]):
        cfg.voc_res_kernels[i] = k
    cfg.voc_n_dil = len(dil[0])
    for i, d in enumerate(dil[0]):
        cfg.voc_res_dilations[i] = d
    # "fp32": fp32 operands on the CUDA cores (exact-fp32 cross-check engine); "fp16": fp16 operands / fp32 accumulate on
    # tcgen05 (the fast mode); "split": split-fp16 operands (x_hi W_hi + x_hi W_lo + x_lo W_hi, fp32 accumulate) on tcgen05 --
    # fp32-grade results at a third of the fp16 mode's tensor rate
    cfg.voc_precision = {"fp32": 0, "fp16": 1, "split": 2}[voc_precision]
    if cfg.voc_precision == 2 and not voc_tensor_cores:
        raise ValueError("voc_precision='split' is a tensor-core mode")
    cfg.voc_use_tensor_cores = int(bool(voc_tensor_cores) and cfg.voc_precision >= 1)
    cfg.voc_group = voc_group
    # fp16-operand vocoder: recover the resblock residual from the activated fp16 context rows (halves the HBM traffic of
    # every second conv; costs 1.7 dB of the 59 dB SNR).  The fp32 vocoder keeps its fp32 residual stream.
    cfg.voc_residual_from_ctx = int(cfg.voc_precision == 1 if voc_residual_from_ctx is None else bool(voc_residual_from_ctx))
    # Emformer / Conan contractions: split-fp16 tensor-core GEMMs (fp32-grade) by default whenever the tensor-core
    # vocoder is on; lin_tensor_cores=False keeps them on the exact-fp32 FFMA engine
    cfg.lin_use_tensor_cores = int(cfg.voc_use_tensor_cores if lin_tensor_cores is None else bool(lin_tensor_cores))
    # whole residual blocks as one kernel at the 32 / 64 channel scales (activations stay in shared memory)
    if cfg.voc_precision == 2:
        cfg.voc_residual_from_ctx = 0      # fp32 residual stream next to the split operands
    can_fuse = bool(cfg.voc_use_tensor_cores and cfg.voc_residual_from_ctx and cfg.voc_precision == 1)
    cfg.voc_fuse_resblocks = int(can_fuse if voc_fuse_resblocks is None else (bool(voc_fuse_resblocks) and can_fuse))
    cfg.lin_fuse_ffn = int(bool(cfg.lin_use_tensor_cores) if lin_fuse_ffn is None else (bool(lin_fuse_ffn) and bool(cfg.lin_use_tensor_cores)))
    # session setup: the style encoder (95 % of the setup FLOPs) on the tensor cores, same split-fp16 operand format
    cfg.ses_use_tensor_cores = int(bool(cfg.lin_use_tensor_cores) if ses_tensor_cores is None else bool(ses_tensor_cores))
    # torchaudio Emformer max_memory_size: the reference never passes it (modules/Emformer/emformer.py:14-22 -> 0); an
    # `emformer_memory_size` hparam / argument enables the memory bank for checkpoints trained with one
    cfg.emformer_memory_size = int(hp.get("emformer_memory_size", 0) if emformer_memory_size is None else emformer_memory_size)
    cfg.step_graphs = int(bool(step_graphs))
    cfg.lin_fuse_blocks = int(bool(cfg.lin_use_tensor_cores) if lin_fuse_blocks is None else (bool(lin_fuse_blocks) and bool(cfg.lin_use_tensor_cores)))
    return cfg


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_PACK_CACHE: Dict[tuple, Dict[str, torch.Tensor]] = {}


def _packed_weights(sd_conan, sd_emformer, sd_voc, cfg):
    """pack_engine_weights memoised on the state-dict objects and the config fields the packing depends on (a process that
    builds several engines over the same checkpoints -- bench.py's sweeps, the test suite -- packs them once)."""
    key = (id(sd_conan), id(sd_emformer), id(sd_voc), cfg.voc_precision, cfg.voc_use_tensor_cores, cfg.lin_use_tensor_cores,
           cfg.lin_fuse_ffn, cfg.ses_use_tensor_cores, cfg.lin_fuse_blocks, cfg.max_ref_frames, cfg.emformer_layers, cfg.right_context)
    hit = _PACK_CACHE.get(key)
    if hit is None or hit[0] is not sd_conan:
        if len(_PACK_CACHE) >= 4:
            _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
        hit = (sd_conan, pack_engine_weights(sd_conan, sd_emformer, sd_voc, cfg))
        _PACK_CACHE[key] = hit
    return hit[1]


class Engine:
    """Owns the native engine, the bound weight tensors and the resident per-slot state."""

    def __init__(self, sd_conan, sd_emformer, sd_voc, cfg: _lib.ConanConfig):
        if not torch.cuda.is_available():
            raise RuntimeError("conan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", cfg.device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        _lib.check(self.lib.conan_engine_create(C.byref(cfg), C.byref(h)), "engine_create")
        self.h = h
        packed = _packed_weights(sd_conan, sd_emformer, sd_voc, cfg)
        self._weights = {}
        nw = self.lib.conan_engine_num_weights(self.h)
        name, numel, dtype = C.c_char_p(), C.c_size_t(), C.c_int()
        for i in range(nw):
            _lib.check(self.lib.conan_engine_weight_info(self.h, i, C.byref(name), C.byref(numel), C.byref(dtype)), "weight_info")
            key = name.value.decode()
            if key not in packed:
                raise KeyError(f"packer produced no tensor for engine weight '{key}'")
            want = {0: torch.float32, 1: torch.float16, 2: torch.int32}[dtype.value]
            t = packed[key].to(want).contiguous().to(self.device)
            self._weights[key] = t
            _lib.check(self.lib.conan_engine_bind_weight(self.h, key.encode(), _ptr(t), t.numel(), dtype.value), f"bind {key}")
        _lib.check(self.lib.conan_engine_finalize(self.h), "finalize")
        # fp32 copy of the content-token projection for the module-level `emformer.proj` view
        self.aux = {"emf.proj.w": sd_emformer["proj.weight"].float().contiguous().to(self.device),
                    "emf.proj.b": sd_emformer["proj.bias"].float().contiguous().to(self.device)} if "proj.weight" in sd_emformer else {}
        self.segment = cfg.segment
        self.rows_in = cfg.segment + cfg.right_context
        self.hop_out = cfg.segment
        for i in range(cfg.voc_n_ups):
            self.hop_out *= cfg.voc_rates[i]
        self.n_mels = cfg.n_mels

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.conan_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def state_bytes(self) -> int:
        return int(self.lib.conan_engine_state_bytes(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.conan_engine_launch_count(self.h))

    @property
    def graph_replays(self) -> int:
        return int(self.lib.conan_engine_graph_replays(self.h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _host_ids(slots: Sequence[int]):
        arr = np.ascontiguousarray(np.asarray(slots, dtype=np.int32))
        return arr, arr.ctypes.data_as(C.c_void_p)

    def ids_tensor(self, slots: Sequence[int]) -> torch.Tensor:
        return torch.as_tensor(np.asarray(slots, dtype=np.int32), device=self.device)

    # ------------------------------------------------------------------ state
    def reset_slots(self, slots: Sequence[int], parts: int = PARTS_ALL):
        arr, p = self._host_ids(slots)
        _lib.check(self.lib.conan_slots_reset(self.h, len(arr), p, parts, self._stream()), "slots_reset")

    def open_sessions(self, slots: Sequence[int], ref_mel: torch.Tensor):
        """ref_mel [n, T_ref, n_mels] fp32 on the device (same T_ref for the whole call)."""
        assert ref_mel.dim() == 3 and ref_mel.shape[0] == len(slots) and ref_mel.shape[2] == self.n_mels
        ref_mel = ref_mel.to(self.device, torch.float32).contiguous()
        arr, p = self._host_ids(slots)
        _lib.check(self.lib.conan_session_open(self.h, len(arr), p, _ptr(ref_mel), ref_mel.shape[1], self._stream()), "session_open")

    # ------------------------------------------------------------------ steps (device tensors)
    def emformer_step(self, ids: torch.Tensor, chunk: torch.Tensor, want_enc=False, want_logits=False):
        n = ids.numel()
        assert chunk.shape == (n, self.rows_in, self.cfg.emformer_dim) and chunk.dtype == torch.float32 and chunk.is_contiguous()
        tokens = torch.empty(n, self.segment, dtype=torch.int32, device=self.device)
        enc = torch.empty(n, self.segment, self.cfg.emformer_dim, device=self.device) if want_enc else None
        logits = torch.empty(n, self.segment, self.cfg.emformer_output_dim, device=self.device) if want_logits else None
        _lib.check(self.lib.conan_emformer_step(self.h, n, _ptr(ids), _ptr(chunk), _ptr(enc), _ptr(logits), _ptr(tokens), self._stream()), "emformer_step")
        return tokens, enc, logits

    def emformer_forward(self, slots: Sequence[int], inp: torch.Tensor, want_enc=True, want_logits=False, want_tokens=False):
        """Full-utterance forward (EmformerDistillModel.forward / torchaudio Emformer.forward): inp [n, T + rc, dim], the
        utterance right-padded with the look-ahead frames -> (enc [n, T, dim], logits [n, T, out_dim], tokens [n, T]).
        Resets the Emformer state of `slots`."""
        n, frames, D = inp.shape
        rc = self.rows_in - self.segment
        assert n == len(slots) and D == self.cfg.emformer_dim and frames > rc
        inp = inp.to(self.device, torch.float32).contiguous()
        T = frames - rc
        enc = torch.empty(n, T, D, device=self.device) if want_enc else None
        logits = torch.empty(n, T, self.cfg.emformer_output_dim, device=self.device) if want_logits else None
        tokens = torch.empty(n, T, dtype=torch.int32, device=self.device) if want_tokens else None
        arr, p = self._host_ids(slots)
        _lib.check(self.lib.conan_emformer_forward(self.h, n, p, _ptr(inp), frames, _ptr(enc), _ptr(logits), _ptr(tokens), self._stream()),
                   "emformer_forward")
        return enc, logits, tokens

    def decoder_step(self, ids: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        n = ids.numel()
        assert tokens.shape == (n, self.segment) and tokens.dtype == torch.int32 and tokens.is_contiguous()
        mel = torch.empty(n, self.segment, self.n_mels, device=self.device)
        _lib.check(self.lib.conan_decoder_step(self.h, n, _ptr(ids), _ptr(tokens), _ptr(mel), self._stream()), "decoder_step")
        return mel

    def vocoder_step(self, ids: torch.Tensor, mel: torch.Tensor) -> torch.Tensor:
        n = ids.numel()
        assert mel.shape == (n, self.segment, self.n_mels) and mel.dtype == torch.float32 and mel.is_contiguous()
        wav = torch.empty(n, self.hop_out, device=self.device)
        _lib.check(self.lib.conan_vocoder_step(self.h, n, _ptr(ids), _ptr(mel), _ptr(wav), self._stream()), "vocoder_step")
        return wav

    def step(self, ids: torch.Tensor, chunk: torch.Tensor, wav: Optional[torch.Tensor] = None,
             mel: Optional[torch.Tensor] = None, tokens: Optional[torch.Tensor] = None, want_mel=True, want_tokens=True):
        n = ids.numel()
        assert chunk.shape == (n, self.rows_in, self.cfg.emformer_dim) and chunk.dtype == torch.float32 and chunk.is_contiguous()
        if wav is None:
            wav = torch.empty(n, self.hop_out, device=self.device)
        if mel is None and want_mel:
            mel = torch.empty(n, self.segment, self.n_mels, device=self.device)
        if tokens is None and want_tokens:
            tokens = torch.empty(n, self.segment, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.conan_step(self.h, n, _ptr(ids), _ptr(chunk), _ptr(wav), _ptr(mel), _ptr(tokens), self._stream()), "step")
        return wav, mel, tokens

    # ------------------------------------------------------------------ step with HOST buffers (the plugin call)
    def step_host(self, slots: np.ndarray, chunk: np.ndarray, wav_out: np.ndarray, mel_out: Optional[np.ndarray] = None,
                  tokens_out: Optional[np.ndarray] = None):
        n = len(slots)
        assert slots.dtype == np.int32 and chunk.dtype == np.float32 and chunk.shape == (n, self.rows_in, self.cfg.emformer_dim)
        assert wav_out.dtype == np.float32 and wav_out.shape == (n, self.hop_out)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _lib.check(self.lib.conan_step_host(self.h, n, vp(slots), vp(chunk), vp(wav_out), vp(mel_out), vp(tokens_out), self._stream()), "step_host")

    def step_host_submit(self, slots: np.ndarray, chunk: np.ndarray, wav_out: np.ndarray, mel_out: Optional[np.ndarray] = None,
                         tokens_out: Optional[np.ndarray] = None) -> int:
        """Pipelined step_host: returns a ticket at once; the result copies of this step overlap the next step's compute.
        At most two steps in flight; every array passed must stay alive and untouched until step_host_wait(ticket)."""
        n = len(slots)
        assert slots.dtype == np.int32 and chunk.dtype == np.float32 and chunk.shape == (n, self.rows_in, self.cfg.emformer_dim)
        assert wav_out.dtype == np.float32 and wav_out.shape == (n, self.hop_out)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        t = C.c_int(-1)
        _lib.check(self.lib.conan_step_host_submit(self.h, n, vp(slots), vp(chunk), vp(wav_out), vp(mel_out), vp(tokens_out), self._stream(),
                                                   C.byref(t)), "step_host_submit")
        return t.value

    def step_host_wait(self, ticket: int):
        _lib.check(self.lib.conan_step_host_wait(self.h, ticket), "step_host_wait")

    # ------------------------------------------------------------------ measurement
    def set_profiling(self, enabled: bool):
        _lib.check(self.lib.conan_engine_set_profiling(self.h, int(enabled)), "set_profiling")

    def profile_read(self, category: int):
        """category 0 FFMA, 1 tcgen05 ring (fp16), 2 tcgen05 window, 3 tcgen05 ring (split fp16), 4 fused residual block,
        5 fused Emformer FFN, 6 fused Conan block
        -> (ms, launches, algorithmic flops, algorithmic bytes)."""
        ms, n, fl, by = C.c_double(), C.c_uint64(), C.c_double(), C.c_double()
        _lib.check(self.lib.conan_engine_profile_read(self.h, category, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)), "profile_read")
        return ms.value, n.value, fl.value, by.value

    # ------------------------------------------------------------------ debug
    def debug_read(self, name: str, slot: int) -> torch.Tensor:
        numel = C.c_size_t()
        _lib.check(self.lib.conan_debug_read(self.h, name.encode(), slot, None, 0, C.byref(numel), self._stream()), "debug_read")
        out = torch.empty(numel.value, device=self.device)
        _lib.check(self.lib.conan_debug_read(self.h, name.encode(), slot, _ptr(out), out.numel(), C.byref(numel), self._stream()), "debug_read")
        return out
