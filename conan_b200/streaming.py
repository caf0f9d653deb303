"""Drop-in host surface of the reference's streaming driver.

`StreamingVoiceConversion` keeps the constructor / `infer_once` / `test_multiple_sentences`
signatures, hparams keys and checkpoint layouts of `inference/Conan.py:20-176`; underneath, the
three PyTorch modules are replaced by the native engine (one resident slot per stream, incremental
chunk steps instead of the reference's full-history recompute).  The module-level call surface the
reference loop uses is kept as thin views over the same engine:
    emformer.emformer.infer(chunk, lengths, state) / emformer.proj / emformer.mode   (:115-120)
    model(content=..., ref=..., infer=True, ...)["mel_out"]                              (:131-142)
    vocoder.spec2wav(mel[T, 80]) -> wav[T*hop]                                            (:149)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import audio, ckpt, synth
from .frontend import GpuLogMel
from .engine import Engine, PARTS_CONAN, PARTS_EMFORMER, PARTS_VOCODER, make_config
from .hparams import hparams, set_hparams
from .scheduler import ChunkScheduler

REGISTERED_VOCODERS: Dict[str, type] = {}


def register_vocoder(name):
    """tasks/tts/vocoder_infer/base_vocoder.py:9-15"""
    def _f(cls):
        REGISTERED_VOCODERS[name] = cls
        return cls
    return _f


def get_vocoder_cls(vocoder_name):
    return REGISTERED_VOCODERS.get(vocoder_name)


def load_voc_config(vocoder_ckpt: str) -> dict:
    """`{vocoder_ckpt}/config.yaml` (tasks/tts/vocoder_infer/hifigan.py:14-16)."""
    return set_hparams(f"{vocoder_ckpt}/config.yaml", global_hparams=False, print_hparams=False)


class _LazyDefaults:
    """Initial tensors of a spec, generated only if a checkpoint really leaves a hole."""

    def __init__(self, spec, seed):
        self.spec, self.seed, self._sd = spec, seed, None

    def __getitem__(self, key):
        if self._sd is None:
            self._sd = synth.make_state_dict(self.spec, self.seed)
        return self._sd[key]


def build_engine(hp: Dict, *, max_streams: int = 8, max_ref_frames: int = 1024, voc_precision: str = "fp16",
                 voc_tensor_cores: bool = True, device: int = 0) -> Engine:
    """Reads the three checkpoints named by the hparams (work_dir / emformer_ckpt / vocoder_ckpt, same
    selection rules as utils/commons/ckpt_utils.py:26-66) and builds the native engine."""
    voc_hp = load_voc_config(hp["vocoder_ckpt"])
    # strict=False (inference/Conan.py:37,51): shape-mismatched / missing keys keep the initial value; the reference's
    # nn.Module supplies it, here it is the seeded default init of conan_b200.synth (hparams `seed`)
    seed = int(hp.get("seed", 1234))
    spec_c, spec_e = synth.conan_spec(hp), synth.emformer_spec(hp)
    sd_c = ckpt.filter_to_spec(ckpt.load_state_dict(hp["work_dir"], "model"), spec_c, strict=False,
                               defaults=_LazyDefaults(spec_c, seed))
    sd_e = ckpt.filter_to_spec(ckpt.load_state_dict(hp["emformer_ckpt"], "model"), spec_e, strict=False,
                               defaults=_LazyDefaults(spec_e, seed))
    sd_v = ckpt.filter_to_spec(ckpt.load_state_dict(hp["vocoder_ckpt"], "model_gen"), synth.hifigan_spec(voc_hp), strict=True)
    cfg = make_config(hp, voc_hp, max_slots=max_streams, max_ref_frames=max_ref_frames, device=device,
                      voc_precision=voc_precision, voc_tensor_cores=voc_tensor_cores)
    return Engine(sd_c, sd_e, sd_v, cfg)


# ------------------------------------------------------------------------------------------------
# module-level views
# ------------------------------------------------------------------------------------------------
class _EmformerCore:
    """`.infer(chunk [B,6,80], lengths [B], state)` -> (out [B,4,80], lengths - rc, state)  (TA:745-803).
    `state` is an opaque handle on engine slots (None starts new streams, like the reference)."""

    def __init__(self, eng: Engine, slot_pool: List[int]):
        self.eng, self.pool = eng, slot_pool

    def infer(self, input: torch.Tensor, lengths: torch.Tensor, states=None):
        B = input.shape[0]
        if input.shape[1] != self.eng.rows_in:
            raise ValueError(f"expected size of {self.eng.rows_in} for dimension 1 of input, but got {input.shape[1]}.")
        if states is None:
            if B > len(self.pool):
                raise RuntimeError("not enough free slots for a new Emformer state")
            states = {"slots": self.pool[:B]}
            self.eng.reset_slots(states["slots"], PARTS_EMFORMER)
        ids = self.eng.ids_tensor(states["slots"])
        chunk = input.to(self.eng.device, torch.float32).contiguous()
        _, enc, _ = self.eng.emformer_step(ids, chunk, want_enc=True)
        return enc, torch.clamp(lengths - (self.eng.rows_in - self.eng.segment), min=0), states

    def forward(self, input: torch.Tensor, lengths: torch.Tensor):
        """torchaudio Emformer.forward (TA:709-743): input [B, T + rc, D] right-padded with the look-ahead frames ->
        (output [B, T, D], lengths).  Every batch element is taken at full length (the reference's callers pass full lengths)."""
        B = input.shape[0]
        if B > len(self.pool):
            raise RuntimeError("not enough free slots for a full-utterance forward")
        enc, _, _ = self.eng.emformer_forward(self.pool[:B], input)
        return enc, lengths

    __call__ = forward


class EmformerView:
    """Stands in for modules/Emformer/emformer.py::EmformerDistillModel on the inference path."""

    def __init__(self, eng: Engine, slot_pool: List[int]):
        self.eng = eng
        self.emformer = _EmformerCore(eng, slot_pool)
        self.mode = None
        self.segment_length = eng.segment
        self.right_context_len = eng.rows_in - eng.segment

    def forward(self, mel_input: torch.Tensor, lengths: torch.Tensor):
        """EmformerDistillModel.forward (modules/Emformer/emformer.py:31-47): (proj(emformer(mel_input, lengths)), lengths)."""
        output, lengths = self.emformer(mel_input, lengths)
        return self.proj(output), lengths

    __call__ = forward

    def proj(self, x: torch.Tensor) -> torch.Tensor:
        """Linear(80 -> emformer_output_dim) through the FFMA conv-GEMM operator."""
        from . import ops
        B, T, D = x.shape
        w, b = self.eng.aux["emf.proj.w"], self.eng.aux["emf.proj.b"]
        y = torch.empty(B, T, w.shape[0], device=self.eng.device)
        ops.conv_gemm(x.to(self.eng.device, torch.float32).contiguous(), w, b, k=1, dil=1, L=T, row0=0, y=y)
        return y


class ConanView:
    """`model(content=codes [B,T], ref=[B,T_ref,80], infer=True, ...)["mel_out"]` computed from scratch
    over all T tokens (session setup + T/4 causal chunk steps) -- the reference's full-history call."""

    def __init__(self, eng: Engine, slot_pool: List[int]):
        self.eng, self.pool = eng, slot_pool

    def __call__(self, content, spk_embed=None, target=None, ref=None, f0=None, uv=None, infer=True, global_steps=0, **kw):
        if ref is None or spk_embed is not None or not infer:
            raise ValueError("only the inference call (ref given, spk_embed None, infer=True) is on the hot path")
        B, T = content.shape
        if B > len(self.pool):
            raise RuntimeError("not enough free slots")
        slots = self.pool[:B]
        seg = self.eng.segment
        self.eng.reset_slots(slots, PARTS_CONAN)
        self.eng.open_sessions(slots, ref.to(self.eng.device, torch.float32))
        ids = self.eng.ids_tensor(slots)
        pad = (-T) % seg
        tok = torch.nn.functional.pad(content.to(torch.int32), (0, pad)).to(self.eng.device)
        mels = [self.eng.decoder_step(ids, tok[:, i:i + seg].contiguous()) for i in range(0, T + pad, seg)]
        return {"mel_out": torch.cat(mels, 1)[:, :T], "content": content}


@register_vocoder("HifiGAN")
class HifiGAN:
    """tasks/tts/vocoder_infer/hifigan.py:11-31: `spec2wav(mel np[T,80]) -> wav np[T*hop]`."""
    _engine_factory = None      # set by StreamingVoiceConversion so the vocoder shares its engine

    def __init__(self, eng: Optional[Engine] = None, slot: int = 0):
        if eng is None:
            if HifiGAN._engine_factory is None:
                raise RuntimeError("HifiGAN() needs an engine: construct it through StreamingVoiceConversion "
                                   "or pass eng=build_engine(hparams)")
            eng, slot = HifiGAN._engine_factory()
        self.eng, self.slot = eng, slot
        self.device = eng.device

    def spec2wav(self, mel, **kwargs):
        mel = np.asarray(mel, dtype=np.float32)
        T, seg = mel.shape[0], self.eng.segment
        if T == 0:
            return np.zeros(0, np.float32)
        pad = (-T) % seg
        m = torch.from_numpy(np.pad(mel, ((0, pad), (0, 0)), mode="edge"))[None].to(self.device)
        self.eng.reset_slots([self.slot], PARTS_VOCODER)
        ids = self.eng.ids_tensor([self.slot])
        wav = torch.cat([self.eng.vocoder_step(ids, m[:, i:i + seg].contiguous()) for i in range(0, T + pad, seg)], 1)
        hop = self.eng.hop_out // seg
        return wav[0, :T * hop].cpu().numpy()


# ------------------------------------------------------------------------------------------------
class StreamingVoiceConversion:
    """Streaming style-transfer inference (drop-in for inference/Conan.py:20-176)."""
    tokens_per_chunk: int = 4

    def __init__(self, hp: Dict, max_streams: int = 8, **engine_kw):
        if not torch.cuda.is_available():
            raise RuntimeError("conan_b200 has no CPU path: a CUDA (sm_100a) device is required")
        self.hparams = hp
        self.device = "cuda"
        if get_vocoder_cls(hp["vocoder"]) is None:
            raise ValueError(f"Vocoder '{hp['vocoder']}' is not registered. Check vocoder name and registration.")
        self.engine = build_engine(hp, max_streams=max_streams + 3, **engine_kw)
        # the last three slots back the module-level views; the others belong to the scheduler
        view_slots = list(range(max_streams, max_streams + 3))
        self.scheduler = ChunkScheduler(self.engine, max_streams)
        self.model = ConanView(self.engine, view_slots[:1])
        self.emformer = EmformerView(self.engine, view_slots[1:2])
        # the registry contract is a zero-argument constructor (base_vocoder.py:17 / inference/Conan.py:40-45); the engine is handed
        # over through a factory that is in effect for the duration of this one constructor call only, so a second
        # StreamingVoiceConversion cannot rebind the vocoder of the first
        voc_cls = get_vocoder_cls(hp["vocoder"])
        prev_factory = getattr(voc_cls, "_engine_factory", None)
        voc_cls._engine_factory = lambda: (self.engine, view_slots[2])
        try:
            self.vocoder = voc_cls()
        finally:
            voc_cls._engine_factory = prev_factory
        self.frontend = GpuLogMel(hp)
        self._vocoder_warm_zero()

    def _vocoder_warm_zero(self):
        _ = self.vocoder.spec2wav(np.zeros((4, 80), dtype=np.float32))

    def _wav_to_mel(self, path: str) -> np.ndarray:
        """inference/Conan.py:58-70: wav file -> clipped log-mel [T, 80], computed on the device (`conan_logmel`)."""
        wav = audio.load_wav(path, self.hparams["audio_sample_rate"])
        if self.hparams.get("loud_norm", False):
            wav = audio.loudness_normalize(wav, self.hparams["audio_sample_rate"])
        return self.frontend.offline(wav)[0].cpu().numpy()

    def infer_mels(self, ref_mel: np.ndarray, src_mel: np.ndarray):
        """The loop of infer_once on precomputed mels: (wav float32 [T*hop], mel float32 [T, 80])."""
        sid = self.scheduler.open(ref_mel)
        wavs, mels = [], []
        try:
            self.scheduler.push(sid, src_mel)
            self.scheduler.end(sid)
            while not self.scheduler.finished(sid):
                w, m, _ = self.scheduler.step()[sid]
                wavs.append(w), mels.append(m)
        finally:
            self.scheduler.close(sid)              # the slot goes back to the pool even if a step raised
        if not wavs:
            return np.zeros(0, np.float32), np.zeros((0, 80), np.float32)
        return np.concatenate(wavs), np.concatenate(mels)

    def infer_once(self, inp: Dict):
        ref_mel = self._wav_to_mel(inp["ref_wav"])
        src_mel = self._wav_to_mel(inp["src_wav"])
        return self.infer_mels(ref_mel, src_mel)

    def test_multiple_sentences(self, test_cases: List[Dict]):
        os.makedirs("infer_out_demo", exist_ok=True)
        for inp in test_cases:
            wav, _ = self.infer_once(inp)
            ref_name = os.path.splitext(os.path.basename(inp["ref_wav"]))[0]
            src_name = os.path.splitext(os.path.basename(inp["src_wav"]))[0]
            save_path = f"infer_out_demo/{ref_name}_{src_name}.wav"
            audio.save_wav(wav, save_path, self.hparams["audio_sample_rate"])
            print(f"Saved output: {save_path}")


if __name__ == "__main__":
    set_hparams()
    demo = [{"ref_wav": "path/to/reference_audio.wav", "src_wav": "path/to/source_audio.wav"}]
    engine = StreamingVoiceConversion(hparams)
    engine.test_multiple_sentences(demo)
