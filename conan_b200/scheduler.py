"""Chunk scheduler: resident per-stream state + packing of variable-readiness streams into one
launch sequence (north-star component 4).

The reference has no scheduler -- its loop (`inference/Conan.py:95-156`) serves one utterance and
re-runs the models on all history.  The contract kept here is that loop's chunk assembly:
  emit = min(seg, T - pos); look = min(rc, T - pos - emit); the chunk is frames [pos, pos+emit+look)
  padded to seg+rc rows by repeating its last frame (:97-110); a stream advances by `emit` frames.
A stream whose source is still arriving is *ready* when seg+rc frames past `pos` are buffered; once
`end()` was called the remaining frames are flushed with the same padding rule.  `step()` gathers the
ready streams' chunks into one [n, seg+rc, 80] host buffer, makes ONE engine call (slot ids + chunks
in, wav/mel/tokens out) and hands each stream its `emit` frames of output.

The engine is duck-typed (`reset_slots`, `open_sessions`, `step_host`, `segment`, `rows_in`,
`hop_out`, `n_mels`) so the host logic is unit-tested on CPU with a recording fake.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class _Stream:
    slot: int
    frames: List[np.ndarray] = field(default_factory=list)   # pending source mel, [n_i, 80] pieces
    buffered: Optional[np.ndarray] = None                     # concatenated view, lazily rebuilt
    pos: int = 0                                              # frames already consumed (emitted)
    ended: bool = False

    def mel(self) -> np.ndarray:
        if self.frames:
            parts = ([self.buffered] if self.buffered is not None else []) + self.frames
            self.buffered = np.concatenate(parts, axis=0)
            self.frames = []
        return self.buffered if self.buffered is not None else np.zeros((0, 80), np.float32)


class ChunkScheduler:
    def __init__(self, engine, max_streams: int):
        self.eng = engine
        self.seg, self.rows = engine.segment, engine.rows_in
        self.rc = self.rows - self.seg
        self.free = list(range(max_streams - 1, -1, -1))
        self.streams: Dict[int, _Stream] = {}
        self._next_id = 0
        self._chunk_buf = np.zeros((max_streams, self.rows, engine.n_mels), np.float32)
        self._wav_buf = np.zeros((max_streams, engine.hop_out), np.float32)
        self._mel_buf = np.zeros((max_streams, self.seg, engine.n_mels), np.float32)
        self._tok_buf = np.zeros((max_streams, self.seg), np.int32)

    # ------------------------------------------------------------------ session admission
    def open(self, ref_mel) -> int:
        return self.open_many([ref_mel])[0]

    def open_many(self, ref_mels) -> List[int]:
        """ref_mels: list of [T_ref, 80] arrays.  Streams with equal T_ref share one session-setup call."""
        import torch
        if len(ref_mels) > len(self.free):
            raise RuntimeError(f"no free stream slot ({len(ref_mels)} requested, {len(self.free)} free)")
        ids, by_len = [], {}
        for m in ref_mels:
            m = np.asarray(m, dtype=np.float32)
            slot = self.free.pop()
            sid = self._next_id
            self._next_id += 1
            self.streams[sid] = _Stream(slot=slot)
            ids.append(sid)
            by_len.setdefault(m.shape[0], []).append((slot, m))
        self.eng.reset_slots([self.streams[s].slot for s in ids])
        for _, group in by_len.items():
            slots = [g[0] for g in group]
            ref = torch.from_numpy(np.stack([g[1] for g in group]))
            self.eng.open_sessions(slots, ref.to(self.eng.device) if hasattr(self.eng, "device") else ref)
        return ids

    def close(self, sid: int):
        st = self.streams.pop(sid)
        self.free.append(st.slot)

    # ------------------------------------------------------------------ input side
    def push(self, sid: int, mel_frames):
        st = self.streams[sid]
        if st.ended:
            raise RuntimeError("push() after end()")
        f = np.asarray(mel_frames, dtype=np.float32)
        if f.ndim != 2 or f.shape[1] != self.eng.n_mels:
            raise ValueError("mel_frames must be [n, n_mels]")
        if f.shape[0]:
            st.frames.append(f)

    def end(self, sid: int):
        self.streams[sid].ended = True

    def _ready(self, st: _Stream) -> bool:
        avail = st.mel().shape[0] - st.pos
        return avail >= self.rows or (st.ended and avail > 0)

    def ready(self) -> List[int]:
        return [sid for sid, st in self.streams.items() if self._ready(st)]

    def finished(self, sid: int) -> bool:
        st = self.streams[sid]
        return st.ended and st.mel().shape[0] - st.pos <= 0

    # ------------------------------------------------------------------ one packed step
    def assemble(self, st: _Stream) -> Tuple[np.ndarray, int]:
        mel = st.mel()
        T = mel.shape[0] if st.ended else max(mel.shape[0], st.pos + self.rows)
        emit = min(self.seg, T - st.pos)
        look = min(self.rc, T - (st.pos + emit))
        chunk = mel[st.pos:st.pos + emit + look]
        need = self.rows - chunk.shape[0]
        if need > 0:
            chunk = np.concatenate([chunk, np.repeat(chunk[-1:], need, axis=0)], axis=0)
        return chunk, emit

    def step(self, max_batch: Optional[int] = None) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        """Runs one chunk step for every ready stream (one launch sequence).  Returns
        {stream id: (wav [emit*hop], mel [emit, 80], tokens [emit])}."""
        sids = self.ready()
        if max_batch is not None:
            sids = sids[:max_batch]
        n = len(sids)
        if n == 0:
            return {}
        slots = np.empty(n, np.int32)
        emits = []
        for i, sid in enumerate(sids):
            st = self.streams[sid]
            self._chunk_buf[i], emit = self.assemble(st)
            slots[i] = st.slot
            emits.append(emit)
        self.eng.step_host(slots, self._chunk_buf[:n], self._wav_buf[:n], self._mel_buf[:n], self._tok_buf[:n])
        hop = self.eng.hop_out // self.seg
        out = {}
        for i, sid in enumerate(sids):
            e = emits[i]
            self.streams[sid].pos += e
            out[sid] = (self._wav_buf[i, :e * hop].copy(), self._mel_buf[i, :e].copy(), self._tok_buf[i, :e].copy())
        return out


def shard_streams(n_streams: int, world_size: int, rank: int) -> range:
    """Contiguous block partition of stream ids over ranks: sessions are independent,
    every rank owns a disjoint block and its own weight replica; there is no data-path collective."""
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
