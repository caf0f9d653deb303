"""Chunk scheduler: resident per-stream state + packing of variable-readiness streams into one
launch sequence (north-star component 4).

The reference has no scheduler -- its loop (`inference/Conan.py:95-156`) serves one utterance and
re-runs the models on all history.  The contract kept here is that loop's chunk assembly:
  emit = min(seg, T - pos); look = min(rc, T - pos - emit); the chunk is frames [pos, pos+emit+look)
  padded to seg+rc rows by repeating its last frame (:97-110); a stream advances by `emit` frames.
A stream whose source is still arriving is *ready* when seg+rc frames past `pos` are buffered; once
`end()` was called the remaining frames are flushed with the same padding rule.  One step gathers the
ready streams' chunks into one [n, seg+rc, 80] staging buffer, makes ONE engine call (slot ids + chunks
in, wav/mel/tokens out) and hands each stream its `emit` frames of output.

Host data structures (all indexed by slot, struct-of-arrays, so a step is a handful of vectorised numpy
operations whatever the number of streams -- there is no per-stream Python loop on the step path):
  * `ring  [max_streams, cap, n_mels]` fp32: the not-yet-consumed source frames of every stream, frame f of a
    stream at row f mod cap.  Bounded: a stream holds at most `cap` frames here whatever its length;
  * `recv / pos / ended / total`: frames written to the ring, frames consumed, end-of-stream flag, final length;
  * a per-stream backlog (only for callers that push more than the ring holds, e.g. a whole utterance at
    once): the caller's arrays wait there and refill the ring as frames are consumed.
Chunk assembly is one `np.take` with the index  min(pos + i, last) mod cap  (the clamp IS the reference's
repeat-last-frame padding) straight into the staging buffer the engine copies from.

Two stepping styles: `step_packed()` (synchronous: results of this step) and `submit()` / `collect()`
(two steps in flight: the result copy of step i runs under the compute of step i+1).

The engine is duck-typed (`reset_slots`, `open_sessions`, `step_host`, optionally `step_host_submit` /
`step_host_wait`, `segment`, `rows_in`, `hop_out`, `n_mels`) so the host logic is unit-tested on CPU with a
recording fake.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Deque, Dict, List, Optional, Tuple

import numpy as np


@dataclass
class PackedStep:
    """Result of one packed step: row i belongs to stream `sids[i]` (slot `slots[i]`); only the first
    `emits[i]` frames (emits[i] * hop samples) of a row are output, the rest is look-ahead padding."""
    sids: np.ndarray       # [n] int64
    slots: np.ndarray      # [n] int32
    emits: np.ndarray      # [n] int32
    wav: np.ndarray        # [n, seg * hop] fp32
    mel: np.ndarray        # [n, seg, n_mels] fp32
    tokens: np.ndarray     # [n, seg] int32
    hop: int

    def __len__(self):
        return len(self.sids)

    def per_stream(self) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        return {int(s): (self.wav[i, :e * self.hop].copy(), self.mel[i, :e].copy(), self.tokens[i, :e].copy())
                for i, (s, e) in enumerate(zip(self.sids, self.emits))}


class _StreamView:
    """Read-only view of one stream's scheduler state (kept for callers that look at `.slot` / `.pos`)."""

    def __init__(self, sch: "ChunkScheduler", slot: int):
        self._s, self.slot = sch, slot

    @property
    def pos(self) -> int:
        return int(self._s.pos[self.slot])

    @property
    def ended(self) -> bool:
        return bool(self._s.ended[self.slot])


def _pinned_empty(shape, dtype):
    """Page-locked staging memory when a CUDA runtime is present (the engine copies from / to it asynchronously)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(shape, dtype={np.float32: torch.float32, np.int32: torch.int32}[dtype]).pin_memory()
            return t.numpy(), t
    except Exception:
        pass
    return np.zeros(shape, dtype), None


class ChunkScheduler:
    def __init__(self, engine, max_streams: int, capacity_frames: int = 64):
        self.eng = engine
        self.seg, self.rows = engine.segment, engine.rows_in
        self.rc = self.rows - self.seg
        self.n_mels = engine.n_mels
        self.hop = engine.hop_out // self.seg
        self.S = max_streams
        self.cap = max(int(capacity_frames), 2 * self.rows)
        self.free = list(range(max_streams - 1, -1, -1))
        self.streams: Dict[int, _StreamView] = {}
        self._next_id = 0
        S = max_streams
        self.ring = np.zeros((S, self.cap, self.n_mels), np.float32)
        self._ring2d = self.ring.reshape(S * self.cap, self.n_mels)
        self.recv = np.zeros(S, np.int64)          # frames written to the ring so far
        self.pos = np.zeros(S, np.int64)           # frames consumed (emitted)
        self.total = np.zeros(S, np.int64)         # frames pushed in all (ring + backlog)
        self.ended = np.zeros(S, bool)
        self.active = np.zeros(S, bool)
        self.sid_of = np.full(S, -1, np.int64)
        self._backlog: Dict[int, Deque[np.ndarray]] = {}     # slot -> frames waiting for ring space
        self._row_off = np.arange(self.rows, dtype=np.int64)[None, :]
        # two sets of staging buffers (pipelined stepping alternates between them)
        self._keep = []
        self._stage = []
        for _ in range(2):
            bufs = {}
            for name, shape, dt in (("chunk", (S, self.rows, self.n_mels), np.float32), ("wav", (S, engine.hop_out), np.float32),
                                    ("mel", (S, self.seg, self.n_mels), np.float32), ("tok", (S, self.seg), np.int32),
                                    ("slots", (S,), np.int32)):
                arr, keep = _pinned_empty(shape, dt)
                bufs[name] = arr
                self._keep.append(keep)
            self._stage.append(bufs)
        self._flip = 0
        self._inflight: Deque[Tuple[int, int, np.ndarray, np.ndarray, np.ndarray, int]] = deque()

    def warm(self, max_batch: int = 64, full: bool = False):
        """Start-up warm-up of a serving loop: runs throw-away steps for every ready-count bucket (multiples of 8) up to
        `max_batch` (and the full size with full=True), through both pipelined buffer sets, so that the engine has captured its
        CUDA graphs before the first real chunk arrives (a capture costs a few ms -- fine at start-up, a p99 outlier in service).
        Must run before any session is opened; the slots it touched are reset."""
        if self.streams:
            raise RuntimeError("warm() must run before sessions are opened")
        sizes = sorted(set(list(range(8, min(max_batch, self.S) + 1, 8)) + ([self.S] if full else [])))
        pipelined = hasattr(self.eng, "step_host_submit")
        for n in sizes:
            slots = np.arange(n, dtype=np.int32)
            reps = 3 if n <= 256 else 5                               # eager run(s), capture, replay -- per buffer set
            for rep in range(2 * reps if pipelined else reps):
                b = self._stage[rep & 1]
                b["chunk"][:n] = -3.0
                b["slots"][:n] = slots
                if pipelined:
                    self.eng.step_host_wait(self.eng.step_host_submit(b["slots"][:n], b["chunk"][:n], b["wav"][:n], b["mel"][:n], b["tok"][:n]))
                else:
                    self.eng.step_host(b["slots"][:n], b["chunk"][:n], b["wav"][:n], b["mel"][:n], b["tok"][:n])
        if sizes:
            self.eng.reset_slots(list(range(max(sizes))))

    # ------------------------------------------------------------------ session admission
    def open(self, ref_mel) -> int:
        return self.open_many([ref_mel])[0]

    def open_many(self, ref_mels) -> List[int]:
        """ref_mels: list of [T_ref, 80] arrays.  Streams with equal T_ref share one session-setup call.
        Nothing is allocated if a reference is malformed; if the engine rejects a session every slot taken by
        this call is returned to the pool."""
        import torch
        refs = [np.asarray(m, dtype=np.float32) for m in ref_mels]
        max_ref = getattr(getattr(self.eng, "cfg", None), "max_ref_frames", None)
        for m in refs:
            if m.ndim != 2 or m.shape[1] != self.n_mels or m.shape[0] < 1:
                raise ValueError(f"reference mel must be [T_ref >= 1, {self.n_mels}], got {m.shape}")
            if max_ref is not None and m.shape[0] > max_ref:
                raise ValueError(f"reference mel has {m.shape[0]} frames; this engine was built for at most {max_ref} "
                                 "(max_ref_frames)")
        if len(refs) > len(self.free):
            raise RuntimeError(f"no free stream slot ({len(refs)} requested, {len(self.free)} free)")
        ids, slots, by_len = [], [], {}
        for m in refs:
            slot = self.free.pop()
            sid = self._next_id
            self._next_id += 1
            self._activate(slot, sid)
            ids.append(sid)
            slots.append(slot)
            by_len.setdefault(m.shape[0], []).append((slot, m))
        try:
            self.eng.reset_slots(slots)
            for _, group in by_len.items():
                ref = torch.from_numpy(np.stack([g[1] for g in group]))
                self.eng.open_sessions([g[0] for g in group], ref.to(self.eng.device) if hasattr(self.eng, "device") else ref)
        except Exception:
            for sid in ids:
                self.close(sid)
            raise
        return ids

    def _activate(self, slot: int, sid: int):
        self.recv[slot] = self.pos[slot] = self.total[slot] = 0
        self.ended[slot] = False
        self.active[slot] = True
        self.sid_of[slot] = sid
        self.streams[sid] = _StreamView(self, slot)

    def close(self, sid: int):
        st = self.streams.pop(sid)
        self.active[st.slot] = False
        self.sid_of[st.slot] = -1
        self._backlog.pop(st.slot, None)
        self.free.append(st.slot)

    # ------------------------------------------------------------------ input side
    def _ring_write(self, slot: int, f: np.ndarray) -> int:
        """Copies as many leading frames of f as the ring has room for; returns how many."""
        room = self.cap - int(self.recv[slot] - self.pos[slot])
        n = min(room, f.shape[0])
        if n > 0:
            w = int(self.recv[slot] % self.cap)
            first = min(n, self.cap - w)
            self.ring[slot, w:w + first] = f[:first]
            if n > first:
                self.ring[slot, :n - first] = f[first:n]
            self.recv[slot] += n
        return n

    def push(self, sid: int, mel_frames):
        st = self.streams[sid]
        slot = st.slot
        if self.ended[slot]:
            raise RuntimeError("push() after end()")
        f = np.asarray(mel_frames, dtype=np.float32)
        if f.ndim != 2 or f.shape[1] != self.n_mels:
            raise ValueError("mel_frames must be [n, n_mels]")
        if not f.shape[0]:
            return
        self.total[slot] += f.shape[0]
        bl = self._backlog.get(slot)
        if bl:                                   # keep arrival order: frames queue behind the backlog
            bl.append(f)
            return
        n = self._ring_write(slot, f)
        if n < f.shape[0]:
            self._backlog.setdefault(slot, deque()).append(f[n:])

    def push_many(self, slots: np.ndarray, frames: np.ndarray):
        """Vectorised push for streams fed in lock-step (a serving loop that receives the same number of new frames for a
        batch of streams): slots [n] (distinct), frames [n, f, n_mels].  Every stream must have ring room for f frames."""
        slots = np.asarray(slots, dtype=np.int64)
        f = frames.shape[1]
        if f == 0 or len(slots) == 0:
            return
        if self._backlog and any(int(s) in self._backlog for s in slots):
            raise BufferError("push_many: a stream still has a backlog from push(); drain it first")
        if (self.cap - (self.recv[slots] - self.pos[slots]) < f).any():
            raise BufferError("push_many: ring full for at least one stream (consumer is behind: back-pressure)")
        idx = (self.recv[slots][:, None] + np.arange(f, dtype=np.int64)[None, :]) % self.cap + slots[:, None] * self.cap
        self._ring2d[idx.reshape(-1)] = frames.reshape(-1, self.n_mels)
        self.recv[slots] += f
        self.total[slots] += f

    def end(self, sid: int):
        self.ended[self.streams[sid].slot] = True

    def buffered_frames(self, sid: int) -> int:
        """Frames held for this stream in the scheduler's own ring (bounded by the ring capacity)."""
        slot = self.streams[sid].slot
        return int(self.recv[slot] - self.pos[slot])

    def pending_frames(self, sid: int) -> int:
        slot = self.streams[sid].slot
        return int(self.total[slot] - self.pos[slot])

    # ------------------------------------------------------------------ readiness
    def _ready_mask(self) -> np.ndarray:
        avail = self.recv - self.pos
        flush = self.ended & (self.recv == self.total) & (avail > 0)
        return self.active & ((avail >= self.rows) | flush)

    def ready_slots(self) -> np.ndarray:
        return np.flatnonzero(self._ready_mask()).astype(np.int32)

    def ready(self) -> List[int]:
        return [int(s) for s in self.sid_of[self.ready_slots()]]

    def finished(self, sid: int) -> bool:
        slot = self.streams[sid].slot
        return bool(self.ended[slot]) and int(self.total[slot] - self.pos[slot]) <= 0

    # ------------------------------------------------------------------ one packed step
    def _assemble(self, slots: np.ndarray, out: np.ndarray) -> np.ndarray:
        """chunk rows of `slots` -> out[:n]; returns emit[n].  inference/Conan.py:97-110 as index arithmetic."""
        s64 = slots.astype(np.int64)
        pos, ended = self.pos[s64], self.ended[s64]
        remaining = np.where(ended, self.total[s64] - pos, np.int64(1 << 40))
        emit = np.minimum(self.seg, remaining)
        last = pos + np.minimum(self.rows, remaining) - 1                 # last real frame of the chunk (then repeated)
        src = np.minimum(pos[:, None] + self._row_off, last[:, None])
        flat = (src % self.cap + s64[:, None] * self.cap).reshape(-1)
        np.take(self._ring2d, flat, axis=0, out=out[:len(slots)].reshape(-1, self.n_mels), mode="clip")
        return emit.astype(np.int32)

    def _advance(self, slots: np.ndarray, emits: np.ndarray):
        self.pos[slots.astype(np.int64)] += emits
        for slot in [s for s in self._backlog if self._backlog[s]]:      # only streams that over-filled the ring
            bl = self._backlog[slot]
            while bl:
                n = self._ring_write(slot, bl[0])
                if n < bl[0].shape[0]:
                    bl[0] = bl[0][n:]
                    break
                bl.popleft()
            if not bl:
                del self._backlog[slot]

    def _pick(self, max_batch: Optional[int]) -> np.ndarray:
        # a stream may be part of both in-flight steps: the engine runs them in submission order on one CUDA stream, and
        # `pos` advanced when the first one was staged
        slots = self.ready_slots()
        if max_batch is not None:
            slots = slots[:max_batch]
        return slots

    def submit(self, max_batch: Optional[int] = None) -> Optional[int]:
        """Assembles the ready streams' chunks and enqueues one engine step without waiting for it (needs an engine with
        step_host_submit / step_host_wait).  Returns a ticket for collect(), or None when no stream is ready.
        At most two steps may be in flight."""
        if len(self._inflight) >= 2:
            raise RuntimeError("two steps already in flight: collect() first")
        slots = self._pick(max_batch)
        n = len(slots)
        if n == 0:
            return None
        b = self._stage[self._flip]
        b["slots"][:n] = slots
        emits = self._assemble(slots, b["chunk"])
        sids = self.sid_of[slots.astype(np.int64)].copy()
        ticket = self.eng.step_host_submit(b["slots"][:n], b["chunk"][:n], b["wav"][:n], b["mel"][:n], b["tok"][:n])
        self._advance(slots, emits)              # the chunk is staged: the ring rows are free for new frames
        self._inflight.append((ticket, self._flip, slots.copy(), emits, sids, n))
        self._flip ^= 1
        return ticket

    def collect(self, ticket: Optional[int] = None) -> PackedStep:
        """Waits for the oldest in-flight step and returns its outputs (views of the staging buffers: valid until the
        second-next submit())."""
        if not self._inflight:
            raise RuntimeError("no step in flight")
        t, flip, slots, emits, sids, n = self._inflight.popleft()
        if ticket is not None and ticket != t:
            raise RuntimeError("steps must be collected in submission order")
        self.eng.step_host_wait(t)
        b = self._stage[flip]
        return PackedStep(sids, slots, emits, b["wav"][:n], b["mel"][:n], b["tok"][:n], self.hop)

    def step_packed(self, max_batch: Optional[int] = None) -> Optional[PackedStep]:
        """One synchronous chunk step for every ready stream (one launch sequence)."""
        if self._inflight:
            raise RuntimeError("step_packed() while pipelined steps are in flight")
        slots = self._pick(max_batch)
        n = len(slots)
        if n == 0:
            return None
        b = self._stage[self._flip]
        b["slots"][:n] = slots
        emits = self._assemble(slots, b["chunk"])
        sids = self.sid_of[slots.astype(np.int64)].copy()
        self.eng.step_host(b["slots"][:n], b["chunk"][:n], b["wav"][:n], b["mel"][:n], b["tok"][:n])
        self._advance(slots, emits)
        return PackedStep(sids, slots.copy(), emits, b["wav"][:n], b["mel"][:n], b["tok"][:n], self.hop)

    def step(self, max_batch: Optional[int] = None) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        """Runs one chunk step for every ready stream.  Returns
        {stream id: (wav [emit*hop], mel [emit, 80], tokens [emit])} (copies; small-scale convenience over step_packed)."""
        r = self.step_packed(max_batch)
        return {} if r is None else r.per_stream()


def shard_streams(n_streams: int, world_size: int, rank: int) -> range:
    """Contiguous block partition of stream ids over ranks: sessions are independent,
    every rank owns a disjoint block and its own weight replica; there is no data-path collective."""
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
