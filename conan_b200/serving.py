"""Serving shell around the chunk scheduler, and the JSON batch runner.

Two layers, both host-side Python like the reference's own drivers:

`StreamServer` -- what turns "streams per GPU" into a service loop:
  * session admission with a bounded wait queue: `admit()` opens a session when a slot is free, queues the request when the
    pool is full, and refuses it (returns None) when the queue is full too -- back-pressure towards the caller instead of
    unbounded growth;
  * audio in either form: `feed_mel()` (log-mel frames) or `feed_pcm()` (16 kHz PCM; a per-stream sample buffer feeds the GPU
    log-mel front-end `conan_logmel` whenever whole frames are available).  Both report how much was ACCEPTED: a stream whose
    scheduler ring is full (the consumer is behind) is not fed further -- per-stream input back-pressure;
  * `pump()` = one scheduler step for every ready stream (synchronous, or pipelined with two steps in flight), whose wav lands
    in a per-stream output jitter buffer that `read()` drains in arbitrary-sized pieces;
  * `end()` / automatic close: a stream whose input ended and whose output has been produced releases its slot, and the next
    queued admission takes it.

`VoiceConversionRunner` -- drop-in for `inference/run_voice_conversion_nvae.py:15-176`: same JSON config
(`total_pairs`, `conversion_pairs[*].{ref_wav, src_wav, src_corpus, src_utt_id, output_name}`), same output naming
(`{src_corpus}_{src_utt_id}.wav` under `output_dir`), same `conversion_progress.json` / `final_report.json` keys, same
per-pair error capture.  The difference is underneath: pairs are converted CONCURRENTLY, as many sessions in flight as the
engine has slots, one packed launch sequence per 80 ms of all of them, instead of one utterance at a time.
"""
from __future__ import annotations

import json
import os
import time
from collections import deque
from dataclasses import dataclass, field
from datetime import datetime
from typing import Deque, Dict, List, Optional, Tuple

import numpy as np

from .scheduler import ChunkScheduler


@dataclass
class _Session:
    sid: int
    pcm: Optional[np.ndarray] = None                 # samples received but not yet framed (feed_pcm)
    pcm_received: int = 0                            # samples received in all
    frames_made: int = 0                             # mel frames produced from PCM so far
    out: Deque[np.ndarray] = field(default_factory=deque)    # output jitter buffer: wav pieces not yet read
    out_samples: int = 0
    mel_out: List[np.ndarray] = field(default_factory=list)
    ended: bool = False
    tag: object = None


class StreamServer:
    def __init__(self, engine, max_streams: int, *, frontend=None, max_queue: int = 64, keep_mel: bool = False,
                 pipelined: bool = False, capacity_frames: int = 64):
        self.sch = ChunkScheduler(engine, max_streams, capacity_frames=capacity_frames)
        self.fe = frontend                               # conan_b200.frontend.GpuLogMel (only needed for feed_pcm)
        self.max_queue = max_queue
        self.keep_mel = keep_mel
        self.pipelined = pipelined and hasattr(engine, "step_host_submit")
        self.sessions: Dict[int, _Session] = {}
        self.queue: Deque[Tuple[int, np.ndarray, object]] = deque()     # (ticket, ref_mel, tag)
        self.admitted: Dict[int, int] = {}               # ticket -> sid once a queued admission got its slot
        self._ticket = 0
        self._pending = None                             # ticket of the scheduler step in flight (pipelined mode)
        self._ending = set()                             # streams whose input ended and whose slot is not yet released
        self.stats = {"admitted": 0, "queued": 0, "refused": 0, "closed": 0, "steps": 0, "stream_chunks": 0,
                      "input_backpressure_events": 0}

    # ------------------------------------------------------------------ admission
    def admit(self, ref_mel: np.ndarray, tag=None) -> Optional[int]:
        """Returns a ticket (resolve it with `session_of(ticket)`), or None when the wait queue is full (back-pressure)."""
        ref_mel = np.asarray(ref_mel, dtype=np.float32)
        t = self._ticket
        if self.sch.free and not self.queue:
            self._ticket += 1
            self._open(t, ref_mel, tag)
            return t
        if len(self.queue) >= self.max_queue:
            self.stats["refused"] += 1
            return None
        self._ticket += 1
        self.queue.append((t, ref_mel, tag))
        self.stats["queued"] += 1
        return t

    def _open(self, ticket: int, ref_mel: np.ndarray, tag):
        sid = self.sch.open(ref_mel)                     # raises on a malformed reference; nothing is leaked (scheduler.open_many)
        self.sessions[sid] = _Session(sid=sid, tag=tag)
        self.admitted[ticket] = sid
        self.stats["admitted"] += 1

    def session_of(self, ticket: int) -> Optional[int]:
        """Stream id of an admission ticket, or None while it is still waiting for a slot."""
        return self.admitted.get(ticket)

    def _drain_queue(self):
        while self.queue and self.sch.free:
            t, ref, tag = self.queue.popleft()
            try:
                self._open(t, ref, tag)
            except Exception as ex:                      # a bad reference must not wedge the queue
                self.admitted[t] = -1
                print(f"| admission {t} failed: {ex}")

    # ------------------------------------------------------------------ input
    def _room(self, sid: int) -> int:
        return self.sch.cap - self.sch.buffered_frames(sid) if not self.sch._backlog.get(self.sch.streams[sid].slot) else 0

    def feed_mel(self, sid: int, frames: np.ndarray) -> int:
        """Accepts as many leading frames as the stream's ring has room for; returns that count (0 = back-pressure)."""
        frames = np.asarray(frames, dtype=np.float32)
        n = min(self._room(sid), frames.shape[0])
        if n < frames.shape[0]:
            self.stats["input_backpressure_events"] += 1
        if n > 0:
            self.sch.push(sid, frames[:n])
        return n

    def feed_pcm(self, sid: int, pcm: np.ndarray, final: bool = False) -> int:
        """16 kHz PCM (float32 in [-1, 1]) -> log-mel frames on the GPU (`conan_logmel`), streamed: frame f is produced as soon as
        its samples [f*hop - n_fft/2, f*hop + n_fft/2) exist (the reference computes the whole utterance offline with the same
        centre-padded STFT, utils/audio/__init__.py:62-72).  Returns the number of SAMPLES accepted (all of them unless the
        stream's frame ring is full).  `final=True` flushes the tail frames (zero centre padding) and ends the stream."""
        import torch
        if self.fe is None:
            raise RuntimeError("feed_pcm needs a GpuLogMel front-end")
        fe, s = self.fe, self.sessions[sid]
        pcm = np.asarray(pcm, dtype=np.float32).reshape(-1)
        room_frames = self._room(sid)
        # samples we may take without producing more frames than there is room for
        max_new = max(0, (s.frames_made + room_frames) * fe.hop + fe.n_fft - fe.pad - s.pcm_received - 1) if not final else pcm.shape[0]
        take = min(pcm.shape[0], max_new)
        if take < pcm.shape[0]:
            self.stats["input_backpressure_events"] += 1
            final = False
        if take:
            s.pcm = pcm[:take].copy() if s.pcm is None else np.concatenate([s.pcm, pcm[:take]])
            s.pcm_received += take
        # frames now computable: frame f needs padded samples [f*hop, f*hop + n_fft) of (pad zeros | signal | zeros)
        if final:
            avail = fe.n_frames_for(s.pcm_received)
        else:
            avail = (s.pcm_received + fe.pad - fe.n_fft) // fe.hop + 1 if s.pcm_received + fe.pad >= fe.n_fft else 0
        n_new = avail - s.frames_made
        if n_new > 0:
            # the rows the new frames read: padded-signal samples [frames_made*hop, (avail - 1)*hop + n_fft)
            start = s.frames_made * fe.hop                       # in padded coordinates
            rows = n_new + fe.taps - 1
            sig = np.zeros(rows * fe.hop, np.float32)
            base = s.pcm_received - (0 if s.pcm is None else s.pcm.shape[0])      # signal index of s.pcm[0]
            lo, hi = start - fe.pad, start - fe.pad + rows * fe.hop              # signal index range wanted
            a, b = max(lo, base, 0), min(hi, s.pcm_received)
            if b > a:
                sig[a - lo:b - lo] = s.pcm[a - base:b - base]
            mel = fe.frames(torch.from_numpy(sig).to(fe.device).view(1, rows, fe.hop), 0, n_new)[0].cpu().numpy()
            self.sch.push(sid, mel)
            s.frames_made = avail
            # samples before the next frame's window are no longer needed
            keep_from = max(base, s.frames_made * fe.hop - fe.pad)
            s.pcm = s.pcm[keep_from - base:]
        if final:
            self.end(sid)
        return take

    def end(self, sid: int):
        self.sessions[sid].ended = True
        self.sch.end(sid)
        self._ending.add(sid)

    # ------------------------------------------------------------------ stepping / output
    def _deliver(self, r):
        if r is None:
            return 0
        for i, sid in enumerate(r.sids):
            s = self.sessions[int(sid)]
            e = int(r.emits[i])
            w = r.wav[i, :e * r.hop].copy()
            s.out.append(w)
            s.out_samples += w.shape[0]
            if self.keep_mel:
                s.mel_out.append(r.mel[i, :e].copy())
        self.stats["steps"] += 1
        self.stats["stream_chunks"] += len(r)
        return len(r)

    def pump(self) -> int:
        """One scheduler step over every ready stream; returns how many stream-chunks were delivered to the output buffers.
        Pipelined mode keeps one step in flight: this call submits the next step and delivers the previous one."""
        if not self.pipelined:
            n = self._deliver(self.sch.step_packed())
        else:
            t = self.sch.submit()
            n = 0
            if self._pending is not None:
                n = self._deliver(self.sch.collect(self._pending))
            self._pending = t
        self._reap()
        return n

    def flush(self):
        """Delivers the step still in flight (pipelined mode)."""
        if self._pending is not None:
            self._deliver(self.sch.collect(self._pending))
            self._pending = None
            self._reap()

    def _reap(self):
        """A finished stream keeps its session object (its output may not have been read yet) but gives its slot back."""
        if self._ending:
            busy = set()
            for t in self.sch._inflight:                     # a stream's last chunk may still be in the step in flight
                busy.update(int(x) for x in t[4])
            for sid in [x for x in self._ending if x not in busy and self.sch.finished(x)]:
                self._ending.discard(sid)
                if sid in self.sch.streams:
                    self.sch.close(sid)
                    self.stats["closed"] += 1
        self._drain_queue()

    def available(self, sid: int) -> int:
        return self.sessions[sid].out_samples

    def read(self, sid: int, n_samples: Optional[int] = None) -> np.ndarray:
        """Drains up to n_samples (all, if None) of the stream's output jitter buffer."""
        s = self.sessions[sid]
        want = s.out_samples if n_samples is None else min(n_samples, s.out_samples)
        parts, got = [], 0
        while got < want:
            p = s.out[0]
            if p.shape[0] <= want - got:
                parts.append(s.out.popleft())
                got += p.shape[0]
            else:
                parts.append(p[:want - got])
                s.out[0] = p[want - got:]
                got = want
        s.out_samples -= got
        return np.concatenate(parts) if parts else np.zeros(0, np.float32)

    def done(self, sid: int) -> bool:
        """Input ended, everything computed (the slot has been released); unread output may remain in the jitter buffer."""
        return self.sessions[sid].ended and sid not in self.sch.streams

    def release(self, sid: int) -> Optional[np.ndarray]:
        s = self.sessions.pop(sid)
        self._ending.discard(sid)
        if sid in self.sch.streams:
            self.sch.close(sid)
            self._drain_queue()
        return np.concatenate(s.mel_out) if s.mel_out else None


# ----------------------------------------------------------------------------------------------
class VoiceConversionRunner:
    """Run voice conversion on all prepared pairs (inference/run_voice_conversion_nvae.py:15-176), many pairs in flight."""

    def __init__(self, config_file="voice_conversion_config.json", hparams=None, engine=None, output_dir="test_output_nvae_conan",
                 max_concurrent: Optional[int] = None):
        self.config_file = config_file
        self.config = self.load_config()
        self.output_dir = output_dir
        self.setup_output_dir()
        self.hparams = hparams
        if engine is None:
            from .streaming import StreamingVoiceConversion
            print("Initializing StreamingVoiceConversion engine...")
            engine = StreamingVoiceConversion(hparams, max_streams=max_concurrent or 64)
            print("Engine initialized successfully!")
        self.engine = engine
        self.max_concurrent = max_concurrent

    def load_config(self):
        if not os.path.exists(self.config_file):
            raise FileNotFoundError(f"Configuration file {self.config_file} not found")
        with open(self.config_file, "r") as f:
            config = json.load(f)
        print(f"Loaded {config['total_pairs']} conversion pairs")
        return config

    def setup_output_dir(self):
        os.makedirs(self.output_dir, exist_ok=True)
        print(f"Output directory: {self.output_dir}")

    def _save(self, pair, wav) -> str:
        from . import audio
        output_name = f"{pair['src_corpus']}_{pair['src_utt_id']}.wav"
        output_path = os.path.join(self.output_dir, output_name)
        audio.save_wav(wav, output_path, self.hparams["audio_sample_rate"])
        return output_path

    def run_single_conversion(self, pair, pair_idx, spk_emb=None):
        """One pair through `infer_once`, as the reference does (errors are captured, not raised)."""
        try:
            wav_pred, _ = self.engine.infer_once({"ref_wav": pair["ref_wav"], "src_wav": pair["src_wav"]})
            return True, self._save(pair, wav_pred)
        except Exception as e:
            print(f"ERROR in pair {pair_idx}: {str(e)}")
            return False, str(e)

    def run_all_conversions(self, start_idx=0, end_idx=None, batch_size=50, embs=None):
        """All pairs [start_idx, end_idx), concurrently: a pair is admitted as soon as a slot is free, every scheduler step
        advances all admitted pairs by one chunk, a finished pair is saved and its slot reused.  Same progress / report files
        as the reference (`embs` is accepted and ignored: inference/Conan.py::infer_once takes no speaker embedding)."""
        pairs = self.config["conversion_pairs"]
        total_pairs = len(pairs)
        end_idx = total_pairs if end_idx is None else end_idx
        n_total = end_idx - start_idx
        print(f"Running conversions from {start_idx} to {end_idx} (total: {n_total})")
        eng = self.engine
        sch = eng.scheduler
        limit = min(self.max_concurrent or sch.S, sch.S)
        successful = failed = processed = 0
        errors: List[str] = []
        start_time = time.time()
        progress_file = os.path.join(self.output_dir, "conversion_progress.json")
        nxt = start_idx
        live: Dict[int, Tuple[int, List[np.ndarray]]] = {}             # sid -> (pair index, wav pieces)

        def finish(i, ok, result):
            nonlocal successful, failed, processed
            processed += 1
            if ok:
                successful += 1
                print(f"  [{i + 1}/{total_pairs}] saved: {result}")
            else:
                failed += 1
                errors.append(f"Pair {i}: {result}")
                print(f"  [{i + 1}/{total_pairs}] failed: {result}")
            if processed % batch_size == 0:
                elapsed = time.time() - start_time
                avg = elapsed / processed
                with open(progress_file, "w") as f:
                    json.dump({"processed": processed, "total": n_total, "successful": successful, "failed": failed,
                               "elapsed_time": elapsed, "estimated_remaining": (n_total - processed) * avg,
                               "current_batch_end": i, "errors": errors}, f, indent=2)

        while nxt < end_idx or live:
            while nxt < end_idx and len(live) < limit and sch.free:      # admission
                i, pair = nxt, pairs[nxt]
                nxt += 1
                try:
                    ref_mel = eng._wav_to_mel(pair["ref_wav"])
                    src_mel = eng._wav_to_mel(pair["src_wav"])
                    sid = sch.open(ref_mel)
                    sch.push(sid, src_mel)
                    sch.end(sid)
                    live[sid] = (i, [])
                except Exception as e:
                    print(f"ERROR in pair {i}: {str(e)}")
                    finish(i, False, str(e))
            if not live:
                continue
            r = sch.step_packed()
            if r is not None:
                for k, sid in enumerate(r.sids):
                    live[int(sid)][1].append(r.wav[k, :int(r.emits[k]) * r.hop].copy())
            for sid in [s for s in live if sch.finished(s)]:
                i, pieces = live.pop(sid)
                sch.close(sid)
                try:
                    finish(i, True, self._save(pairs[i], np.concatenate(pieces) if pieces else np.zeros(0, np.float32)))
                except Exception as e:
                    finish(i, False, str(e))
        total_time = time.time() - start_time
        denom = max(n_total, 1)
        print(f"\n=== Final Summary ===\nTotal processed: {n_total}\nSuccessful: {successful}\nFailed: {failed}")
        print(f"Success rate: {100 * successful / denom:.1f}%\nTotal time: {total_time / 60:.1f} minutes")
        print(f"Average time per file: {total_time / denom:.1f}s\nOutput directory: {self.output_dir}")
        final_report = {"start_idx": start_idx, "end_idx": end_idx, "total_processed": n_total, "successful": successful,
                        "failed": failed, "success_rate": 100 * successful / denom, "total_time_minutes": total_time / 60,
                        "avg_time_per_file_seconds": total_time / denom, "output_directory": self.output_dir,
                        "errors": errors, "timestamp": datetime.now().isoformat()}
        with open(os.path.join(self.output_dir, "final_report.json"), "w") as f:
            json.dump(final_report, f, indent=2)
        if failed > 0:
            print("\nErrors encountered:")
            for error in errors[:10]:
                print(f"  {error}")
            if len(errors) > 10:
                print(f"  ... and {len(errors) - 10} more errors")
        return final_report


def main():
    """`python -m conan_b200.serving --config ... --exp_name ...` : the reference's `main()` (run_voice_conversion_nvae.py:165-176)."""
    from .hparams import hparams, set_hparams
    set_hparams()
    runner = VoiceConversionRunner("voice_conversion_config.json", hparams=hparams)
    runner.run_all_conversions(start_idx=0, end_idx=None, batch_size=50)


if __name__ == "__main__":
    main()
