#!/usr/bin/env python
"""bench.py -- concurrent real-time 80 ms-chunk voice streams on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference] [--lean]

A "step" is one pass of the whole chunk path (Emformer step -> proj/argmax -> Conan chunk
decoder -> causal shuffle HiFi-GAN) over one 80 ms chunk of every resident stream: S = 1024
streams per GPU (BASELINE.json configs[3] at N = 1; N x 1024 = configs[4] at N = 8; weak
scaling, streams are independent, there is no collective on the data path).

  value    = (stream-chunks processed per second, all ranks) x 0.08 s, measured over a SUSTAINED window: the K-step
             block is repeated until >= --min-seconds (3 s) of device time has elapsed; `burst` holds the first K steps
             alone.  p50 / p99 step latency come from every step of that window (>= 500 samples).
  e2e      = the same through the plugin calls conan_step_host_submit / _wait (two steps in flight): slot ids +
             mel chunks copied from pinned host memory, wav + mel + tokens copied back, every step, inside the
             timed region; the synchronous conan_step_host figure is reported next to it.
  roofline = the dominant kernel family: algorithmic FLOPs / CUDA-event time of its launches (engine profiling hooks,
             separate short pass) against the measured BURST 16-bit tensor peak of MEASURED_PEAKS.json;
             `path` scores the sustained window against the SUSTAINED peak.
  cpu_baseline / --impl reference: the CPU oracle (oracle/incremental.py, the incremental PyTorch restatement of the
             reference's loop; the reference itself is Python and /root/reference does not exist on the GPU box).

Extra keys of the N = 1 line (skipped with --lean, and on multi-GPU runs except `config5`):
  sweep       resident-slot sweep (2048 ... 14336 slots really allocated, sessions really opened): ms/step, p50/p99,
              the largest S with p99 < 20 ms and the largest S that is still real-time (step < 80 ms)
  scheduler   ChunkScheduler in the loop: lock-step pipelined throughput, and a real-time run with jittered arrivals
              (every stream delivers 4 frames per 80 ms of wall clock at its own phase) reporting arrival-to-wav p50/p99
  session     session setup: ms per batch of 64 sessions (T_ref 150); throughput with a fraction of the slots re-opened
              every step inside the timed window (churn)
  config2     Emformer only, 64 streams (BASELINE.json configs[1]); lock-step and staggered ages
  config3     vocoder only, 256 streams (configs[2])
  fp32_grade  the same workload on the fp32-grade engines: split-fp16 operands on tcgen05 (voc_precision split) and
              fp32 FFMA; the headline's fp16-operand vocoder is the reduced-precision variant north_star words separately
  torch_cuda  the oracle's PyTorch restatement on the same B200 through torch CUDA (cuDNN / cuBLAS eager), same workload
  config5     (N > 1 too) 8192 streams sharded over the N GPUs (8192 / N resident slots per GPU)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_S = 0.08
FLOP_PER_STREAM_CHUNK = 2.630e9
FLOP_EMFORMER, FLOP_VOCODER = 26.15e6, 2.527e9
WORKLOAD = "full Conan pipeline (style encoder once + chunk loop), {S} concurrent streams per B200"
SEG, ROWS, MELS, HOP = 4, 6, 80, 1280


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d.get("bf16_tflops", 1657.0), "sustained": d.get("bf16_tflops_sustained", 1380.4),
                "hbm": d.get("hbm_gbs", 6550.7), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1650.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
            try:
                pw.append(float(f[6]))
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _pct(sorted_ms, q):
    return sorted_ms[min(len(sorted_ms) - 1, int(len(sorted_ms) * q))]


def _lat(per_step):
    ps = sorted(per_step)
    return {"p50": ps[len(ps) // 2], "p99": _pct(ps, 0.99), "max": ps[-1], "samples": len(ps)}


# ----------------------------------------------------------------------------------------------
def cpu_oracle_throughput(n_streams: int, n_chunks: int, warm_chunks: int = 1, threads: int = 0):
    """Times the CPU oracle (incremental restatement of the reference loop) on a bounded sample.
    Returns (real-time streams sustained, seconds, threads)."""
    import torch
    from conan_b200 import synth
    from oracle.incremental import StreamingOracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sds = synth.make_all_state_dicts(1234)
    o = StreamingOracle(*sds)
    ref = torch.stack([synth.synth_mel(150, 100 + s) for s in range(n_streams)])
    T = (warm_chunks + n_chunks) * 4 + 2
    src = torch.stack([synth.synth_mel(T, 200 + s) for s in range(n_streams)])
    with torch.no_grad():
        o.emf.reset(n_streams)
        o.conan.open(ref)
        o.voc.reset(n_streams)
        pos = 0
        for _ in range(warm_chunks):
            o.step(src, pos)
            pos += 4
        t0 = time.perf_counter()
        for _ in range(n_chunks):
            o.step(src, pos)
            pos += 4
        dt = time.perf_counter() - t0
    return n_streams * n_chunks / dt * CHUNK_S, dt, threads


def run_reference(args):
    """--impl reference: the reference's CPU formulation (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 16
    val, dt, threads = cpu_oracle_throughput(B, args.steps, warm_chunks=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "concurrent real-time 80 ms-chunk streams", "value": val, "unit": "streams",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(S=1024), "note": "CPU arm: each step is one 80 ms chunk of a bounded "
                   f"sample of {B} lock-step streams of that workload (same weights, same synthetic mel)"},
        "cpu_baseline": {"value": val, "unit": "streams", "cores": threads, "kind": "port",
                         "sample": f"{B} streams x {args.steps} chunks, oracle/incremental.py (incremental restatement "
                                   "of inference/Conan.py:95-156; the literal O(T^2) loop is slower still)"},
        "e2e": {"value": val, "unit": "streams", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
class Rig:
    """One engine with S resident streams (sessions opened), a pool of device-resident synthetic chunks, and timing helpers."""

    def __init__(self, S, local, sds, *, voc_precision="fp16", tensor_cores=True, voc_group=0, fuse=None, open_sessions=True,
                 lin_tensor_cores=None, max_ref_frames=160):
        import numpy as np
        import torch
        from conan_b200 import synth
        from conan_b200.engine import Engine, make_config
        self.torch, self.np, self.S = torch, np, S
        self.dev = torch.device("cuda", local)
        self.cfg = make_config(max_slots=S, max_ref_frames=max_ref_frames, device=local, voc_precision=voc_precision,
                               voc_tensor_cores=tensor_cores, voc_group=voc_group, voc_fuse_resblocks=fuse,
                               lin_tensor_cores=lin_tensor_cores)
        self.eng = Engine(*sds, self.cfg)
        self.slots = np.arange(S, dtype=np.int32)
        self.ids = self.eng.ids_tensor(self.slots)
        self.eng.reset_slots(self.slots)
        self.refs = torch.stack([synth.synth_mel(150, 100 + s) for s in range(8)]).to(self.dev)
        self.session_ms = None
        if open_sessions:
            # one session per stream: 8 distinct 3 s reference utterances, repeated; batches of 64 sessions
            torch.cuda.synchronize(self.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for g in range(0, S, 64):
                n = min(64, S - g)
                self.eng.open_sessions(self.slots[g:g + n], self.refs[torch.arange(g, g + n, device=self.dev) % 8])
            e1.record()
            torch.cuda.synchronize(self.dev)
            self.session_ms = e0.elapsed_time(e1)
        # synthetic source mel: a pool of chunks larger than one step so every step reads fresh input
        self.n_pool = 8
        pool = torch.stack([synth.synth_mel(ROWS * self.n_pool, 300 + s) for s in range(64)])          # [64, 6*n_pool, 80]
        self.chunks = [pool[:, ROWS * i:ROWS * i + ROWS].repeat(S // 64 + 1, 1, 1)[:S].contiguous().to(self.dev)
                       for i in range(self.n_pool)]
        self.wav = torch.empty(S, self.eng.hop_out, device=self.dev)
        self.mel = torch.empty(S, SEG, MELS, device=self.dev)
        self.tok = torch.empty(S, SEG, dtype=torch.int32, device=self.dev)
        self._i = 0

    def step(self):
        self.eng.step(self.ids, self.chunks[self._i % self.n_pool], self.wav, self.mel, self.tok)
        self._i += 1

    def timed(self, K, fn=None):
        """K steps with a CUDA event after each: (per-step ms list, total ms)."""
        torch = self.torch
        fn = fn or self.step
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        evs[0].record()
        for i in range(K):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize(self.dev)
        return [evs[i].elapsed_time(evs[i + 1]) for i in range(K)], evs[0].elapsed_time(evs[K])

    def sustained(self, K, blocks, fn=None):
        per, total = [], 0.0
        for _ in range(blocks):
            p, t = self.timed(K, fn)
            per += p
            total += t
        return per, total

    def close(self):
        self.eng.close()
        del self.eng, self.chunks, self.wav, self.mel, self.tok, self.refs
        self.torch.cuda.empty_cache()


def _short_window(rig, K, seconds, warm=3):
    """warm-up, one K-step probe, then K-step blocks for ~`seconds`: dict(ms_per_step, latency, value, steps)."""
    for _ in range(warm + (4 * rig.n_pool if rig.cfg.step_graphs else 0)):
        rig.step()
    rig.torch.cuda.synchronize(rig.dev)
    _, probe = rig.timed(K)
    blocks = max(1, int(seconds * 1e3 / max(probe, 1e-3) + 0.999))
    per, total = rig.sustained(K, blocks)
    return {"streams_resident": rig.S, "steps": len(per), "seconds": total * 1e-3, "ms_per_step": total / len(per),
            "value": rig.S * len(per) / (total * 1e-3) * CHUNK_S, "latency_ms": _lat(per),
            "realtime": total / len(per) < CHUNK_S * 1e3, "state_gib": rig.eng.state_bytes / 2 ** 30}


# ---------------------------------------------------------------------------------------------- extra legs
def leg_sweep(local, sds, sizes, K):
    out = []
    for S in sizes:
        try:
            rig = Rig(S, local, sds)
            r = _short_window(rig, K, 1.5)
            r["session_open_ms_total"] = rig.session_ms
            if S == 8192:            # config 5 at N = 1 as a SERVICE: 8192 real-time streams with jittered arrivals on one GPU
                r["realtime_jittered"] = realtime_leg(rig, 3.0, warm_batch=256)
            rig.close()
        except Exception as ex:          # e.g. out of memory at the largest size: report, keep the line
            r = {"streams_resident": S, "error": str(ex)[:200]}
        out.append(r)
    ok = [r for r in out if "error" not in r]
    p99_ok = [r["streams_resident"] for r in ok if r["latency_ms"]["p99"] < 20.0]
    rt_ok = [r["streams_resident"] for r in ok if r["latency_ms"]["p99"] < CHUNK_S * 1e3]
    return {"runs": out, "largest_streams_with_p99_below_20ms": max(p99_ok) if p99_ok else None,
            "largest_streams_realtime_p99_below_80ms": max(rt_ok) if rt_ok else None,
            "note": "every run allocates S slots, opens S sessions and steps all S streams; latency = device time of one packed step"}


def _adopt_sessions(rig):
    """A ChunkScheduler over the rig's already opened sessions: slot i <-> stream i (session setup is measured separately)."""
    from conan_b200.scheduler import ChunkScheduler
    sch = ChunkScheduler(rig.eng, rig.S)
    sch.free = []
    for slot in range(rig.S):
        sch._activate(slot, slot)
    sch._next_id = rig.S
    rig.eng.reset_slots(rig.slots, 1 | 2 | 4)        # zero the stream state; the session caches stay
    return sch


def _frame_pool(S, P=16):
    import numpy as np
    from conan_b200 import synth
    pool = np.stack([synth.synth_mel(SEG * P, 500 + s).numpy() for s in range(32)])                   # [32, 4P, 80]
    return np.ascontiguousarray(np.tile(pool, (S // 32 + 1, 1, 1))[:S].reshape(S, P, SEG, MELS))


def realtime_leg(rig, seconds=3.0, warm_batch=96):
    """Every stream delivers 4 frames per 80 ms of WALL CLOCK at its own phase with exponential jitter; the loop ingests what has
    arrived, runs one packed step for whatever is ready (ChunkScheduler.step_packed: results on the host) and records, per
    stream-chunk, the time from the arrival that completed the chunk to its wav being on the host."""
    import numpy as np
    from conan_b200.scheduler import ChunkScheduler
    eng, S = rig.eng, rig.S
    # start-up warm-up, as a server would do it: the step graphs of the ready-count buckets are captured on throw-away steps
    warm = ChunkScheduler(eng, S)
    warm.warm(max_batch=warm_batch)
    del warm
    sch = _adopt_sessions(rig)
    P = 16
    pool = _frame_pool(S, P)
    rng = np.random.default_rng(0)
    n_ev = int(seconds / CHUNK_S)
    phase = rng.uniform(0, CHUNK_S, S)
    jitter = rng.exponential(0.004, (S, n_ev))                          # ~Poisson network jitter, mean 4 ms
    t_arr = phase[:, None] + CHUNK_S * np.arange(n_ev)[None, :] + jitter      # arrival k of stream s: 4 new frames
    order = np.argsort(t_arr, axis=None)
    ev_t, ev_s, ev_k = t_arr.reshape(-1)[order], (order // n_ev), (order % n_ev)
    last_arrival = np.zeros(S)
    lat, batch_sizes = [], []
    nxt, N = 0, len(ev_t)
    t0 = time.perf_counter()
    while nxt < N or sch.ready_slots().size:
        now = time.perf_counter() - t0
        hi = np.searchsorted(ev_t, now, side="right")
        if hi > nxt:
            s_, k_ = ev_s[nxt:hi], ev_k[nxt:hi]
            for k in np.unique(k_):                                       # a stream appears once per arrival index
                m = k_ == k
                sch.push_many(s_[m], pool[s_[m], k % P])
            last_arrival[s_] = ev_t[nxt:hi]                               # (the latest one, if a stream arrived twice)
            nxt = hi
        r = sch.step_packed()
        if r is None:
            if nxt < N:
                time.sleep(max(0.0, min(0.0005, ev_t[nxt] - (time.perf_counter() - t0))))
            continue
        done = time.perf_counter() - t0
        lat.append(done - last_arrival[r.slots])
        batch_sizes.append(len(r))
    total = time.perf_counter() - t0
    lat = np.sort(np.concatenate(lat)) * 1e3
    return {"api": "ChunkScheduler.push_many / step_packed (synchronous, results on the host)", "streams": S, "seconds": total,
            "arrival_model": "every stream delivers 4 frames per 80 ms of wall clock, uniform phase, exponential jitter (mean 4 ms)",
            "chunk_steps": int(len(lat)), "engine_calls": len(batch_sizes), "mean_batch": float(np.mean(batch_sizes)),
            "max_batch": int(np.max(batch_sizes)),
            "arrival_to_wav_ms": {"p50": float(lat[len(lat) // 2]), "p99": float(lat[int(len(lat) * 0.99)]), "max": float(lat[-1]),
                                  "samples": int(len(lat))},
            "kept_up": bool(total < seconds + 0.25), "graph_replays_total": eng.graph_replays}


def leg_scheduler(rig, K, seconds=3.0):
    """ChunkScheduler in the measured loop: (a) lock-step pipelined throughput through push_many / submit / collect,
    (b) real time with jittered arrivals: arrival-to-wav latency through step_packed."""
    import numpy as np
    S = rig.S
    sch = _adopt_sessions(rig)
    P = 16
    pool = _frame_pool(S, P)
    slots = np.arange(S)
    sch.push_many(slots, np.ascontiguousarray(pool[:, 0, :2]))          # prime the look-ahead (rc = 2 frames)
    prev, n_steps = None, 0
    t0 = time.perf_counter()
    while True:
        sch.push_many(slots, pool[:, n_steps % P])
        t = sch.submit()
        if prev is not None:
            sch.collect(prev)
        prev = t
        n_steps += 1
        if n_steps >= 3 * K and time.perf_counter() - t0 > min(seconds, 2.0):
            break
    sch.collect(prev)
    rig.torch.cuda.synchronize(rig.dev)
    dt = time.perf_counter() - t0
    lock = {"api": "ChunkScheduler.push_many / submit / collect (two steps in flight)", "steps": n_steps, "seconds": dt,
            "ms_per_step": dt / n_steps * 1e3, "value": S * n_steps / dt * CHUNK_S, "clock": "host wall clock"}
    return {"lockstep_pipelined": lock, "realtime_jittered": realtime_leg(rig, seconds)}


def leg_session(rig, K):
    """Session setup cost and throughput under churn (slots re-opened inside the timed window)."""
    torch, np, eng, S = rig.torch, rig.np, rig.eng, rig.S
    n = min(64, S)
    evs = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.open_sessions(rig.slots[:n], rig.refs[torch.arange(n, device=rig.dev) % 8])
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(rig.dev)
    ms = sorted(a.elapsed_time(b) for a, b in evs[1:])
    out = {"batch": n, "ref_frames": 150, "session_open_ms": ms[len(ms) // 2], "gflop_per_session": 12.8,
           "tflops": n * 12.8e9 / (ms[len(ms) // 2] * 1e-3) / 1e12,
           "tensor_cores": bool(rig.cfg.ses_use_tensor_cores), "all_sessions_open_ms": rig.session_ms}
    churn = {}
    for frac in (0.0027, 0.01):           # 0.27 % per step = 30 s sessions at 80 ms chunks; 1 % = 8 s sessions
        k = max(1, int(round(S * frac)))
        state = {"i": 0}

        def fn():
            lo = (state["i"] * k) % (S - k + 1)
            sl = rig.slots[lo:lo + k]
            eng.reset_slots(sl)
            eng.open_sessions(sl, rig.refs[torch.arange(k, device=rig.dev) % 8])
            rig.step()
            state["i"] += 1
        for _ in range(2):
            fn()
        per, total = rig.timed(max(K, 20), fn)
        key = f"{frac * 100:.2f}%_of_slots_per_step"
        churn[key] = {"sessions_opened_per_step": k, "same_stream": {"ms_per_step": total / len(per),
                      "value": S * len(per) / (total * 1e-3) * CHUNK_S, "latency_ms": _lat(per)}}
        # the serving arrangement: session setup on a side stream while the chunk step of the OTHER streams runs; a re-opened
        # slot joins the ready list at the next step (event dependency), as a new session would
        side = torch.cuda.Stream(device=rig.dev)
        main = torch.cuda.current_stream(rig.dev)
        ev_open = torch.cuda.Event()
        state = {"i": 0, "pending": None}
        all_slots = rig.slots

        def fn2():
            lo = (state["i"] * k) % (S - k + 1)
            sl = all_slots[lo:lo + k]
            side.wait_stream(main)                           # the slots being re-opened were last used by the previous step
            with torch.cuda.stream(side):
                eng.reset_slots(sl)
                eng.open_sessions(sl, rig.refs[torch.arange(k, device=rig.dev) % 8])
                ev_open.record(side)
            rest = np.concatenate([all_slots[:lo], all_slots[lo + k:]])
            ids_rest = state.get("ids", {}).get(lo)
            if ids_rest is None:
                ids_rest = eng.ids_tensor(rest)
                state.setdefault("ids", {})[lo] = ids_rest
            ch = rig.chunks[rig._i % rig.n_pool]
            eng.step(ids_rest, ch[:S - k], rig.wav[:S - k], rig.mel[:S - k], rig.tok[:S - k])
            rig._i += 1
            main.wait_event(ev_open)                         # next step may use the re-opened slots
            state["i"] += 1
        for _ in range(3):
            fn2()
        per, total = rig.timed(max(K, 20), fn2)
        churn[key]["side_stream"] = {"ms_per_step": total / len(per), "value": (S - k) * len(per) / (total * 1e-3) * CHUNK_S,
                                     "latency_ms": _lat(per)}
    out["churn"] = churn
    return out


def leg_config2(local, sds, K):
    """BASELINE.json configs[1]: Emformer content extractor only, 64 concurrent streams."""
    import numpy as np
    import torch
    from conan_b200 import synth
    S = 64
    rig = Rig(S, local, sds, open_sessions=False)
    eng = rig.eng

    def lock():
        eng.emformer_step(rig.ids, rig.chunks[rig._i % rig.n_pool])
        rig._i += 1
    for _ in range(5):
        lock()
    per, total = rig.sustained(50, 10, lock)
    res = {"streams": S, "api": "conan_emformer_step", "lockstep": {"ms_per_step": total / len(per), "latency_ms": _lat(per),
           "value": S * len(per) / (total * 1e-3) * CHUNK_S, "rtf": (total / len(per)) / (CHUNK_S * 1e3)}}
    # staggered ages: four groups of 16 streams, a group joins / leaves every few steps (per-stream past_len differs)
    groups = [eng.ids_tensor(rig.slots[g * 16:(g + 1) * 16]) for g in range(4)]
    ch16 = [c[:16].contiguous() for c in rig.chunks]
    ch32 = [c[:32].contiguous() for c in rig.chunks]
    ids32 = eng.ids_tensor(rig.slots[16:48])

    def stag():
        i = rig._i
        if i % 3 == 0:
            eng.emformer_step(groups[i % 4], ch16[i % rig.n_pool])
        elif i % 3 == 1:
            eng.emformer_step(ids32, ch32[i % rig.n_pool])
        else:
            eng.emformer_step(rig.ids, rig.chunks[i % rig.n_pool])
        rig._i += 1
    per, total = rig.sustained(60, 5, stag)
    n_sc = sum(16 if i % 3 == 0 else (32 if i % 3 == 1 else 64) for i in range(len(per)))
    res["staggered_ages"] = {"ms_per_step": total / len(per), "latency_ms": _lat(per), "stream_chunks": n_sc,
                             "value": n_sc / (total * 1e-3) * CHUNK_S}
    res["tflops"] = S * FLOP_EMFORMER / (res["lockstep"]["ms_per_step"] * 1e-3) / 1e12
    rig.close()
    return res


def leg_config1(local, sds):
    """BASELINE.json configs[0] on the GPU: ONE stream, 80 ms chunks, full pipeline.  Chunk latency two ways: device time of one step
    (chunk resident in HBM) and the wall clock of the synchronous plugin call conan_step_host (pinned host chunk in, wav + mel + tokens
    back on the host when it returns): what a single caller waits per chunk.  RTF = latency / 80 ms."""
    import numpy as np
    import torch
    rig = Rig(1, local, sds)
    for _ in range(40):
        rig.step()
    torch.cuda.synchronize(rig.dev)
    per, total = rig.sustained(100, 5)
    eng = rig.eng
    h_chunks = [c.cpu().pin_memory().numpy() for c in rig.chunks]
    out = (torch.empty(1, eng.hop_out).pin_memory().numpy(), torch.empty(1, SEG, MELS).pin_memory().numpy(),
           torch.empty(1, SEG, dtype=torch.int32).pin_memory().numpy())
    slots = np.zeros(1, np.int32)
    wall = []
    for i in range(540):
        t0 = time.perf_counter()
        eng.step_host(slots, h_chunks[i % rig.n_pool], *out)
        wall.append((time.perf_counter() - t0) * 1e3)
    wall = wall[40:]
    res = {"streams": 1, "device": {"ms_per_step": total / len(per), "latency_ms": _lat(per), "rtf": (total / len(per)) / (CHUNK_S * 1e3)},
           "host_call": {"api": "conan_step_host (synchronous; wall clock around the call)", "ms_per_chunk": sum(wall) / len(wall),
                         "latency_ms": _lat(wall), "rtf": (sum(wall) / len(wall)) / (CHUNK_S * 1e3)},
           "graph_replays": eng.graph_replays}
    rig.close()
    return res


def leg_config3(local, sds, K):
    """BASELINE.json configs[2]: causal shuffle HiFi-GAN vocoder only, 256 concurrent streams, mel ~ N(0, 0.6^2)."""
    import torch
    S = 256
    rig = Rig(S, local, sds, open_sessions=False)
    g = torch.Generator().manual_seed(3)
    mels = [(torch.randn(S, SEG, MELS, generator=g) * 0.6).to(rig.dev) for _ in range(8)]

    def fn():
        rig.eng.vocoder_step(rig.ids, mels[rig._i % 8])
        rig._i += 1
    for _ in range(5):
        fn()
    per, total = rig.sustained(50, 12, fn)
    res = {"streams": S, "api": "conan_vocoder_step", "ms_per_step": total / len(per), "latency_ms": _lat(per),
           "value": S * len(per) / (total * 1e-3) * CHUNK_S, "rtf": (total / len(per)) / (CHUNK_S * 1e3),
           "tflops": S * FLOP_VOCODER / (total / len(per) * 1e-3) / 1e12}
    rig.close()
    return res


def leg_fp32_grade(local, sds, S, K):
    """The fp32-grade engines on the headline workload (north_star: fp32 primary, reduced precision reported separately)."""
    out = {}
    rig = Rig(S, local, sds, voc_precision="split")
    r = _short_window(rig, K, 1.5)
    r["engine"] = ("split-fp16 operands (x_hi*W_hi + x_hi*W_lo + x_lo*W_hi), fp32 accumulate, tcgen05, everywhere: wav max-abs "
                   "<= 1e-4 vs the reference (tests/test_gpu_parity.py::test_vocoder_split_fp16_tensor_cores_is_fp32_grade)")
    r["path_tflops_algorithmic"] = r["value"] / CHUNK_S * FLOP_PER_STREAM_CHUNK / 1e12
    out["tensor_core_split_fp16"] = r
    rig.close()
    Sf = min(S, 256)
    rig = Rig(Sf, local, sds, voc_precision="fp32", tensor_cores=False, lin_tensor_cores=False)
    for _ in range(2):
        rig.step()
    per, total = rig.timed(5)
    out["cuda_core_fp32_ffma"] = {"streams_resident": Sf, "steps": 5, "ms_per_step": total / 5, "value": Sf * 5 / (total * 1e-3) * CHUNK_S,
                                  "engine": "fp32 operands, fp32 FFMA on CUDA cores (exact-fp32 cross-check engine)"}
    rig.close()
    return out


def leg_torch_cuda(local, S, n_chunks=10):
    """The oracle's PyTorch restatement run on this B200 through torch CUDA (cuDNN / cuBLAS, eager): the number a stock
    PyTorch port of the incremental formulation would give.  A baseline leg: the oracle is the thing timed, not the product."""
    import torch
    from conan_b200 import synth
    from oracle.incremental import StreamingOracle
    dev = torch.device("cuda", local)
    sds = [{k: v.to(dev) for k, v in sd.items()} for sd in synth.make_all_state_dicts(1234)]
    res = {"streams": S, "backend": f"torch {torch.__version__} eager, cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}, "
                                    f"matmul.allow_tf32={torch.backends.cuda.matmul.allow_tf32} (stock defaults)"}
    T = (3 + n_chunks) * 4 + 2
    ref = torch.stack([synth.synth_mel(150, 100 + s) for s in range(8)]).to(dev).repeat(S // 8, 1, 1)       # (CPU generators)
    src = torch.stack([synth.synth_mel(T, 200 + s) for s in range(64)]).to(dev).repeat(S // 64 + 1, 1, 1)[:S]
    with torch.device(dev), torch.no_grad():
        o = StreamingOracle(*sds)
        o.emf.reset(S)
        t0 = time.perf_counter()
        o.conan.open(ref)
        torch.cuda.synchronize(dev)
        res["session_open_ms_all"] = (time.perf_counter() - t0) * 1e3
        o.voc.reset(S)
        pos = 0
        for _ in range(3):
            o.step(src, pos)
            pos += 4
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_chunks):
            o.step(src, pos)
            pos += 4
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    res.update(steps=n_chunks, ms_per_step=ms / n_chunks, value=S * n_chunks / (ms * 1e-3) * CHUNK_S)
    del o, sds
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from conan_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout when the communicator is created
        # (seen on the GPU box), so fd 1 points at stderr while the process group and its first collective come up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            torch.cuda.set_device(local)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, K, W = args.streams, args.steps, max(args.warmup, 3)

    sds = synth.make_all_state_dicts(1234)
    rig = Rig(S, local, sds, voc_precision=args.voc_precision, tensor_cores=not args.no_tensor_cores, voc_group=args.voc_group,
              fuse=False if args.no_fuse else None)
    eng, cfg, ids, slots = rig.eng, rig.cfg, rig.ids, rig.slots

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def allmax(*vals):
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---------------- device-resident timing
    for _ in range(W):
        rig.step()
    # untimed priming beyond the W warm-up steps: a chunk step is captured as a CUDA graph the second time its (ready count,
    # buffer set) is seen, and the input pool rotates over n_pool device buffers
    n_prime = 4 * rig.n_pool if cfg.step_graphs else 0       # (a ready count above 256 is captured at its 4th sighting)
    for _ in range(n_prime):
        rig.step()
    sync_all()
    if args.ncu_step:
        # profiler window for `ncu --profile-from-start off`: exactly one warmed-up step, no bench line
        torch.cuda.profiler.start()
        rig.step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        eng.close()
        return
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    burst_per, burst_ms = rig.timed(K)                                   # the first K steps alone (burst clocks)
    launches = eng.launch_count - l0
    sync_all()
    (burst_ms,) = allmax(burst_ms)
    # sustained window: the K-step block repeated until >= min_seconds of device time (same block count on every rank)
    blocks = max(1, int(args.min_seconds * 1e3 / max(burst_ms, 1e-3) + 0.999))
    sync_all()
    sus_per, sus_ms = rig.sustained(K, blocks)
    sync_all()
    (sus_ms,) = allmax(sus_ms)
    clocks = sampler.stop()
    n_sus = len(sus_per)

    # ---------------- end-to-end through the plugin call (host buffers: chunks in; wav + mel + tokens out)
    n_pool = rig.n_pool
    h_chunks = [torch.empty(S, ROWS, MELS).pin_memory() for _ in range(n_pool)]
    for h, d in zip(h_chunks, rig.chunks):
        h.copy_(d.cpu())
    h_out = [(torch.empty(S, eng.hop_out).pin_memory(), torch.empty(S, SEG, MELS).pin_memory(),
              torch.empty(S, SEG, dtype=torch.int32).pin_memory()) for _ in range(2)]
    np_chunks = [h.numpy() for h in h_chunks]
    np_out = [tuple(t.numpy() for t in o) for o in h_out]
    Ke = max(3, min(K, 20))
    for i in range(6):                           # untimed: a ready count above 256 is captured as a graph at its 4th sighting
        eng.step_host(slots, np_chunks[i % n_pool], *np_out[0])
    sync_all()
    # (a) synchronous plugin call, one step at a time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        eng.step_host(slots, np_chunks[i % n_pool], *np_out[0])
    e1.record()
    sync_all()
    e2e_sync_ms = e0.elapsed_time(e1)     # device timeline: includes the H2D/D2H copies and every host gap between steps
    # (b) the serving loop: submit step i, then collect step i-1 (its result copies overlap step i's compute).  Every step still
    # copies its chunks in from pinned memory and wav / mel / tokens out; e1 is recorded after the last result has landed.
    e2e_api = "conan_step_host_submit / conan_step_host_wait (two steps in flight: result copy of step i under the compute of step i+1)"
    n_e2e = max(Ke, int(args.e2e_seconds * 1e3 / max(burst_ms / K, 1e-3)))
    try:
        for i in range(10):                      # untimed: first use of the engine's copy stream and events, graph capture per buffer set
            eng.step_host_wait(eng.step_host_submit(slots, np_chunks[i % 2], *np_out[i % 2]))
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prev = None
        for i in range(n_e2e):
            t = eng.step_host_submit(slots, np_chunks[i % n_pool], *np_out[i % 2])
            if prev is not None:
                eng.step_host_wait(prev)
            prev = t
        eng.step_host_wait(prev)
        e1.record()
        sync_all()
        e2e_ms = e0.elapsed_time(e1)
    except Exception as ex:                      # never lose the bench line: report the synchronous call instead
        print(f"pipelined host stepping failed ({ex}); e2e falls back to conan_step_host", file=sys.stderr)
        e2e_ms, e2e_api, n_e2e = e2e_sync_ms, "conan_step_host", Ke
    # ---------------- roofline of the dominant kernel family (separate pass, CUDA events around every conv launch)
    eng.set_profiling(True)
    NPROF = 2
    for i in range(NPROF):
        rig.step()
    prof = {cat: eng.profile_read(cat) for cat in range(7)}
    eng.set_profiling(False)

    e2e_ms, e2e_sync_ms = allmax(e2e_ms, e2e_sync_ms)
    line = None
    if rank == 0:
        pk = _peaks()
        value = world * S * n_sus / (sus_ms * 1e-3) * CHUNK_S
        burst_value = world * S * K / (burst_ms * 1e-3) * CHUNK_S
        e2e_val = world * S * n_e2e / (e2e_ms * 1e-3) * CHUNK_S
        names = {0: ("conv_gemm_ffma_kernel (fp32 CUDA-core engine)", "tensor"),
                 1: ("conv_gemm_tc_kernel (tcgen05 implicit-GEMM causal conv, fp16 operands: vocoder scales 0-1 + upsampling)", "tensor"),
                 2: ("conv_window_tc_kernel (tcgen05, persistent, weights resident, one input window per tile: vocoder scales 2-3)", "hbm"),
                 3: ("conv_gemm_tc_kernel, split-fp16 operands (Emformer / Conan linear + conv contractions, 3 MMAs per product)", "tensor"),
                 4: ("resblock_fused_kernel (tcgen05, one HiFi-GAN residual block = six convs per launch, activations in shared memory: vocoder scales 2-3)", "tensor"),
                 5: ("ffn_fused_kernel (tcgen05, Emformer 80 -> 2048 -> 80 feed-forward in one launch, split-fp16 operands, hidden activation in shared memory)", "tensor"),
                 6: ("block_fused_kernel (tcgen05, Conan decoder residual block body / aligner feed-forward: two GEMMs per launch, split-fp16 operands)", "tensor")}
        roofs = {}
        for cat, (ms, nl, fl, by) in prof.items():
            if nl == 0:
                continue
            kname, bound = names[cat]
            tf, gb = fl / (ms * 1e-3) / 1e12, by / (ms * 1e-3) / 1e9
            r = {"kernel": kname, "bound": bound, "launches_per_step": int(nl // NPROF), "ms_per_step": ms / NPROF,
                 "algorithmic_gflop_per_step": fl / NPROF / 1e9, "algorithmic_gbyte_per_step": by / NPROF / 1e9,
                 "tflops": tf, "gbs": gb, "traffic": _measured_traffic(kname),
                 "peak_source": pk["source"] + ": burst 16-bit tensor peak (the kernels are event-timed in a short pass)"}
            if bound == "tensor":
                r.update(achieved=tf, peak=pk["burst"], unit="TFLOP/s", frac=tf / pk["burst"])
            else:
                r.update(achieved=gb, peak=pk["hbm"], unit="GB/s", frac=gb / pk["hbm"])
            roofs[cat] = r
        top = max(roofs, key=lambda c_: roofs[c_]["ms_per_step"])
        roof = dict(roofs[top])
        roof["other_kernels"] = [roofs[c_] for c_ in sorted(roofs) if c_ != top]
        path_tf = world * S * n_sus * FLOP_PER_STREAM_CHUNK / (sus_ms * 1e-3) / 1e12
        roof["path"] = {"window": "sustained", "tflops_algorithmic": path_tf / world, "peak": pk["sustained"],
                        "frac": path_tf / world / pk["sustained"], "peak_source": pk["source"] + ": sustained 16-bit tensor peak"}
        dtype = {"fp16": "split-f16 operands, f32 accumulate = f32-grade (Emformer/Conan) + f16 operands / f32 accumulate (vocoder)",
                 "split": "split-f16 operands, f32 accumulate = f32-grade (whole path)", "fp32": "f32"}[args.voc_precision]
        line = {
            "metric": "concurrent real-time 80 ms-chunk streams", "value": value, "unit": "streams", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": sus_ms / n_sus, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD.format(S=S), "streams_per_gpu": S, "chunk_ms": 80, "ref_frames": 150,
                       "weights": "synthetic seeded (conan_b200.synth, reference state_dict layout)",
                       "l2": f"per-step working set (resident state {eng.state_bytes / 2**30:.1f} GiB) is larger than L2; no flush needed",
                       "voc_precision": args.voc_precision, "voc_tensor_cores": not args.no_tensor_cores, "voc_group": args.voc_group,
                       "voc_fuse_resblocks": bool(cfg.voc_fuse_resblocks), "step_graphs": bool(cfg.step_graphs),
                       "graph_priming_steps_untimed": n_prime, "graph_replays": eng.graph_replays,
                       "value_window": f"sustained: the {K}-step block repeated {blocks}x = {n_sus} steps, {sus_ms * 1e-3:.2f} s of device time "
                                       "(value, ms_per_step, latency_ms, rtf all from this window)"},
            "sustained": {"blocks": blocks, "steps_timed": n_sus, "seconds": sus_ms * 1e-3},
            "burst": {"steps": K, "ms_per_step": burst_ms / K, "value": burst_value, "latency_ms": _lat(burst_per),
                      "sustained_over_burst": value / burst_value},
            "latency_ms": _lat(sus_per),
            "rtf": (sus_ms / n_sus) / (CHUNK_S * 1e3),
            "path_tflops": path_tf,
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "streams", "h2d_bytes_per_step": int(S * (ROWS * MELS * 4 + 4)),
                    "d2h_bytes_per_step": int(S * (eng.hop_out * 4 + SEG * MELS * 4 + SEG * 4)), "steps": n_e2e,
                    "ms_per_step": e2e_ms / n_e2e, "seconds": e2e_ms * 1e-3, "api": e2e_api, "results_copied": "wav + mel + tokens",
                    "synchronous_call": {"value": world * S * Ke / (e2e_sync_ms * 1e-3) * CHUNK_S, "ms_per_step": e2e_sync_ms / Ke,
                                         "api": "conan_step_host"}},
            "roofline": roof,
            "session": {"all_sessions_open_ms": rig.session_ms, "sessions": S, "tensor_cores": bool(cfg.ses_use_tensor_cores)},
        }
    # ---------------- extra legs
    extras = world == 1 and not args.lean and args.voc_precision == "fp16" and not args.no_tensor_cores

    def leg(name, fn, *a):
        if line is None:
            return
        t0 = time.perf_counter()
        try:
            line[name] = fn(*a)
        except Exception as ex:
            line[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
        line[name + "_leg_seconds"] = round(time.perf_counter() - t0, 1)

    if extras:
        leg("scheduler", leg_scheduler, rig, K)
        leg("session", leg_session, rig, K)
    rig.close()
    if extras:
        leg("config1", leg_config1, local, sds)
        leg("config2", leg_config2, local, sds, K)
        leg("config3", leg_config3, local, sds, K)
        leg("fp32_grade", leg_fp32_grade, local, sds, S, K)
        leg("torch_cuda", leg_torch_cuda, local, S)
        leg("sweep", leg_sweep, local, sds, [2048, 3072, 4096, 8192, 12288, 14336], K)
    # config 5: 8192 streams sharded over the N GPUs (every rank runs its own shard; max over ranks)
    S5 = 8192 // world
    if not args.lean and world > 1 and S5 != S:
        r5 = Rig(S5, local, sds)
        for _ in range(3):
            r5.step()
        sync_all()
        per5, ms5 = r5.timed(K)
        sync_all()
        (ms5,) = allmax(ms5)
        if line is not None:
            line["config5"] = {"total_streams": S5 * world, "streams_per_gpu": S5, "steps": K, "ms_per_step": ms5 / K,
                               "value": world * S5 * K / (ms5 * 1e-3) * CHUNK_S, "latency_ms_rank0": _lat(per5),
                               "realtime": ms5 / K < CHUNK_S * 1e3}
        r5.close()
    if line is not None:
        if world == 1 and not args.no_cpu_baseline:
            cv, cdt, cthreads = cpu_oracle_throughput(16, 60)
            line["cpu_baseline"] = {"value": cv, "unit": "streams", "cores": cthreads, "kind": "port",
                                    "sample": f"16 lock-step streams x 60 chunks of the same workload in {cdt:.1f} s "
                                              "(oracle/incremental.py on all host cores)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _measured_traffic(kernel_name: str):
    """DRAM bytes per launch of a kernel family, measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum) on the same
    command and committed under profiles/ (tools/summarize_profiles.py traffic); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_traffic.json")))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
    except Exception:
        return None
    for fam, v in d.items():
        if not fam.startswith("_") and kernel_name.startswith(fam):
            return v["dram_bytes_per_launch"]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=1024, help="resident streams per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voc-precision", default="fp16", choices=["fp16", "split", "fp32"])
    ap.add_argument("--no-tensor-cores", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="run the 32 / 64 channel residual blocks conv by conv (A/B against the fused kernel)")
    ap.add_argument("--voc-group", type=int, default=0)
    ap.add_argument("--min-seconds", type=float, default=3.0, help="device time of the sustained window the headline value is measured over")
    ap.add_argument("--e2e-seconds", type=float, default=1.5)
    ap.add_argument("--lean", action="store_true", help="headline + e2e + roofline only (no sweep / scheduler / config legs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true", help="open a cudaProfiler window around one step and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.voc_precision == "fp32":
            args.no_tensor_cores = True
        run_b200(args)


if __name__ == "__main__":
    main()
