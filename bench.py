#!/usr/bin/env python
"""bench.py -- concurrent real-time 80 ms-chunk voice streams on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference]

A "step" is one pass of the whole chunk path (Emformer step -> proj/argmax -> Conan chunk
decoder -> causal shuffle HiFi-GAN) over one 80 ms chunk of every resident stream: S = 1024
streams per GPU (BASELINE.json configs[3] at N = 1; N x 1024 = configs[4] at N = 8; weak
scaling, streams are independent, there is no collective on the data path).

  value  = (stream-chunks processed per second, all ranks) x 0.08 s
         = concurrent real-time streams the job sustains with inputs already resident in HBM
  e2e    = the same through the plugin calls conan_step_host_submit / _wait (two steps in flight): slot ids +
           mel chunks copied from pinned host memory and wav copied back, every step, inside the timed region;
           the synchronous conan_step_host figure is reported next to it
  roofline: the tcgen05 implicit-GEMM conv kernels of the vocoder (96 % of the path's FLOPs),
           algorithmic FLOPs / CUDA-event time of those launches, vs the measured dense
           16-bit tensor peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU oracle (oracle/incremental.py, the incremental
           PyTorch restatement of the reference's loop; the reference itself is Python and
           /root/reference does not exist on the GPU box) on all host cores.
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
WORKLOAD = "full Conan pipeline (style encoder once + chunk loop), {S} concurrent streams per B200"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1380.4), d.get("hbm_gbs", 6550.7), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], [], set()
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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from conan_b200 import synth
    from conan_b200.engine import Engine, make_config

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
    cfg = make_config(max_slots=S, max_ref_frames=160, device=local, voc_precision=args.voc_precision,
                      voc_tensor_cores=not args.no_tensor_cores, voc_group=args.voc_group,
                      voc_fuse_resblocks=False if args.no_fuse else None)
    eng = Engine(*sds, cfg)
    slots = np.arange(S, dtype=np.int32)
    ids = eng.ids_tensor(slots)
    eng.reset_slots(slots)
    # one session per stream: 8 distinct 3 s reference utterances, repeated (setup, untimed)
    refs = torch.stack([synth.synth_mel(150, 100 + s) for s in range(8)]).to(dev)
    for g in range(0, S, 64):
        n = min(64, S - g)
        eng.open_sessions(slots[g:g + n], refs[torch.arange(g, g + n, device=dev) % 8])
    # synthetic source mel: a pool of chunks larger than one step so every step reads fresh input
    n_pool = 8
    pool = torch.stack([synth.synth_mel(6 * n_pool, 300 + s) for s in range(64)])                 # [64, 6*n_pool, 80]
    chunks_dev = [pool[:, 6 * i:6 * i + 6].repeat(S // 64 + 1, 1, 1)[:S].contiguous().to(dev) for i in range(n_pool)]
    wav = torch.empty(S, eng.hop_out, device=dev)
    mel = torch.empty(S, 4, 80, device=dev)
    tok = torch.empty(S, 4, dtype=torch.int32, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- device-resident timing
    for i in range(W):
        eng.step(ids, chunks_dev[i % n_pool], wav, mel, tok)
    sync_all()
    if args.ncu_step:
        # profiler window for `ncu --profile-from-start off`: exactly one warmed-up step, no bench line
        torch.cuda.profiler.start()
        eng.step(ids, chunks_dev[W % n_pool], wav, mel, tok)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        eng.close()
        return
    sampler = ClockSampler(local)
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    l0 = eng.launch_count
    evs[0].record()
    for i in range(K):
        eng.step(ids, chunks_dev[i % n_pool], wav, mel, tok)
        evs[i + 1].record()
    sync_all()
    launches = eng.launch_count - l0
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = evs[0].elapsed_time(evs[K])
    # ---------------- end-to-end through the plugin call (host buffers)
    h_chunks = [torch.empty(S, 6, 80).pin_memory() for _ in range(n_pool)]
    for h, d in zip(h_chunks, chunks_dev):
        h.copy_(d.cpu())
    h_wavs = [torch.empty(S, eng.hop_out).pin_memory() for _ in range(2)]
    np_chunks = [h.numpy() for h in h_chunks]
    np_wavs = [h.numpy() for h in h_wavs]
    Ke = max(3, min(K, 20))
    for i in range(2):
        eng.step_host(slots, np_chunks[i % n_pool], np_wavs[0])
    sync_all()
    # (a) synchronous plugin call, one step at a time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        eng.step_host(slots, np_chunks[i % n_pool], np_wavs[0])
    e1.record()
    sync_all()
    e2e_sync_ms = e0.elapsed_time(e1)     # device timeline: includes the H2D/D2H copies and every host gap between steps
    # (b) the serving loop: submit step i, then collect step i-1 (its wav copy overlaps step i's compute).  Every step still
    # copies its chunks in from pinned memory and its wav out; e1 is recorded after the last result has landed on the host.
    assert n_pool >= 2
    e2e_api = "conan_step_host_submit / conan_step_host_wait (two steps in flight: result copy of step i under the compute of step i+1)"
    try:
        for i in range(2):                       # untimed: first use of the engine's copy stream and events
            eng.step_host_wait(eng.step_host_submit(slots, np_chunks[i % 2], np_wavs[i % 2]))
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prev = None
        for i in range(Ke):
            t = eng.step_host_submit(slots, np_chunks[i % 2], np_wavs[i % 2])
            if prev is not None:
                eng.step_host_wait(prev)
            prev = t
        eng.step_host_wait(prev)
        e1.record()
        sync_all()
        e2e_ms = e0.elapsed_time(e1)
    except Exception as ex:                      # never lose the bench line: report the synchronous call instead
        print(f"pipelined host stepping failed ({ex}); e2e falls back to conan_step_host", file=sys.stderr)
        e2e_ms, e2e_api = e2e_sync_ms, "conan_step_host"
    clocks = sampler.stop()
    # ---------------- roofline of the dominant kernel family (separate, untimed-by-the-headline pass)
    eng.set_profiling(True)
    NPROF = 2
    for i in range(NPROF):
        eng.step(ids, chunks_dev[i % n_pool], wav, mel, tok)
    prof = {cat: eng.profile_read(cat) for cat in range(6)}
    eng.set_profiling(False)

    t = torch.tensor([total_ms, e2e_ms, e2e_sync_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_sync_ms = t.tolist()
    if rank == 0:
        tf_peak, hbm_peak, peak_src = _peaks()
        value = world * S * K / (total_ms * 1e-3) * CHUNK_S
        e2e_val = world * S * Ke / (e2e_ms * 1e-3) * CHUNK_S
        ps = sorted(per_step)
        names = {0: ("conv_gemm_ffma_kernel (fp32 CUDA-core engine)", "tensor"),
                 1: ("conv_gemm_tc_kernel (tcgen05 implicit-GEMM causal conv, fp16 operands: vocoder scales 0-1 + upsampling)", "tensor"),
                 2: ("conv_window_tc_kernel (tcgen05, persistent, weights resident, one input window per tile: vocoder scales 2-3)", "hbm"),
                 3: ("conv_gemm_tc_kernel, split-fp16 operands (Emformer / Conan linear + conv contractions, 3 MMAs per product)", "tensor"),
                 4: ("resblock_fused_kernel (tcgen05, one HiFi-GAN residual block = six convs per launch, activations in shared memory: vocoder scales 2-3)", "tensor"),
                 5: ("ffn_fused_kernel (tcgen05, Emformer 80 -> 2048 -> 80 feed-forward in one launch, split-fp16 operands, hidden activation in shared memory)", "tensor")}
        roofs = {}
        for cat, (ms, nl, fl, by) in prof.items():
            if nl == 0:
                continue
            kname, bound = names[cat]
            tf, gb = fl / (ms * 1e-3) / 1e12, by / (ms * 1e-3) / 1e9
            r = {"kernel": kname, "bound": bound, "launches_per_step": int(nl // NPROF), "ms_per_step": ms / NPROF,
                 "algorithmic_gflop_per_step": fl / NPROF / 1e9, "algorithmic_gbyte_per_step": by / NPROF / 1e9,
                 "tflops": tf, "gbs": gb, "traffic": _measured_traffic(kname), "peak_source": peak_src}
            if bound == "tensor":
                r.update(achieved=tf, peak=tf_peak, unit="TFLOP/s", frac=tf / tf_peak)
            else:
                r.update(achieved=gb, peak=hbm_peak, unit="GB/s", frac=gb / hbm_peak)
            roofs[cat] = r
        top = max(roofs, key=lambda c_: roofs[c_]["ms_per_step"])
        roof = dict(roofs[top])
        roof["other_kernels"] = [roofs[c_] for c_ in sorted(roofs) if c_ != top]
        line = {
            "metric": "concurrent real-time 80 ms-chunk streams", "value": value, "unit": "streams", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "split-f16 operands, f32 accumulate = f32-grade (Emformer/Conan) + f16 operands / f32 accumulate (vocoder)"
            if args.voc_precision == "fp16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(S=S), "streams_per_gpu": S, "chunk_ms": 80, "ref_frames": 150,
                       "weights": "synthetic seeded (conan_b200.synth, reference state_dict layout)",
                       "l2": f"per-step working set (resident state {eng.state_bytes / 2**30:.1f} GiB) is larger than L2; no flush needed",
                       "voc_precision": args.voc_precision, "voc_tensor_cores": not args.no_tensor_cores, "voc_group": args.voc_group,
                       "voc_fuse_resblocks": bool(cfg.voc_fuse_resblocks)},
            "latency_ms": {"p50": ps[len(ps) // 2], "p99": ps[min(len(ps) - 1, int(len(ps) * 0.99))], "max": ps[-1]},
            "rtf": (total_ms / K) / (CHUNK_S * 1e3),
            "path_tflops": world * S * K * FLOP_PER_STREAM_CHUNK / (total_ms * 1e-3) / 1e12,
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "streams", "h2d_bytes_per_step": int(S * (6 * 80 * 4 + 4)),
                    "d2h_bytes_per_step": int(S * eng.hop_out * 4), "steps": Ke, "ms_per_step": e2e_ms / Ke,
                    "api": e2e_api,
                    "synchronous_call": {"value": world * S * Ke / (e2e_sync_ms * 1e-3) * CHUNK_S, "ms_per_step": e2e_sync_ms / Ke,
                                         "api": "conan_step_host"}},
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            cv, cdt, cthreads = cpu_oracle_throughput(16, 120)
            line["cpu_baseline"] = {"value": cv, "unit": "streams", "cores": cthreads, "kind": "port",
                                    "sample": f"16 lock-step streams x 120 chunks of the same workload in {cdt:.1f} s "
                                              "(oracle/incremental.py on all host cores)"}
        print(json.dumps(line), flush=True)
    eng.close()
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
    ap.add_argument("--voc-precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-tensor-cores", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="run the 32 / 64 channel residual blocks conv by conv (A/B against the fused kernel)")
    ap.add_argument("--voc-group", type=int, default=0)
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
