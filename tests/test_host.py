"""CPU: host-side logic -- config loader, checkpoint layout, weight packing, mel front-end,
chunk scheduler (with a recording fake engine) and the multi-rank stream partition (gloo, world 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conan_b200 import audio, ckpt, hparams as hp_mod, synth
from conan_b200.scheduler import ChunkScheduler, shard_streams
from conan_b200.weights import pack_conv, sinusoid_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ hparams
def test_set_hparams_inheritance_and_overrides(tmp_path, monkeypatch):
    (tmp_path / "base.yaml").write_text("a: 1\nb: {c: 2, d: 3}\nlst: [1, 2]\nflag: false\n")
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "mid.yaml").write_text("base_config: ../base.yaml\na: 5\nname: x\n")
    (tmp_path / "top.yaml").write_text("base_config:\n  - ./sub/mid.yaml\nb: {c: 9}\n")
    monkeypatch.chdir(tmp_path)
    cfg = hp_mod.set_hparams(config="top.yaml", hparams_str="a=7,b.d=4,lst=[3 4 5],flag=True,name=y", print_hparams=False)
    assert cfg["a"] == 7 and cfg["b"] == {"c": 9, "d": 4} and cfg["lst"] == [3, 4, 5] and cfg["flag"] is True and cfg["name"] == "y"
    assert cfg["work_dir"] == "" and hp_mod.hparams["a"] == 7          # global dict is populated
    local = hp_mod.set_hparams(config="base.yaml", global_hparams=False, print_hparams=False)
    assert local["a"] == 1 and hp_mod.hparams["a"] == 7                # global untouched


def test_shipped_configs_carry_the_hot_path_keys(monkeypatch):
    monkeypatch.chdir(ROOT)
    cfg = hp_mod.set_hparams(config="egs/conan_emformer.yaml", print_hparams=False, global_hparams=False)
    for k, v in synth.DEFAULT_HP.items():
        assert cfg[k] == v, k
    voc = hp_mod.set_hparams(config="egs/hifi_16k320_shuffle.yaml", print_hparams=False, global_hparams=False)
    for k in ("upsample_rates", "upsample_kernel_sizes", "upsample_initial_channel", "resblock_kernel_sizes", "resblock_dilation_sizes"):
        assert voc[k] == synth.DEFAULT_VOC_HP[k], k
    from conan_b200.engine import make_config
    c = make_config(cfg, voc, max_slots=4, max_ref_frames=100)
    assert (c.segment, c.right_context, c.hidden_size, c.voc_n_ups, list(c.voc_rates)[:4]) == (4, 2, 256, 4, [8, 5, 4, 2])
    with pytest.raises(ValueError):
        make_config({**cfg, "f0_gen": "flow"}, voc)


# ------------------------------------------------------------------ checkpoints
def test_checkpoint_layout_roundtrip_and_selection(tmp_path):
    sd = {"a.weight": torch.randn(3, 2), "b": torch.randn(4)}
    ckpt.save_checkpoint(sd, str(tmp_path / "m"), "model", steps=10)
    ckpt.save_checkpoint({k: v + 1 for k, v in sd.items()}, str(tmp_path / "m"), "model", steps=200)
    got = ckpt.load_state_dict(str(tmp_path / "m"), "model")
    assert torch.equal(got["b"], sd["b"] + 1)                          # newest step wins
    got = ckpt.load_state_dict(str(tmp_path / "m" / "model_ckpt_steps_10.ckpt"), "model")
    assert torch.equal(got["b"], sd["b"])
    flat = {"state_dict": {"model_gen.a.weight": sd["a.weight"], "model_disc.x": torch.zeros(1)}}
    os.makedirs(tmp_path / "f")
    torch.save(flat, tmp_path / "f" / "model_ckpt_steps_1.ckpt")
    got = ckpt.load_state_dict(str(tmp_path / "f"), "model_gen")
    assert list(got) == ["a.weight"]
    os.makedirs(tmp_path / "empty")
    with pytest.raises(AssertionError):
        ckpt.load_state_dict(str(tmp_path / "empty"), "model")
    assert ckpt.load_state_dict(str(tmp_path / "empty"), "model", force=False) is None


def test_filter_to_spec_follows_load_state_dict_semantics(capsys):
    """utils/commons/ckpt_utils.py:48-58: strict=False drops shape-mismatched keys (printing `| Unmatched keys:`) and lets
    missing keys keep their initial value; strict=True raises like nn.Module.load_state_dict."""
    spec = [("a.weight", (3, 2), "x"), ("b", (4,), "x")]
    init = {"a.weight": torch.full((3, 2), 7.0), "b": torch.full((4,), 8.0)}
    got = ckpt.filter_to_spec({"a.weight": torch.zeros(2, 2), "b": torch.ones(4), "extra": torch.zeros(1)}, spec, strict=False, defaults=init)
    assert "| Unmatched keys:  a.weight (3, 2) (2, 2)" in capsys.readouterr().out
    assert torch.equal(got["a.weight"], init["a.weight"]) and torch.equal(got["b"], torch.ones(4)) and "extra" not in got
    got = ckpt.filter_to_spec({"b": torch.ones(4)}, spec, strict=False, defaults=init)           # missing key -> initial value
    assert torch.equal(got["a.weight"], init["a.weight"])
    with pytest.raises(KeyError):
        ckpt.filter_to_spec({"b": torch.ones(4)}, spec, strict=False)                            # no initial values to keep
    for bad in ({"a.weight": torch.zeros(2, 2), "b": torch.ones(4)}, {"b": torch.ones(4)},
                {"a.weight": torch.zeros(3, 2), "b": torch.ones(4), "extra": torch.zeros(1)}):
        with pytest.raises(RuntimeError):
            ckpt.filter_to_spec(bad, spec, strict=True)


def test_set_hparams_assigns_infer_debug_validate_unconditionally(tmp_path, monkeypatch):
    """utils/commons/hparams.py:113-116: a saved checkpoints/<exp>/config.yaml cannot mask the flags."""
    monkeypatch.chdir(tmp_path)
    os.makedirs("checkpoints/e")
    (tmp_path / "checkpoints" / "e" / "config.yaml").write_text("infer: true\ndebug: true\na: 3\n")
    (tmp_path / "c.yaml").write_text("a: 1\n")
    cfg = hp_mod.set_hparams(config="c.yaml", exp_name="e", print_hparams=False, global_hparams=False)
    assert cfg["a"] == 3 and cfg["infer"] is False and cfg["debug"] is False and cfg["validate"] is False
    monkeypatch.setattr(sys, "argv", ["x", "--config", "c.yaml", "--exp_name", "e", "--infer", "--debug"])
    cfg = hp_mod.set_hparams(print_hparams=False, global_hparams=False)
    assert cfg["infer"] is True and cfg["debug"] is True and cfg["validate"] is False


def test_weight_norm_folding_matches_torch():
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 5, 3))
    with torch.no_grad():
        conv.weight_g.mul_(1.7)
    sd = {"c." + k: v for k, v in conv.state_dict().items()}
    x = torch.randn(2, 6, 9)
    w = ckpt.fold_weight_norm(sd, "c")
    assert torch.allclose(F.conv1d(x, w, sd["c.bias"]), conv(x), atol=1e-6)


# ------------------------------------------------------------------ weight packing
def test_pack_conv_is_tap_major():
    w = torch.randn(5, 4, 3)
    p = pack_conv(w)
    assert p.shape == (5, 12) and torch.equal(p[:, 1 * 4 + 2], w[:, 2, 1])


def test_pixel_shuffle_folds_into_weight_row_order():
    """conv with permuted rows, read as [T*r, C], equals CausalPixelShuffle1d(conv) (hifigan_causal.py:186-188)."""
    C, r, cin, k, T = 6, 4, 8, 3, 5
    w, b, x = torch.randn(C * r, cin, k), torch.randn(C * r), torch.randn(1, cin, T + k - 1)
    y = F.conv1d(x, w, b)                                              # [1, C*r, T], channel c*r + j
    ref = y.view(1, C, r, T).permute(0, 1, 3, 2).reshape(1, C, T * r)  # the reference's shuffle
    wp = w.view(C, r, cin, k).permute(1, 0, 2, 3).reshape(r * C, cin, k)
    bp = b.view(C, r).t().reshape(-1)
    yp = F.conv1d(x, wp, bp).transpose(1, 2).reshape(1, T * r, C)      # channels-last rows [t, r*C] -> [t*r + j, c]
    assert torch.allclose(yp.transpose(1, 2), ref, atol=1e-6)


def test_sinusoid_table_layout():
    t = sinusoid_table(20, 256)
    assert (t[0] == 0).all() and abs(float(t[3, 0]) - np.sin(3.0)) < 1e-6 and abs(float(t[3, 128]) - np.cos(3.0)) < 1e-6


# ------------------------------------------------------------------ mel front-end
def test_slaney_mel_basis_matches_torchaudio():
    import torchaudio
    mine = audio.slaney_mel_basis(16000, 1024, 80, 80, 7600)
    ta = torchaudio.functional.melscale_fbanks(513, 80.0, 7600.0, 80, 16000, norm="slaney", mel_scale="slaney").t().numpy()
    assert np.abs(mine - ta).max() < 1e-6
    wav = np.sin(2 * np.pi * 440 * np.arange(16000) / 16000).astype(np.float32) * 0.5
    mel = audio.wav2mel(wav)
    assert mel.shape == (16000 // 320 + 1, 80) and np.isfinite(mel).all()
    assert 8 <= int(mel[10].argmax()) <= 16                         # 440 Hz lands in the low mel bins


def test_bs1770_loudness_normalisation():
    """`loud_norm` branch of librosa_wav2spec (utils/audio/__init__.py:57-62).  BS.1770 anchor: a 997 Hz full-scale sine reads
    -3.01 LUFS (the K-weighting is ~0 dB there), so amplitude 0.1 reads ~ -23.0."""
    sr = 16000
    x = (0.1 * np.sin(2 * np.pi * 997 * np.arange(3 * sr) / sr)).astype(np.float32)
    assert abs(audio.integrated_loudness(x, sr) + 23.01) < 0.15
    y = audio.loudness_normalize(x, sr)
    assert abs(audio.integrated_loudness(y, sr) + 22.0) < 1e-3 and y.dtype == np.float32
    loud = audio.loudness_normalize(x * 9.5, sr, target_lufs=0.0)           # would exceed full scale: peak-limited
    assert abs(np.abs(loud).max() - 1.0) < 1e-6


# ------------------------------------------------------------------ scheduler (fake engine)
class FakeEngine:
    segment, rows_in, hop_out, n_mels = 4, 6, 1280, 80

    def __init__(self):
        self.calls, self.opened, self.resets = [], [], []

    def reset_slots(self, slots, parts=7):
        self.resets.append(list(slots))

    def open_sessions(self, slots, ref):
        self.opened.append((list(slots), tuple(ref.shape)))

    def step_host(self, slots, chunk, wav, mel, tok):
        self.calls.append((slots.copy(), chunk.copy()))
        for i, s in enumerate(slots):
            wav[i] = chunk[i, :4, 0].repeat(320)          # echo: wav sample block t carries mel[t, 0]
            mel[i] = chunk[i, :4]
            tok[i] = s


def test_scheduler_chunk_assembly_matches_reference_loop_rule():
    from oracle.incremental import assemble_chunk
    eng = FakeEngine()
    sch = ChunkScheduler(eng, 4)
    src = synth.synth_mel(23, 1).numpy()                   # 5 full chunks + a 3-frame tail
    sid = sch.open(np.zeros((30, 80), np.float32))
    sch.push(sid, src)
    sch.end(sid)
    pos, mels = 0, []
    while not sch.finished(sid):
        out = sch.step()[sid]
        chunk_ref, emit = assemble_chunk(torch.from_numpy(src)[None], pos)
        assert np.array_equal(eng.calls[-1][1][0], chunk_ref[0].numpy())
        assert out[1].shape[0] == emit and out[0].shape[0] == emit * 320
        mels.append(out[1])
        pos += emit
    assert pos == 23 and np.array_equal(np.concatenate(mels), src)


def test_scheduler_packs_only_ready_streams_and_recycles_slots():
    eng = FakeEngine()
    sch = ChunkScheduler(eng, 3)
    a, b, c = sch.open_many([np.zeros((20, 80), np.float32), np.zeros((20, 80), np.float32), np.zeros((8, 80), np.float32)])
    assert sorted(len(o[0]) for o in eng.opened) == [1, 2]            # equal-length references share one setup call
    with pytest.raises(RuntimeError):
        sch.open(np.zeros((8, 80), np.float32))                        # pool exhausted
    sch.push(a, synth.synth_mel(6, 1).numpy())                         # exactly one chunk + look-ahead
    sch.push(b, synth.synth_mel(5, 2).numpy())                         # look-ahead not yet complete
    assert sch.ready() == [a]
    out = sch.step()
    assert list(out) == [a] and len(eng.calls[-1][0]) == 1
    sch.push(b, synth.synth_mel(3, 3).numpy())                         # now 8 frames buffered
    sch.push(a, synth.synth_mel(2, 4).numpy())                         # a: 8 frames, pos 4 -> needs 10
    assert sch.ready() == [b]
    sch.end(a)
    assert sorted(sch.ready()) == [a, b]
    out = sch.step()
    assert out[a][1].shape[0] == 4 and out[b][1].shape[0] == 4 and len(eng.calls[-1][0]) == 2
    assert sch.finished(a) and not sch.finished(b)
    slot_a = sch.streams[a].slot
    sch.close(a)
    d = sch.open(np.zeros((8, 80), np.float32))
    assert sch.streams[d].slot == slot_a and eng.resets[-1] == [slot_a]
    sch.end(c)
    assert c not in sch.ready()                                        # ended with nothing buffered


def test_scheduler_host_buffers_stay_bounded_over_a_long_stream():
    """per-stream host memory must not grow with the stream's length (1e5 frames ~ 33 minutes of audio)."""
    eng = FakeEngine()
    sch = ChunkScheduler(eng, 2, capacity_frames=32)
    sid = sch.open(np.zeros((10, 80), np.float32))
    ring_bytes = sch.ring.nbytes
    rng = np.random.default_rng(0)
    fed = emitted = 0
    first = None
    while fed < 100_000:
        f = rng.standard_normal((40, 80)).astype(np.float32)        # 40 frames per push: more than the ring has room for at times
        if first is None:
            first = f.copy()
        sch.push(sid, f)
        fed += 40
        while sch.ready():
            out = sch.step()[sid]
            if emitted == 0:
                assert np.array_equal(out[1], first[:4])
            emitted += out[1].shape[0]
        assert sch.buffered_frames(sid) <= sch.cap and sch.pending_frames(sid) < 40 + sch.rows
    sch.end(sid)
    while not sch.finished(sid):
        emitted += sch.step()[sid][1].shape[0]
    assert emitted == fed and sch.ring.nbytes == ring_bytes and not sch._backlog
    assert len(eng.calls) == fed // 4


def test_scheduler_whole_utterance_push_goes_through_the_backlog():
    eng = FakeEngine()
    sch = ChunkScheduler(eng, 1, capacity_frames=16)
    src = synth.synth_mel(203, 7).numpy()
    sid = sch.open(np.zeros((10, 80), np.float32))
    sch.push(sid, src[:150])
    sch.push(sid, src[150:])                                           # queued behind the backlog, order kept
    sch.end(sid)
    mels = []
    while not sch.finished(sid):
        mels.append(sch.step()[sid][1])
    assert np.array_equal(np.concatenate(mels), src)


class PipelinedFakeEngine(FakeEngine):
    """step_host_submit / _wait with the result written only at wait time (like the real copy stream)."""

    def __init__(self):
        super().__init__()
        self.pending = {}
        self.n_sub = 0

    def step_host_submit(self, slots, chunk, wav, mel, tok):
        t = self.n_sub & 1
        assert t not in self.pending
        self.pending[t] = (slots, chunk, wav, mel, tok)
        self.n_sub += 1
        return t

    def step_host_wait(self, t):
        self.step_host(*self.pending.pop(t))


def test_scheduler_vectorised_lockstep_and_pipelined_stepping():
    eng = PipelinedFakeEngine()
    S = 64
    sch = ChunkScheduler(eng, S)
    sids = sch.open_many([np.zeros((12, 80), np.float32)] * S)
    slots = np.array([sch.streams[s].slot for s in sids])
    src = np.stack([synth.synth_mel(46, 100 + i).numpy() for i in range(S)])            # [S, 46, 80]
    got = {s: [] for s in sids}
    prev, fed = None, 0
    for step in range(11):
        sch.push_many(slots, src[:, fed:fed + (6 if step == 0 else 4)])
        fed += 6 if step == 0 else 4
        t = sch.submit()
        assert t is not None and len(sch._inflight) <= 2
        if prev is not None:
            r = sch.collect(prev)
            assert len(r) == S and (r.emits == 4).all()
            for i, s in enumerate(r.sids):
                got[int(s)].append(r.mel[i].copy())
        prev = t
    r = sch.collect(prev)
    for i, s in enumerate(r.sids):
        got[int(s)].append(r.mel[i].copy())
    for k, s in enumerate(sids):
        assert np.array_equal(np.concatenate(got[s]), src[k, :44])
    with pytest.raises(BufferError):
        for _ in range(40):
            sch.push_many(slots, src[:, :4])                           # nobody consumes: back-pressure


def test_scheduler_warm_up_touches_every_bucket_and_resets():
    for eng in (FakeEngine(), PipelinedFakeEngine()):
        sch = ChunkScheduler(eng, 40)
        sch.warm(max_batch=24, full=True)
        sizes = sorted({len(c[0]) for c in eng.calls})
        assert sizes == [8, 16, 24, 40] and eng.resets[-1] == list(range(40))
        sid = sch.open(np.zeros((8, 80), np.float32))
        with pytest.raises(RuntimeError):
            sch.warm()
        sch.close(sid)


def test_scheduler_failed_admission_releases_slots():
    class Rejecting(FakeEngine):
        def open_sessions(self, slots, ref):
            raise RuntimeError("ref_frames outside [1, max_ref_frames]")
    sch = ChunkScheduler(Rejecting(), 2)
    with pytest.raises(RuntimeError):
        sch.open_many([np.zeros((8, 80), np.float32), np.zeros((9, 80), np.float32)])
    assert len(sch.free) == 2 and not sch.streams and not sch.active.any()
    with pytest.raises(ValueError):
        sch.open(np.zeros((8, 79), np.float32))                        # malformed reference: nothing allocated
    assert len(sch.free) == 2


# ------------------------------------------------------------------ multi-rank partition (gloo)
def _rank_main(rank, world, port, n_streams, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_streams(n_streams, world, rank)
    # no data-path collective: the only exchange is bench-style metadata (units per rank, max time)
    counts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([mine.start, mine.stop], dtype=torch.int64))
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, [c.tolist() for c in counts], float(t)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [8192, 1001])
def test_stream_sharding_world2_gloo(n_streams):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (n_streams % 7)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, n_streams, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    for rank, counts, tmax in res:
        assert tmax == 2.0
        spans = sorted(counts)
        assert spans[0][0] == 0 and spans[-1][1] == n_streams and spans[0][1] == spans[1][0]   # disjoint, contiguous, complete
        assert abs((spans[0][1] - spans[0][0]) - (spans[1][1] - spans[1][0])) <= 1


def test_reference_arm_is_rank0_only(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(type("A", (), {"steps": 1, "warmup": 1, "gpus": 2})())
    assert capsys.readouterr().out == ""


# ------------------------------------------------------------------ serving shell + batch runner (fake engine)
def test_stream_server_admission_queue_backpressure_and_jitter_buffer():
    from conan_b200.serving import StreamServer
    eng = PipelinedFakeEngine()
    srv = StreamServer(eng, 2, max_queue=1, keep_mel=True)
    ref = np.zeros((10, 80), np.float32)
    t0, t1 = srv.admit(ref, tag="a"), srv.admit(ref, tag="b")
    a, b = srv.session_of(t0), srv.session_of(t1)
    assert a is not None and b is not None
    t2 = srv.admit(ref, tag="c")                                       # pool full: queued
    assert srv.session_of(t2) is None and len(srv.queue) == 1
    assert srv.admit(ref) is None and srv.stats["refused"] == 1        # queue full too: refused (back-pressure)
    src = synth.synth_mel(200, 3).numpy()
    took = srv.feed_mel(a, src)                                        # more than the ring holds: only part is accepted
    assert 0 < took < 200 and srv.stats["input_backpressure_events"] == 1
    assert srv.feed_mel(a, src[took:]) == 0                            # still full until the consumer runs
    fed = took
    while fed < 200:
        srv.pump()
        fed += srv.feed_mel(a, src[fed:])
    srv.end(a)
    while not srv.done(a):
        srv.pump()
    assert srv.session_of(t2) is not None                              # the freed slot went to the queued admission
    # output jitter buffer: arbitrary-sized reads reassemble the stream (echo engine: sample block t carries mel[t, 0])
    total = srv.available(a)
    assert total == 200 * 320
    pieces = [srv.read(a, 1000), srv.read(a, 7), srv.read(a)]
    wav = np.concatenate(pieces)
    assert wav.shape[0] == total and np.array_equal(wav[::320], src[:, 0])
    mel = srv.release(a)
    assert np.array_equal(mel, src)


def test_stream_server_pipelined_mode_delivers_everything():
    from conan_b200.serving import StreamServer
    eng = PipelinedFakeEngine()
    srv = StreamServer(eng, 3, pipelined=True, keep_mel=True)
    srcs = [synth.synth_mel(30 + 7 * i, 20 + i).numpy() for i in range(3)]
    sids = [srv.session_of(srv.admit(np.zeros((8, 80), np.float32))) for _ in range(3)]
    for sid, src in zip(sids, srcs):
        assert srv.feed_mel(sid, src) == src.shape[0]
        srv.end(sid)
    for _ in range(40):
        srv.pump()
    srv.flush()
    for sid, src in zip(sids, srcs):
        assert srv.done(sid) and srv.available(sid) == src.shape[0] * 320
        assert np.array_equal(srv.release(sid), src)
    assert srv.stats["closed"] == 3 and not srv.sch.streams


def test_voice_conversion_runner_json_contract(tmp_path, monkeypatch):
    """inference/run_voice_conversion_nvae.py: config schema, output naming, progress / final report keys, error capture."""
    from conan_b200.serving import VoiceConversionRunner
    from conan_b200.scheduler import ChunkScheduler

    class FakeSVC:
        def __init__(self):
            self.engine = FakeEngine()
            self.scheduler = ChunkScheduler(self.engine, 2)

        def _wav_to_mel(self, path):
            if "missing" in path:
                raise FileNotFoundError(path)
            return synth.synth_mel(10 + len(path) % 5, len(path)).numpy()

    pairs = [{"ref_wav": f"r{i}.wav", "src_wav": ("missing.wav" if i == 2 else f"s{i}.wav"), "src_corpus": "vctk",
              "src_utt_id": f"p{i:03d}", "output_name": f"o{i}"} for i in range(5)]
    cfg = tmp_path / "voice_conversion_config.json"
    cfg.write_text(__import__("json").dumps({"total_pairs": 5, "conversion_pairs": pairs}))
    out = tmp_path / "out"
    runner = VoiceConversionRunner(str(cfg), hparams={"audio_sample_rate": 16000}, engine=FakeSVC(), output_dir=str(out))
    rep = runner.run_all_conversions(batch_size=2)
    assert rep["total_processed"] == 5 and rep["successful"] == 4 and rep["failed"] == 1 and rep["success_rate"] == 80.0
    assert sorted(os.listdir(out)) == ["conversion_progress.json", "final_report.json", "vctk_p000.wav", "vctk_p001.wav",
                                       "vctk_p003.wav", "vctk_p004.wav"]
    prog = __import__("json").load(open(out / "conversion_progress.json"))
    assert set(prog) == {"processed", "total", "successful", "failed", "elapsed_time", "estimated_remaining", "current_batch_end", "errors"}
    final = __import__("json").load(open(out / "final_report.json"))
    assert set(final) == {"start_idx", "end_idx", "total_processed", "successful", "failed", "success_rate", "total_time_minutes",
                          "avg_time_per_file_seconds", "output_directory", "errors", "timestamp"}
    assert final["errors"] and "Pair 2" in final["errors"][0]
    with pytest.raises(FileNotFoundError):
        VoiceConversionRunner(str(tmp_path / "nope.json"), hparams={}, engine=FakeSVC(), output_dir=str(out))
