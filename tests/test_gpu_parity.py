"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the golden vectors of the unmodified reference, and through
size-independent properties.  Run on the B200 box:  pytest tests -m gpu

Tolerances (BASELINE.json north_star): mel max-abs <= 1e-3 with fp32 operands, waveform
SNR >= 40 dB vs the reference; tokens (argmax) must agree exactly on these seeds.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conan_b200 import synth
from util import snr_ac_db, snr_db

pytestmark = pytest.mark.gpu

MEL_TOL = 1e-3
SNR_MIN_DB = 40.0


NO_TC = bool(os.environ.get("CONAN_TEST_NO_TC"))      # first-bring-up switch: run everything on the FFMA engine only


def _engine(state_dicts, **kw):
    from conan_b200.engine import Engine, make_config
    if NO_TC:
        kw["voc_tensor_cores"] = False
    kw.setdefault("max_slots", 8)
    kw.setdefault("max_ref_frames", 256)
    cfg = make_config(**kw)
    return Engine(*state_dicts, cfg)


@pytest.fixture(scope="module")
def eng_fp32(state_dicts):
    e = _engine(state_dicts, voc_precision="fp32", voc_tensor_cores=False, lin_tensor_cores=False)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_fp16_ffma(state_dicts):
    e = _engine(state_dicts, voc_precision="fp16", voc_tensor_cores=False)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_split(state_dicts):
    """fp32-grade tensor-core engine: split-fp16 operands everywhere (vocoder included)."""
    e = _engine(state_dicts, voc_precision="split", voc_tensor_cores=True)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_tc(state_dicts):
    e = _engine(state_dicts, voc_precision="fp16", voc_tensor_cores=True)
    yield e
    e.close()


# ------------------------------------------------------------------------------------------
# operator level
# ------------------------------------------------------------------------------------------
def _conv_reference(ctx, w, bias, k, dil, L, row0):
    """ctx [S, rows, cin] float64, w [cout, cin, k] float64 -> [S, L, cout]"""
    x = ctx[:, row0:row0 + L + (k - 1) * dil].transpose(1, 2)
    return F.conv1d(x, w, bias, dilation=dil).transpose(1, 2)


@pytest.mark.parametrize("cin,cout,k,dil,L,S", [
    (80, 240, 1, 1, 6, 5), (80, 2048, 1, 1, 6, 3), (2048, 80, 1, 1, 6, 3), (256, 512, 5, 1, 4, 7),
    (256, 256, 3, 1, 4, 2), (80, 100, 1, 1, 4, 9), (256, 512, 31, 1, 50, 2), (32, 32, 11, 5, 1280, 2),
    (128, 2, 1, 1, 4, 3),
])
def test_conv_gemm_ffma_fp32_vs_torch(cin, cout, k, dil, L, S):
    from conan_b200 import ops
    from conan_b200.weights import pack_conv
    g = torch.Generator().manual_seed(cin * 7 + cout + k)
    H = (k - 1) * dil
    ctx = torch.randn(S, H + L, cin, generator=g)
    w = torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(S, L, cout, generator=g)
    ref = _conv_reference(ctx.double(), w.double(), b.double(), k, dil, L, 0)
    ref = F.leaky_relu(ref * 0.5, 0.1) + res.double()
    y = torch.zeros(S, L, cout, device="cuda")
    ops.conv_gemm(ctx.cuda(), pack_conv(w).cuda(), b.cuda(), k=k, dil=dil, L=L, row0=0, scale=0.5, act="lrelu", slope=0.1,
                  res=res.cuda(), y=y)
    err = (y.cpu().double() - ref).abs().max().item()
    assert err < 2e-5, err


def test_conv_gemm_slot_indirection_and_outputs():
    from conan_b200 import ops
    from conan_b200.weights import pack_conv
    g = torch.Generator().manual_seed(3)
    S, cin, cout, k, dil, L = 6, 64, 64, 3, 3, 32
    H = (k - 1) * dil
    ctx = torch.randn(S, H + L, cin, generator=g)
    w = torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ids = torch.tensor([4, 1, 5], dtype=torch.int32)
    y = torch.full((S, L, cout), 7.0, device="cuda")
    y2 = torch.zeros(S, 10 + L, cout, device="cuda")
    mask = (torch.rand(S, L, generator=g) > 0.3).float()
    ops.conv_gemm(ctx.cuda(), pack_conv(w).cuda(), b.cuda(), k=k, dil=dil, L=L, row0=0, slot_ids=ids.cuda(), y=y, accumulate=True,
                  out_scale=1 / 3, rowmask=mask.cuda(), y2=y2, y2_row0=10, act2="lrelu", slope2=0.1)
    ref = _conv_reference(ctx.double(), w.double(), b.double(), k, dil, L, 0) * mask[:, :, None].double() / 3 + 7.0
    yc = y.cpu().double()
    for s in range(S):
        if s in (4, 1, 5):
            assert (yc[s] - ref[s]).abs().max() < 2e-5
            assert (y2.cpu().double()[s, 10:] - F.leaky_relu(ref[s], 0.1)).abs().max() < 2e-5
            assert (y2.cpu()[s, :10] == 0).all()
        else:
            assert (yc[s] == 7.0).all() and (y2.cpu()[s] == 0).all()          # untouched slots


TC_SHAPES = [
    # cin, cout, k, dil, L, S      (the vocoder's layers: MRF convs per scale, upsampling convs)
    (256, 256, 3, 1, 32, 5), (256, 256, 11, 5, 32, 4), (256, 256, 7, 3, 32, 9),
    (128, 128, 7, 3, 160, 3), (128, 128, 11, 1, 160, 2),
    (64, 64, 11, 5, 640, 2), (64, 64, 3, 1, 640, 1),
    (32, 32, 3, 1, 1280, 2), (32, 32, 11, 5, 1280, 1), (32, 32, 7, 3, 1280, 3),
    (256, 640, 10, 1, 32, 6), (128, 256, 8, 1, 160, 2), (64, 64, 4, 1, 640, 2), (512, 2048, 16, 1, 4, 40),
    # enough 256-row tiles for the CTA-pair kernel (tcgen05.mma.cta_group::2): BN = 256 and BN = 128, odd and even tile counts,
    # a stream count that leaves the last tile partly empty
    (256, 256, 3, 1, 32, 601), (256, 256, 7, 3, 32, 596), (128, 128, 11, 1, 160, 121), (512, 2048, 16, 1, 4, 2100),
    (256, 640, 10, 1, 32, 700),
]


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
@pytest.mark.parametrize("cin,cout,k,dil,L,S", TC_SHAPES)
def test_conv_gemm_tcgen05_vs_ffma_and_torch(cin, cout, k, dil, L, S):
    from conan_b200 import ops
    from conan_b200.weights import pack_conv
    g = torch.Generator().manual_seed(cin + 3 * cout + 5 * k + dil)
    H = (k - 1) * dil + 2                       # two spare history rows: row0 = 2 exercises the row offset
    nslots = S + 2
    ctx = (torch.randn(nslots, H + L, cin, generator=g) * 0.5).half()
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).half()
    b = torch.randn(cout, generator=g)
    res = torch.randn(nslots, L, cout, generator=g)
    ids = torch.arange(S, dtype=torch.int32)        # the tcgen05 engine works on compact operands: streams 0..S-1 of S+2 rows
    ref = _conv_reference(ctx.double(), w.double(), b.double(), k, dil, L, 2) + res.double()
    outs = {}
    for name, engine in (("ffma", ops.ENGINE_FFMA), ("tc", ops.ENGINE_TC)):
        y = torch.zeros(nslots, L, cout, device="cuda")
        y2 = torch.zeros(nslots, 4 + L, cout, device="cuda", dtype=torch.float16)
        ops.conv_gemm(ctx.cuda(), pack_conv(w).cuda(), b.cuda(), k=k, dil=dil, L=L, row0=2, n_streams=S, engine=engine,
                      res=res.cuda(), y=y, y2=y2, y2_row0=4, act2="lrelu", slope2=0.1)
        torch.cuda.synchronize()
        outs[name] = (y.cpu(), y2.cpu())
    sel = ids.long()
    for name in ("ffma", "tc"):
        err = (outs[name][0][sel].double() - ref[sel]).abs().max().item()
        assert err < 1e-4, (name, err)
    assert (outs["tc"][0] - outs["ffma"][0]).abs().max().item() < 1e-4
    assert (outs["tc"][1].float() - outs["ffma"][1].float()).abs().max().item() < 4e-3      # one fp16 ulp at |v| < 4
    untouched = [s for s in range(nslots) if s not in set(sel.tolist())]
    assert (outs["tc"][0][untouched] == 0).all()


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
@pytest.mark.parametrize("cin,cout,k,dil,L,S", [(64, 64, 7, 3, 640, 2), (128, 128, 3, 1, 160, 5), (32, 32, 11, 5, 1280, 1)])
def test_conv_gemm_residual_from_activated_fp16_rows(cin, cout, k, dil, L, S):
    """res given as fp16 lrelu(x) rows with the inverse slope: both engines must add x back (x = h >= 0 ? h : 10 h)."""
    from conan_b200 import ops
    from conan_b200.weights import pack_conv
    g = torch.Generator().manual_seed(cin + k)
    H = (k - 1) * dil
    ctx = (torch.randn(S, H + L, cin, generator=g) * 0.5).half()
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).half()
    b = torch.randn(cout, generator=g)
    xres = torch.randn(S, L, cout, generator=g)
    act = F.leaky_relu(xres, 0.1).half()
    x_back = torch.where(act.float() >= 0, act.float(), act.float() * 10.0)
    ref = _conv_reference(ctx.double(), w.double(), b.double(), k, dil, L, 0) + x_back.double()
    outs = []
    for engine in (ops.ENGINE_FFMA, ops.ENGINE_TC):
        y = torch.zeros(S, L, cout, device="cuda")
        ops.conv_gemm(ctx.cuda(), pack_conv(w).cuda(), b.cuda(), k=k, dil=dil, L=L, row0=0, engine=engine, res=act.cuda(),
                      res_inv_slope=10.0, y=y)
        outs.append(y.cpu())
        assert (outs[-1].double() - ref).abs().max().item() < 1e-4
    assert (outs[0] - outs[1]).abs().max().item() < 1e-4


SPLIT_SHAPES = [
    # cin, cout, k, L, S     (Emformer / Conan contractions in fp32-grade tensor-core mode; dims already padded to 32)
    (96, 256, 1, 6, 70), (96, 2048, 1, 6, 33), (2048, 96, 1, 6, 20), (96, 128, 1, 4, 50),
    (256, 256, 3, 4, 40), (256, 512, 5, 4, 35), (512, 256, 1, 4, 64), (256, 2048, 1, 4, 10), (128, 128, 5, 4, 33),
]


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
@pytest.mark.parametrize("cin,cout,k,L,S", SPLIT_SHAPES)
def test_conv_gemm_tcgen05_split_fp16_is_fp32_grade(cin, cout, k, L, S):
    """x_hi*W_hi + x_hi*W_lo + x_lo*W_hi on the tensor cores vs a float64 reference.  The operands carry ~22 bits;
    what remains is the tensor core's fp32 accumulation (truncating adds: the error grows with the number of
    K-steps, measured 3e-6 at K=96 .. 5e-5 at K=2048 for O(1) outputs) -- two orders below the 1e-3 mel budget
    and below the smallest argmax margin of the golden runs (8e-4)."""
    from conan_b200 import ops
    from conan_b200.weights import pack_conv, split_linear
    g = torch.Generator().manual_seed(cin + cout + k + L)
    H = (k - 1) + 1
    nslots = S + 3
    x = torch.randn(nslots, H + L, cin, generator=g)
    w = torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(nslots, L, cout, generator=g)
    xh = x.half()
    ctx = torch.stack([xh, (x - xh.float()).half()]).contiguous()
    w3, bp = split_linear(pack_conv(w), b, k)
    ref = F.gelu(_conv_reference(x.double(), w.double(), b.double(), k, 1, L, 1) * 0.7) + res.double()
    y = torch.zeros(nslots, L, cout, device="cuda")
    y2 = torch.zeros(2, nslots, L + 2, cout, device="cuda", dtype=torch.float16)
    ops.conv_gemm(ctx.cuda(), w3.cuda(), bp.cuda(), k=k, dil=1, L=L, row0=1, n_streams=S, engine=ops.ENGINE_TC, scale=0.7, act="gelu",
                  res=res.cuda(), y=y, y2=y2, y2_row0=2, y2_split=True, x_split=True, acc_scale=1.0 / 1024)
    torch.cuda.synchronize()
    err = (y.cpu()[:S].double() - ref[:S]).abs().max().item()
    y2c = y2.cpu().float()
    pair = (y2c[0] + y2c[1])[:S, 2:]
    err2 = (pair.double() - ref[:S]).abs().max().item()
    print("split-fp16 GEMM max-abs", err, "pair", err2)
    assert err < 1e-4 and err2 < 1e-4
    assert (y.cpu()[S:] == 0).all() and (y2c[:, S:] == 0).all()


# ------------------------------------------------------------------------------------------
# Emformer
# ------------------------------------------------------------------------------------------
def _chunks(src, pos):
    from oracle.incremental import assemble_chunk
    c, emit = assemble_chunk(src, pos)
    return c.contiguous(), emit


@pytest.mark.parametrize("which", ["eng_fp32", "eng_tc"])
def test_emformer_step_vs_oracle_lockstep(state_dicts, which, request):
    from oracle.incremental import EmformerOracle
    eng = request.getfixturevalue(which)
    B, T = 3, 72                                   # 18 chunks: left context saturates at 50 and the 56-row ring wraps
    src = torch.stack([synth.synth_mel(T, 40 + s) for s in range(B)])
    o = EmformerOracle(state_dicts[1])
    o.reset(B)
    slots = [5, 0, 3]
    eng.reset_slots(slots)
    ids = eng.ids_tensor(slots)
    worst, worst_logit = 0.0, 0.0
    for pos in range(0, T, 4):
        chunk, _ = _chunks(src, pos)
        with torch.no_grad():
            enc_ref = o.step(chunk)
            logit_ref = o.logits(enc_ref)
        tok, enc, logits = eng.emformer_step(ids, chunk.cuda(), want_enc=True, want_logits=True)
        worst = max(worst, (enc.cpu() - enc_ref).abs().max().item())
        worst_logit = max(worst_logit, (logits.cpu() - logit_ref).abs().max().item())
        assert (tok.cpu().long() == logit_ref.argmax(-1)).all(), f"token mismatch at pos {pos}"
    print(which, "emformer enc max-abs", worst, "logits max-abs", worst_logit)
    assert worst < 1e-4 and worst_logit < 1e-4
    assert int(eng.debug_read("emformer_past_len", 5).view(torch.int32)[0]) == T


def test_emformer_staggered_streams_match_independent_runs(state_dicts, eng_fp32):
    """Streams of different ages packed into one launch (per-stream past_len) must equal
    independent B=1 runs -- the case torchaudio's batched infer cannot express (TA:392)."""
    from oracle.incremental import EmformerOracle
    eng = eng_fp32
    T = 40
    srcs = [synth.synth_mel(T, 70 + s)[None] for s in range(3)]
    starts = [0, 2, 5]                              # stream s joins at global step starts[s]
    oracles = [EmformerOracle(state_dicts[1]) for _ in range(3)]
    for o in oracles:
        o.reset(1)
    slots = [2, 6, 1]
    eng.reset_slots(slots)
    for step in range(12):
        active = [s for s in range(3) if step >= starts[s] and (step - starts[s]) * 4 < T]
        if not active:
            continue
        chunks, refs = [], []
        for s in active:
            c, _ = _chunks(srcs[s], (step - starts[s]) * 4)
            chunks.append(c)
            with torch.no_grad():
                refs.append(oracles[s].step(c))
        ids = eng.ids_tensor([slots[s] for s in active])
        _, enc, _ = eng.emformer_step(ids, torch.cat(chunks).cuda(), want_enc=True)
        err = (enc.cpu() - torch.cat(refs)).abs().max().item()
        assert err < 1e-4, (step, err)


@pytest.mark.parametrize("M,tc", [(4, True), (4, False), (2, True), (0, True)])
def test_emformer_memory_bank_and_full_utterance_forward_vs_torchaudio(state_dicts, M, tc):
    """SURVEY 8f row f4 / north_star (1) "memory-bank update kernels": the generic Emformer step (memory bank of M vectors per
    layer and stream, summary query masked from the memory keys, clamped memory output, partial segments) against
    torchaudio.models.Emformer(max_memory_size=M) itself: streaming `infer` over 18 chunks (K/V ring and bank both wrap), streams of
    different ages in one launch, and the full-utterance `forward` with its block attention mask (whole and partial last segment)."""
    import torchaudio
    from oracle.incremental import EmformerOracle
    sd = state_dicts[1]
    ta = torchaudio.models.Emformer(80, 8, 2048, 6, 4, left_context_length=50, right_context_length=2, max_memory_size=M).eval()
    ta.load_state_dict({k[len("emformer."):]: v for k, v in sd.items() if k.startswith("emformer.")})
    eng = _engine(state_dicts, emformer_memory_size=M, lin_tensor_cores=tc, voc_precision="fp16" if tc else "fp32", voc_tensor_cores=tc)
    try:
        assert eng.cfg.emformer_memory_size == M
        B, T = 3, 74
        x = torch.stack([synth.synth_mel(T + 2, 60 + b) for b in range(B)])
        slots = [5, 1, 6]
        eng.reset_slots(slots)
        ids = eng.ids_tensor(slots)
        worst, st = 0.0, None
        with torch.no_grad():
            for pos in range(0, 72, 4):
                ref, _, st = ta.infer(x[:, pos:pos + 6], torch.full((B,), 6), st)
                _, enc, _ = eng.emformer_step(ids, x[:, pos:pos + 6].contiguous().cuda(), want_enc=True)
                worst = max(worst, (enc.cpu() - ref).abs().max().item())
        print(f"M={M} tc={tc}: streaming infer max-abs vs torchaudio {worst:.2e}")
        assert worst < 1e-4
        # streams of different ages in one launch: stream b joins at step 2*b (torchaudio runs them one by one)
        eng.reset_slots(slots)
        states = [None] * B
        worst = 0.0
        with torch.no_grad():
            for step in range(10):
                act = [b for b in range(B) if step >= 2 * b]
                chunks, refs = [], []
                for b in act:
                    pos = (step - 2 * b) * 4
                    ch = x[b:b + 1, pos:pos + 6]
                    r, _, states[b] = ta.infer(ch, torch.full((1,), 6), states[b])
                    chunks.append(ch), refs.append(r)
                _, enc, _ = eng.emformer_step(eng.ids_tensor([slots[b] for b in act]), torch.cat(chunks).contiguous().cuda(), want_enc=True)
                worst = max(worst, (enc.cpu() - torch.cat(refs)).abs().max().item())
        assert worst < 1e-4, worst
        # full-utterance forward (block attention mask in the reference) incl. partial last segments of 3 and 1 frames
        o = EmformerOracle(sd, max_memory_size=M)
        for frames in (66, 65, 63):
            with torch.no_grad():
                ref, _ = ta(x[:, :frames], torch.full((B,), frames - 2))
                ref_logits = o.logits(ref)
            enc, logits, tok = eng.emformer_forward(slots, x[:, :frames], want_logits=True, want_tokens=True)
            err = (enc.cpu() - ref).abs().max().item()
            print(f"M={M} tc={tc}: forward({frames} frames) max-abs vs torchaudio {err:.2e}")
            assert enc.shape == ref.shape and err < 1e-4
            assert (logits.cpu() - ref_logits).abs().max().item() < 2e-4
            assert (tok.cpu().long() == ref_logits.argmax(-1)).all()
    finally:
        eng.close()


def test_emformer_view_forward_matches_streaming_inference(state_dicts, eng_tc):
    """EmformerDistillModel.forward (modules/Emformer/emformer.py:31-47) through the drop-in view == the streamed `inference`
    of the same module wherever real look-ahead exists (the last chunk pads by repeating a frame instead, :74-80)."""
    from conan_b200.streaming import EmformerView
    view = EmformerView(eng_tc, [0, 1])
    x = torch.stack([synth.synth_mel(42, 33 + b) for b in range(2)])
    out, lengths = view(x, torch.full((2,), 42))
    assert out.shape == (2, 40, 100) and lengths.tolist() == [42, 42]
    st, toks = None, []
    for pos in range(0, 36, 4):
        enc, _, st = view.emformer.infer(x[:, pos:pos + 6], torch.full((2,), 6), st)
        toks.append(view.proj(enc).argmax(-1).cpu())
    assert torch.equal(torch.cat(toks, 1), out[:, :36].argmax(-1).cpu())


def test_emformer_encoder_bits_do_not_depend_on_the_batch(eng_tc):
    """The split of the fused feed-forward's hidden dimension is a constant, so a stream's encoder rows (not only its argmax
    tokens) are bit-identical whether it shares the launch with 70 other streams or runs alone."""
    eng = eng_tc
    S = 8
    src = torch.stack([synth.synth_mel(14, 830 + s) for s in range(S)])
    eng.reset_slots(list(range(S)))
    ids = eng.ids_tensor(list(range(S)))
    together = []
    for pos in (0, 4, 8):
        _, enc, _ = eng.emformer_step(ids, src[:, pos:pos + 6].contiguous().cuda(), want_enc=True)
        together.append(enc.cpu())
    eng.reset_slots([5])
    one = eng.ids_tensor([5])
    for k, pos in enumerate((0, 4, 8)):
        _, enc, _ = eng.emformer_step(one, src[2:3, pos:pos + 6].contiguous().cuda(), want_enc=True)
        assert torch.equal(enc.cpu()[0], together[k][2])


# ------------------------------------------------------------------------------------------
# Conan main model
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["eng_fp32", "eng_tc"])
@pytest.mark.parametrize("t_ref", [150, 64, 9])
def test_session_open_vs_oracle(state_dicts, which, request, t_ref):
    """eng_fp32: the whole setup on the fp32 FFMA engine; eng_tc: the style encoder's ConvBlocks on tcgen05 with split-fp16
    operands over sessions padded to a multiple of 32 rows (150 -> 160, 9 -> 32; 64 is exact)."""
    from oracle.incremental import ConanOracle
    eng = request.getfixturevalue(which)
    assert bool(eng.cfg.ses_use_tensor_cores) == (which == "eng_tc")
    B = 2
    ref = torch.stack([synth.synth_mel(t_ref, 11 + s) for s in range(B)])
    o = ConanOracle(state_dicts[0])
    with torch.no_grad():
        dbg = o.open(ref, return_debug=True)
    slots = [3, 6]
    eng.open_sessions(slots, ref.cuda())
    Tp = (t_ref - 1) // 4 + 1
    tp_max = (eng.cfg.max_ref_frames - 1) // 4 + 1
    for i, s in enumerate(slots):
        style = eng.debug_read("style", s).cpu()
        assert (style - o.style[i]).abs().max().item() < 1e-4
        kv = eng.debug_read("kv_cache", s).cpu().view(2, tp_max, 2, 256)
        for l in range(2):
            assert (kv[l, :Tp, 0] - o.K[l][i]).abs().max().item() < 2e-4
            assert (kv[l, :Tp, 1] - o.V[l][i]).abs().max().item() < 2e-4
        kpm = eng.debug_read("kpm", s).cpu()
        assert (kpm[:Tp] == o.kpm[i].float()).all() and (kpm[Tp:] == 1).all()
    idx = eng.debug_read("vq_index", 0).view(torch.int32).cpu()[:Tp]
    assert (idx.long() == dbg["vq_idx"][0]).all()


@pytest.mark.parametrize("which", ["eng_fp32", "eng_tc"])
def test_decoder_step_vs_oracle_teacher_forced(state_dicts, which, request):
    from oracle.incremental import ConanOracle
    eng = request.getfixturevalue(which)
    B, n_chunks = 3, 14                              # 56 frames: longer than the 56-frame receptive field of the chunk path
    ref = torch.stack([synth.synth_mel(90, 21 + s) for s in range(B)])
    g = torch.Generator().manual_seed(9)
    tokens = torch.randint(0, 100, (B, n_chunks * 4), generator=g)
    tokens[0, 5:9] = 57                              # silent_token forces unvoiced (Conan.py:335-338)
    o = ConanOracle(state_dicts[0])
    slots = [1, 4, 2]
    eng.reset_slots(slots)
    eng.open_sessions(slots, ref.cuda())
    ids = eng.ids_tensor(slots)
    worst = 0.0
    buckets = set()
    with torch.no_grad():
        o.open(ref)
        for c in range(n_chunks):
            tk = tokens[:, c * 4:(c + 1) * 4]
            mel_ref, dbg = o.step(tk, return_debug=True)
            mel = eng.decoder_step(ids, tk.to(torch.int32).contiguous().cuda())
            worst = max(worst, (mel.cpu() - mel_ref).abs().max().item())
            for i, s in enumerate(slots):
                uvp = eng.debug_read("uv_pred", i).cpu().view(4, 4)      # scratch is compact: index in the ready list
                assert (uvp[:, 3].long() == dbg["pitch"][i]).all(), "f0 bucket mismatch"
                buckets.update(dbg["pitch"][i].tolist())
    print(which, "decoder mel max-abs", worst, "distinct f0 buckets", len(buckets))
    assert worst < MEL_TOL
    assert len(buckets) > 8                          # the bucket arithmetic is genuinely exercised


def test_decoder_long_reference(state_dicts):
    """A 600-frame reference (150 prosody tokens): more keys than the shared-memory-staged aligner attention
    holds, so the general warp-per-(stream, head) kernel runs; also ragged reference lengths in one batch."""
    from oracle.incremental import ConanOracle
    eng = _engine(state_dicts, max_slots=4, max_ref_frames=640)
    try:
        lens = [600, 333]
        g = torch.Generator().manual_seed(4)
        tokens = torch.randint(0, 100, (2, 16), generator=g)
        worst = 0.0
        with torch.no_grad():
            for b, n in enumerate(lens):                 # the oracle takes one reference length per call
                ref = synth.synth_mel(n, 70 + b)[None]
                o = ConanOracle(state_dicts[0])
                o.open(ref)
                eng.reset_slots([b + 1])
                eng.open_sessions([b + 1], ref.cuda())
                ids = eng.ids_tensor([b + 1])
                for c in range(4):
                    tk = tokens[b:b + 1, c * 4:(c + 1) * 4]
                    mel = eng.decoder_step(ids, tk.to(torch.int32).contiguous().cuda())
                    worst = max(worst, (mel.cpu() - o.step(tk)).abs().max().item())
        print("long-reference decoder mel max-abs", worst)
        assert worst < MEL_TOL
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------
# vocoder
# ------------------------------------------------------------------------------------------
def _run_vocoder(eng, mel, slots):
    eng.reset_slots(slots)
    ids = eng.ids_tensor(slots)
    outs = []
    for i in range(0, mel.shape[1], 4):
        outs.append(eng.vocoder_step(ids, mel[:, i:i + 4].contiguous().cuda()).cpu())
    return torch.cat(outs, 1)


def test_vocoder_fp32_vs_oracle_and_golden(state_dicts, eng_fp32, golden_dir):
    from oracle.incremental import HifiGanOracle
    d = np.load(os.path.join(golden_dir, "vocoder_24f.npz"))
    mel = torch.from_numpy(d["mel"])[None].repeat(2, 1, 1)
    mel[1] = mel[1].flip(0)
    o = HifiGanOracle(state_dicts[2])
    o.reset(2)
    with torch.no_grad():
        ref = torch.cat([o.step(mel[:, i:i + 4]) for i in range(0, 24, 4)], 1)
    wav = _run_vocoder(eng_fp32, mel, [7, 2])
    err = (wav - ref).abs().max().item()
    print("vocoder fp32 max-abs vs oracle", err, "SNR_ac vs reference golden", snr_ac_db(d["wav"], wav[0].numpy()))
    assert err < 1e-4
    assert np.abs(wav[0].numpy() - d["wav"]).max() < 1e-4


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
def test_vocoder_split_fp16_tensor_cores_is_fp32_grade(state_dicts, eng_split, golden_dir):
    """voc_precision = 'split': every vocoder conv on tcgen05 with x_hi*W_hi + x_hi*W_lo + x_lo*W_hi, fp32 residual stream.
    Same tolerance as the fp32 FFMA engine: wav max-abs <= 1e-4 against the oracle AND against the unmodified reference."""
    from oracle.incremental import HifiGanOracle
    d = np.load(os.path.join(golden_dir, "vocoder_24f.npz"))
    mel = torch.from_numpy(d["mel"])[None].repeat(2, 1, 1)
    mel[1] = mel[1].flip(0)
    o = HifiGanOracle(state_dicts[2])
    o.reset(2)
    with torch.no_grad():
        ref = torch.cat([o.step(mel[:, i:i + 4]) for i in range(0, 24, 4)], 1)
    l0 = eng_split.launch_count
    wav = _run_vocoder(eng_split, mel, [7, 2])
    assert eng_split.launch_count > l0
    err = (wav - ref).abs().max().item()
    print("vocoder split-fp16 max-abs vs oracle", err, "vs reference golden", np.abs(wav[0].numpy() - d["wav"]).max(),
          "SNR_ac", snr_ac_db(d["wav"], wav[0].numpy()))
    assert err < 1e-4
    assert np.abs(wav[0].numpy() - d["wav"]).max() < 1e-4
    assert snr_ac_db(d["wav"], wav[0].numpy()) > 80.0


@pytest.mark.parametrize("which", ["eng_fp16_ffma", "eng_tc"])
def test_vocoder_fp16_operands_snr(which, request, golden_dir):
    eng = request.getfixturevalue(which)
    d = np.load(os.path.join(golden_dir, "vocoder_24f.npz"))
    mel = torch.from_numpy(d["mel"])[None]
    wav = _run_vocoder(eng, mel, [3])[0].numpy()
    s, sa = snr_db(d["wav"], wav), snr_ac_db(d["wav"], wav)
    print(which, "SNR", s, "SNR_ac", sa)
    assert sa >= SNR_MIN_DB


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
def test_vocoder_tcgen05_matches_ffma_same_operands(eng_fp16_ffma, eng_tc):
    """Same fp16 operands, fp32 accumulation in TMEM vs in registers: only summation order differs."""
    mel = (torch.randn(3, 16, 80, generator=torch.Generator().manual_seed(2)) * 0.6)
    a = _run_vocoder(eng_fp16_ffma, mel, [0, 1, 2])
    b = _run_vocoder(eng_tc, mel, [4, 6, 5])
    err = (a - b).abs().max().item()
    print("tcgen05 vs ffma (fp16 operands) max-abs", err, "SNR", snr_db(a.numpy(), b.numpy()))
    assert snr_db(a.numpy(), b.numpy()) > 55.0


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
def test_vocoder_fused_resblocks_match_per_conv_path(state_dicts):
    """The fused residual-block kernel (six convs per launch, activations in shared memory, per-slot history blocks)
    computes the same fp16-operand / fp32-accumulate arithmetic as the conv-by-conv path: over several chunks
    (history hand-over between steps), with a permuted slot list, a stream count that does not fill the machine, and a
    second utterance on a reset slot."""
    g = torch.Generator().manual_seed(12)
    mel = torch.randn(5, 24, 80, generator=g) * 0.6
    mel2 = torch.randn(5, 8, 80, generator=g) * 0.6
    slots = [6, 0, 3, 7, 2]
    outs = []
    for fuse in (False, True):
        eng = _engine(state_dicts, voc_fuse_resblocks=fuse)
        assert bool(eng.cfg.voc_fuse_resblocks) == fuse
        a = _run_vocoder(eng, mel, slots)
        b = _run_vocoder(eng, mel2, slots[::-1])           # reset + different slot <-> stream mapping
        outs.append((a, b))
        eng.close()
    for x, y in zip(outs[0], outs[1]):
        err = (x - y).abs().max().item()
        print("fused vs per-conv vocoder max-abs", err, "SNR", snr_db(x.numpy(), y.numpy()))
        assert snr_db(x.numpy(), y.numpy()) > 70.0


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
def test_vocoder_tile_range_cuts_do_not_change_bits(state_dicts):
    """The two-lane fused residual-block kernel cuts a launch with fewer streams than lanes into tile ranges; a range that starts
    inside a stream re-derives the window halos from one warm-up tile (resblock_fused2.cu::LaneIter).  Where the cuts fall depends
    on the stream count, so the same stream must give the same bits alone (3 lanes), among 7 (other cut positions) and among 300
    (C = 64 runs whole streams there, C = 32 still cuts), over three chunks so that the staged history hand-over is exercised."""
    g = torch.Generator().manual_seed(21)
    one = torch.randn(1, 12, 80, generator=g) * 0.6
    base = None
    for n in (1, 7, 300):
        eng = _engine(state_dicts, max_slots=n)
        mel = torch.cat([one, torch.randn(n - 1, 12, 80, generator=g) * 0.6]) if n > 1 else one
        wav = _run_vocoder(eng, mel, list(range(n))[::-1])         # stream 0 lives in the last slot
        eng.close()
        if base is None:
            base = wav[0].clone()
        assert torch.equal(wav[0], base), n


def test_vocoder_group_blocking_is_exact(state_dicts):
    """voc_group (L2 blocking over streams) must not change results."""
    mel = (torch.randn(6, 8, 80, generator=torch.Generator().manual_seed(4)) * 0.6)
    e1 = _engine(state_dicts, voc_group=0)
    a = _run_vocoder(e1, mel, [0, 1, 2, 3, 4, 5])
    e1.close()
    e2 = _engine(state_dicts, voc_group=4)
    b = _run_vocoder(e2, mel, [0, 1, 2, 3, 4, 5])
    e2.close()
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------
# end to end against the unmodified reference's golden vectors
# ------------------------------------------------------------------------------------------
def _run_e2e(eng, ref, src, slots):
    B, T, _ = src.shape
    eng.reset_slots(slots)
    eng.open_sessions(slots, ref.cuda())
    ids = eng.ids_tensor(slots)
    wavs, mels, toks = [], [], []
    for pos in range(0, T, 4):
        chunk, emit = _chunks(src, pos)
        wav, mel, tok = eng.step(ids, chunk.cuda())
        wavs.append(wav.cpu()[:, :emit * 320]), mels.append(mel.cpu()[:, :emit]), toks.append(tok.cpu()[:, :emit])
    return torch.cat(wavs, 1), torch.cat(mels, 1), torch.cat(toks, 1)


@pytest.mark.parametrize("which,name", [("eng_fp32", "e2e_short"), ("eng_tc", "e2e_short"), ("eng_tc", "e2e_long"),
                                        ("eng_split", "e2e_short"), ("eng_split", "e2e_long")])
def test_end_to_end_vs_reference_golden(which, name, request, golden_dir):
    eng = request.getfixturevalue(which)
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    ref = synth.synth_mel(int(d["ref_frames"]), int(d["ref_seed"]))[None]
    src = synth.synth_mel(int(d["src_frames"]), int(d["src_seed"]))[None]
    wav, mel, tok = _run_e2e(eng, ref, src, [4])
    wav, mel, tok = wav[0].numpy(), mel[0].numpy(), tok[0].numpy()
    assert wav.shape == d["wav"].shape and mel.shape == d["mel"].shape
    agree = (tok == d["tokens"]).mean()
    mel_err = np.abs(mel - d["mel"]).max()
    s, sa = snr_db(d["wav"], wav), snr_ac_db(d["wav"], wav)
    print(which, name, "token agreement", agree, "mel max-abs", mel_err, "wav SNR", s, "SNR_ac", sa)
    assert agree == 1.0
    assert mel_err <= MEL_TOL
    assert sa >= SNR_MIN_DB
    if which in ("eng_fp32", "eng_split"):          # fp32-grade engines: CUDA-core fp32, and split-fp16 on the tensor cores
        assert np.abs(wav - d["wav"]).max() < 2e-4


@pytest.mark.skipif(NO_TC, reason="CONAN_TEST_NO_TC set")
def test_conan_fused_blocks_match_per_gemm_path(state_dicts):
    """block_fused_kernel (decoder residual block body: conv k5 -> x 5^-1/2 -> GELU -> 1x1; aligner feed-forward 256 -> 2048 -> 256,
    hidden activation in shared memory, partial outputs summed by the next LayerNorm) against the same split-fp16 arithmetic run
    GEMM by GEMM, and against the oracle: 12 chunks (longer than the decoder's receptive field), streams that do not fill a tile."""
    from oracle.incremental import ConanOracle
    B, T = 5, 48
    ref = torch.stack([synth.synth_mel(40 + 3 * s, 510 + s)[:40] for s in range(B)])
    tokens = torch.randint(0, 100, (B, T), generator=torch.Generator().manual_seed(3))
    tokens[:, 5:9] = 57                                                     # silent token: forced unvoiced frames
    o = ConanOracle(state_dicts[0])
    with torch.no_grad():
        o.open(ref)
        mel_o = torch.cat([o.step(tokens[:, i:i + 4]) for i in range(0, T, 4)], 1)
    mels = {}
    for fuse in (False, True):
        eng = _engine(state_dicts, lin_fuse_blocks=fuse)
        assert bool(eng.cfg.lin_fuse_blocks) == fuse
        slots = [6, 1, 4, 0, 3]
        eng.reset_slots(slots)
        eng.open_sessions(slots, ref.cuda())
        ids = eng.ids_tensor(slots)
        l0 = eng.launch_count
        mels[fuse] = torch.cat([eng.decoder_step(ids, tokens[:, i:i + 4].to(torch.int32).contiguous().cuda()).cpu() for i in range(0, T, 4)], 1)
        launches = (eng.launch_count - l0) // (T // 4)
        print("fused" if fuse else "per-GEMM", "Conan chunk step:", launches, "launches, mel max-abs vs oracle", (mels[fuse] - mel_o).abs().max().item())
        eng.close()
    assert (mels[True] - mel_o).abs().max().item() <= MEL_TOL
    assert (mels[True] - mels[False]).abs().max().item() < 2e-4


def test_step_graph_replay_is_bitwise_equal_to_eager_launches(state_dicts):
    """A chunk step replayed from its CUDA graph (captured at the second step with the same ready count and buffers) must equal
    the eager launch sequence bit for bit, while the CONTENT of the slot-id buffer changes from step to step (slot indirection:
    the graph bakes in pointers and grids, not which streams are ready) and the ready count changes in between."""
    outs = {}
    for graphs in (False, True):
        eng = _engine(state_dicts, step_graphs=graphs)
        S = 6
        ref = torch.stack([synth.synth_mel(40, 300 + s) for s in range(S)])
        src = torch.stack([synth.synth_mel(4 * 9 + 2, 400 + s) for s in range(S)])
        eng.reset_slots(list(range(S)))
        eng.open_sessions(list(range(S)), ref.cuda())
        ids4 = eng.ids_tensor([0, 1, 2, 3])
        ids2 = eng.ids_tensor([4, 5])
        ch4, ch2 = torch.empty(4, 6, 80, device="cuda"), torch.empty(2, 6, 80, device="cuda")
        w4, m4, t4 = torch.empty(4, 1280, device="cuda"), torch.empty(4, 4, 80, device="cuda"), torch.empty(4, 4, dtype=torch.int32, device="cuda")
        w2, m2, t2 = torch.empty(2, 1280, device="cuda"), torch.empty(2, 4, 80, device="cuda"), torch.empty(2, 4, dtype=torch.int32, device="cuda")
        got = []
        order = [[0, 1, 2, 3], [3, 2, 1, 0], [1, 3, 0, 2]]
        for step in range(9):
            perm = order[step % 3]
            ids4.copy_(torch.tensor(perm, dtype=torch.int32))                  # same buffer, new ready list
            ch4.copy_(src[perm, step * 4:step * 4 + 6])
            eng.step(ids4, ch4, w4, m4, t4)
            inv = torch.tensor([perm.index(s) for s in range(4)])
            got.append((w4.cpu()[inv], m4.cpu()[inv], t4.cpu()[inv]))
            if step % 2 == 0:                                                  # another ready count in between
                k = step // 2
                ch2.copy_(src[4:6, k * 4:k * 4 + 6])
                eng.step(ids2, ch2, w2, m2, t2)
                got.append((w2.cpu(), m2.cpu(), t2.cpu()))
        env = os.environ.get("CONAN_STEP_GRAPH")                               # the switch overrides the config (A/B runs)
        assert (eng.graph_replays > 0) == (graphs if env is None else env != "0")
        outs[graphs] = got
        eng.close()
    for a, b in zip(outs[False], outs[True]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_step_host_equals_device_step(eng_tc):
    eng = eng_tc
    ref = torch.stack([synth.synth_mel(40, 5), synth.synth_mel(40, 6)])
    src = torch.stack([synth.synth_mel(12, 7), synth.synth_mel(12, 8)])
    wav_a, mel_a, tok_a = _run_e2e(eng, ref, src, [0, 1])
    slots = np.array([2, 3], dtype=np.int32)
    eng.reset_slots(slots)
    eng.open_sessions(slots, ref.cuda())
    wavs = []
    for pos in range(0, 12, 4):
        chunk, _ = _chunks(src, pos)
        wav = np.empty((2, 1280), dtype=np.float32)
        mel = np.empty((2, 4, 80), dtype=np.float32)
        tok = np.empty((2, 4), dtype=np.int32)
        eng.step_host(slots, chunk.numpy(), wav, mel, tok)
        wavs.append(torch.from_numpy(wav.copy()))
    assert torch.equal(torch.cat(wavs, 1), wav_a)


# ------------------------------------------------------------------------------------------
# size-independent properties at a larger stream count
# ------------------------------------------------------------------------------------------
def test_replicated_streams_are_bitwise_identical_and_slot_order_free(state_dicts):
    """64 streams fed the same session/input must produce bit-identical outputs regardless of
    their slot or their position in the ready list (packing never mixes streams)."""
    eng = _engine(state_dicts, max_slots=64, max_ref_frames=64)
    S = 64
    ref = synth.synth_mel(48, 3)[None].repeat(S, 1, 1)
    src = synth.synth_mel(16, 4)[None].repeat(S, 1, 1)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(0)).tolist()
    wav, mel, tok = _run_e2e(eng, ref, src, perm)
    assert (wav == wav[0:1]).all() and (mel == mel[0:1]).all() and (tok == tok[0:1]).all()
    single, mel1, _ = _run_e2e(eng, ref[:1], src[:1], [9])
    assert torch.equal(single[0], wav[0]) and torch.equal(mel1[0], mel[0])
    eng.close()


def test_causality_future_input_does_not_change_past_output(eng_tc):
    """The reference's verify_causality (hifigan_causal.py:550-599) restated on the whole path:
    perturbing source frames >= t must not change wav samples < t*hop."""
    ref = synth.synth_mel(40, 1)[None]
    src = synth.synth_mel(24, 2)[None]
    src2 = src.clone()
    # frame 14 is first seen (as look-ahead) by the chunk that starts at frame 12.  (A uniform offset would be
    # invisible: the Emformer's LayerNorms remove it, so replace the tail with different content.)
    src2[:, 14:] = synth.synth_mel(24, 99)[None][:, 14:]
    a, _, _ = _run_e2e(eng_tc, ref, src, [0])
    b, _, _ = _run_e2e(eng_tc, ref, src2, [1])
    assert torch.equal(a[:, :12 * 320], b[:, :12 * 320])
    assert not torch.equal(a[:, 12 * 320:], b[:, 12 * 320:])


# ------------------------------------------------------------------------------------------
# BASELINE.json configs 1-3 at their stated stream counts (parity cases, not bench lines)
# ------------------------------------------------------------------------------------------
def test_config1_emformer_only_64_streams_vs_oracle(state_dicts):
    """configs[1]: Emformer content extractor only, 64 concurrent streams, chunkwise: every stream's encoder rows within
    1e-4 of the oracle and every token exact, over enough chunks for the left-context ring to wrap."""
    from oracle.incremental import EmformerOracle
    S, T = 64, 64
    eng = _engine(state_dicts, max_slots=S, max_ref_frames=64)
    try:
        src = torch.stack([synth.synth_mel(T, 300 + s) for s in range(S)])
        o = EmformerOracle(state_dicts[1])
        o.reset(S)
        slots = torch.randperm(S, generator=torch.Generator().manual_seed(5)).tolist()
        eng.reset_slots(slots)
        ids = eng.ids_tensor(slots)
        worst = 0.0
        for pos in range(0, T, 4):
            chunk, _ = _chunks(src, pos)
            with torch.no_grad():
                enc_ref = o.step(chunk)
                tok_ref = o.logits(enc_ref).argmax(-1)
            tok, enc, _ = eng.emformer_step(ids, chunk.cuda(), want_enc=True)
            worst = max(worst, (enc.cpu() - enc_ref).abs().max().item())
            assert (tok.cpu().long() == tok_ref).all(), f"token mismatch at pos {pos}"
        print("config 1 (Emformer, 64 streams) enc max-abs", worst)
        assert worst < 1e-4
    finally:
        eng.close()


def test_config2_vocoder_only_256_streams(state_dicts, golden_dir):
    """configs[2]: vocoder only, 256 concurrent streams.  8 distinct mel streams (checked against the oracle: SNR >= 40 dB),
    each replicated 32 times over permuted slots: all replicas bit-identical (the fused residual-block kernel gives every
    CTA several streams; nothing may leak between them)."""
    from oracle.incremental import HifiGanOracle
    S, D, T = 256, 8, 12
    eng = _engine(state_dicts, max_slots=S, max_ref_frames=64)
    try:
        g = torch.Generator().manual_seed(21)
        base = torch.stack([synth.synth_mel(T, 500 + s) for s in range(D)])
        mel = base.repeat(S // D, 1, 1)                       # stream i carries base[i % D]
        slots = torch.randperm(S, generator=g).tolist()
        wav = _run_vocoder(eng, mel, slots)
        for d in range(D):
            assert (wav[d::D] == wav[d:d + 1]).all(), f"replicas of stream {d} differ"
        o = HifiGanOracle(state_dicts[2])
        o.reset(D)
        with torch.no_grad():
            ref = torch.cat([o.step(base[:, i:i + 4]) for i in range(0, T, 4)], 1)
        worst = min(snr_ac_db(ref[d].numpy(), wav[d].numpy()) for d in range(D))
        print("config 2 (vocoder, 256 streams) worst SNR_ac", worst)
        assert worst >= SNR_MIN_DB
    finally:
        eng.close()


def test_config3_full_pipeline_1024_streams_properties(state_dicts):
    """configs[3] at full size: 1024 streams (16 distinct sessions/inputs x 64 replicas) through session setup and three
    chunk steps.  Size-independent checks: replicas bit-identical; a stream's output equals the same stream run alone."""
    S, D = 1024, 16
    eng = _engine(state_dicts, max_slots=S, max_ref_frames=64)
    try:
        ref = torch.stack([synth.synth_mel(40, 700 + s) for s in range(D)]).repeat(S // D, 1, 1)
        src = torch.stack([synth.synth_mel(12, 800 + s) for s in range(D)]).repeat(S // D, 1, 1)
        slots = torch.randperm(S, generator=torch.Generator().manual_seed(8)).tolist()
        eng.reset_slots(slots)
        for b in range(0, S, 32):                             # session batches of 32, as the scheduler does
            eng.open_sessions(slots[b:b + 32], ref[b:b + 32].cuda())
        ids = eng.ids_tensor(slots)
        wavs, mels, toks = [], [], []
        for pos in range(0, 12, 4):
            chunk, emit = _chunks(src, pos)
            wav, mel, tok = eng.step(ids, chunk.cuda())
            wavs.append(wav.cpu()), mels.append(mel.cpu()), toks.append(tok.cpu())
        wav, mel, tok = torch.cat(wavs, 1), torch.cat(mels, 1), torch.cat(toks, 1)
        for d in range(D):
            assert (wav[d::D] == wav[d:d + 1]).all() and (mel[d::D] == mel[d:d + 1]).all() and (tok[d::D] == tok[d:d + 1]).all()
        # ... and the 16 distinct streams against the CPU oracle (lock-step batch of 16), not only replica equality
        from oracle.incremental import StreamingOracle
        o = StreamingOracle(*state_dicts)
        wav_o, mel_o, tok_o = o.infer(ref[:D], src[:D])
        assert torch.equal(tok[:D].long(), tok_o)
        mel_err = (mel[:D] - mel_o).abs().max().item()
        worst = min(snr_ac_db(wav_o[d].numpy(), wav[d].numpy()) for d in range(D))
        print("1024 streams: mel max-abs vs oracle", mel_err, "worst wav SNR_ac over 16 distinct streams", worst)
        assert mel_err <= MEL_TOL and worst >= SNR_MIN_DB
        single, mel1, tok1 = _run_e2e(eng, ref[3:4], src[3:4], [slots[0]])
        assert torch.equal(tok1[0], tok[3]) and torch.equal(mel1[0], mel[3, :mel1.shape[1]])
        assert torch.equal(single[0], wav[3, :single.shape[1]])
    finally:
        eng.close()


def test_full_step_changing_ready_subsets_different_ages_and_recycled_slots(state_dicts, eng_tc):
    """The packing path end to end (ChunkScheduler -> conan_step_host -> compact buffers + history gather / scatter of every
    sub-model): the ready subset changes every step (frames arrive in uneven bursts), streams are opened at different times
    with references of different lengths, finished streams are closed mid-run and their slots reused by new sessions while the
    other streams keep running.  Every stream must equal its own single-stream CPU oracle run."""
    from conan_b200.scheduler import ChunkScheduler
    from oracle.incremental import StreamingOracle
    eng = eng_tc
    n_streams, n_slots = 7, 3                              # 7 sessions over 3 slots: every slot is recycled at least once
    rng = np.random.default_rng(5)
    T_ref = [40, 23, 64, 40, 9, 31, 50]
    T_src = [26, 41, 18, 33, 22, 37, 14]
    refs = [synth.synth_mel(T_ref[i], 900 + i) for i in range(n_streams)]
    srcs = [synth.synth_mel(T_src[i], 950 + i) for i in range(n_streams)]
    expect = []
    for i in range(n_streams):
        o = StreamingOracle(*state_dicts)
        expect.append(o.infer(refs[i][None], srcs[i][None]))
    sch = ChunkScheduler(eng, n_slots, capacity_frames=16)
    waiting = list(range(n_streams))
    live = {}                                              # stream index -> [sid, frames fed]
    got = {i: ([], [], []) for i in range(n_streams)}
    subsets, slots_used = set(), {}
    for it in range(400):
        while waiting and len(live) < n_slots and (it % 3 == 0 or not live):      # staggered admission
            i = waiting.pop(0)
            sid = sch.open(refs[i].numpy())
            live[i] = [sid, 0]
            slots_used.setdefault(sch.streams[sid].slot, []).append(i)
        for i, st in live.items():                          # uneven arrivals: 0..7 frames per stream per iteration
            n = int(rng.integers(0, 8))
            n = min(n, T_src[i] - st[1])
            if n:
                sch.push(st[0], srcs[i][st[1]:st[1] + n].numpy())
                st[1] += n
            if st[1] == T_src[i] and not sch.streams[st[0]].ended:
                sch.end(st[0])
        out = sch.step()
        if out:
            subsets.add(tuple(sorted(out)))
        by_sid = {st[0]: i for i, st in live.items()}
        for sid, (w, m, t) in out.items():
            i = by_sid[sid]
            got[i][0].append(w), got[i][1].append(m), got[i][2].append(t)
        for i in [i for i, st in live.items() if sch.finished(st[0])]:
            sch.close(live.pop(i)[0])
        if not waiting and not live:
            break
    assert not waiting and not live
    assert len(subsets) >= 6 and any(len(v) >= 2 for v in slots_used.values())      # subsets really changed, slots really reused
    for i in range(n_streams):
        wav, mel, tok = np.concatenate(got[i][0]), np.concatenate(got[i][1]), np.concatenate(got[i][2])
        wav_o, mel_o, tok_o = (x[0].numpy() for x in expect[i])
        assert wav.shape == wav_o.shape and mel.shape == mel_o.shape
        assert (tok == tok_o).all(), i
        mel_err = np.abs(mel - mel_o).max()
        sa = snr_ac_db(wav_o, wav)
        print(f"stream {i} (T_ref {T_ref[i]}, T {T_src[i]}): mel max-abs {mel_err:.2e}, wav SNR_ac {sa:.1f} dB")
        assert mel_err <= MEL_TOL and sa >= SNR_MIN_DB


def test_fast_system_right_context_zero(state_dicts):
    """the released "*_fast" system runs the same graphs with right_context = 0 (4-row chunks, no look-ahead
    keys).  Emformer rows and tokens against the oracle configured the same way, then the whole step end to end against
    the streaming oracle driven with 4-frame chunks."""
    from oracle.incremental import EmformerOracle, assemble_chunk
    hp = dict(synth.DEFAULT_HP, right_context=0)
    eng = _engine(state_dicts, hp=hp)
    try:
        assert eng.rows_in == 4
        B, T = 2, 64
        src = torch.stack([synth.synth_mel(T, 900 + s) for s in range(B)])
        o = EmformerOracle(state_dicts[1], right_context=0)
        o.reset(B)
        slots = [3, 1]
        eng.reset_slots(slots)
        ids = eng.ids_tensor(slots)
        worst = 0.0
        for pos in range(0, T, 4):
            chunk, _ = assemble_chunk(src, pos, rc=0)
            with torch.no_grad():
                enc_ref = o.step(chunk.contiguous())
                tok_ref = o.logits(enc_ref).argmax(-1)
            tok, enc, _ = eng.emformer_step(ids, chunk.contiguous().cuda(), want_enc=True)
            worst = max(worst, (enc.cpu() - enc_ref).abs().max().item())
            assert (tok.cpu().long() == tok_ref).all()
        print("rc=0 emformer enc max-abs", worst)
        assert worst < 1e-4
    finally:
        eng.close()


def test_emformer_fused_ffn_matches_two_gemm_path(state_dicts):
    """The fused feed-forward kernel (hidden activation kept in shared memory, partial sums reduced by the LayerNorm) against
    the two-GEMM path on the same split-fp16 operands, for a stream count that leaves a partial row tile."""
    from oracle.incremental import EmformerOracle
    B, T = 37, 24                                   # 37 * 6 = 222 rows: one full and one partial 128-row tile
    src = torch.stack([synth.synth_mel(T, 600 + s) for s in range(B)])
    outs = []
    for fuse in (False, True):
        eng = _engine(state_dicts, max_slots=40, lin_fuse_ffn=fuse)
        assert bool(eng.cfg.lin_fuse_ffn) == fuse
        slots = list(range(B))[::-1]
        eng.reset_slots(slots)
        ids = eng.ids_tensor(slots)
        encs, toks = [], []
        for pos in range(0, T, 4):
            chunk, _ = _chunks(src, pos)
            tok, enc, _ = eng.emformer_step(ids, chunk.cuda(), want_enc=True)
            encs.append(enc.cpu()), toks.append(tok.cpu())
        outs.append((torch.cat(encs, 1), torch.cat(toks, 1)))
        eng.close()
    err = (outs[0][0] - outs[1][0]).abs().max().item()
    print("fused FFN vs two-GEMM enc max-abs", err)
    assert err < 5e-5 and torch.equal(outs[0][1], outs[1][1])
    o = EmformerOracle(state_dicts[1])
    o.reset(B)
    with torch.no_grad():
        ref = torch.cat([o.step(_chunks(src, pos)[0]) for pos in range(0, T, 4)], 1)
    assert (outs[1][0] - ref).abs().max().item() < 1e-4


def test_step_host_pipelined_equals_synchronous(eng_tc):
    """conan_step_host_submit / _wait (two steps in flight, result copies on the engine's copy stream) returns exactly what the
    synchronous conan_step_host returns, step by step."""
    eng = eng_tc
    ref = torch.stack([synth.synth_mel(40, 5), synth.synth_mel(40, 6)])
    src = torch.stack([synth.synth_mel(24, 7), synth.synth_mel(24, 8)])
    slots = np.array([2, 3], dtype=np.int32)

    def run(pipelined):
        eng.reset_slots(slots)
        eng.open_sessions(slots, ref.cuda())
        chunks = [np.ascontiguousarray(_chunks(src, pos)[0].numpy()) for pos in range(0, 24, 4)]
        wavs = [np.empty((2, 1280), dtype=np.float32) for _ in chunks]
        mels = [np.empty((2, 4, 80), dtype=np.float32) for _ in chunks]
        toks = [np.empty((2, 4), dtype=np.int32) for _ in chunks]
        if not pipelined:
            for c, w, m, t in zip(chunks, wavs, mels, toks):
                eng.step_host(slots, c, w, m, t)
        else:
            prev = None
            for c, w, m, t in zip(chunks, wavs, mels, toks):
                tk = eng.step_host_submit(slots, c, w, m, t)
                if prev is not None:
                    eng.step_host_wait(prev)
                prev = tk
            eng.step_host_wait(prev)
            with pytest.raises(RuntimeError):
                eng.step_host_wait(prev)                     # nothing in flight any more
        return np.concatenate(wavs, 1), np.concatenate(mels, 1), np.concatenate(toks, 1)

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_vocoder_other_kernel_sizes_and_dilations():
    """A vocoder config other than the reference's (resblock kernels 5 / 9, dilations 1, 2, 4): the fused residual-block kernels
    take their generic-kernel-size path; fp16-operand result against the oracle, and the fused path against the conv-by-conv path."""
    from conan_b200.engine import Engine, make_config
    from oracle.incremental import HifiGanOracle
    voc_hp = dict(synth.DEFAULT_VOC_HP, resblock_kernel_sizes=[5, 9], resblock_dilation_sizes=[[1, 2, 4], [1, 2, 4]])
    sds = synth.make_all_state_dicts(77, voc_hp=voc_hp)
    mel = torch.randn(3, 12, 80, generator=torch.Generator().manual_seed(6)) * 0.6
    o = HifiGanOracle(sds[2], resblock_kernel_sizes=(5, 9), resblock_dilation_sizes=((1, 2, 4), (1, 2, 4)))
    o.reset(3)
    with torch.no_grad():
        ref = torch.cat([o.step(mel[:, i:i + 4]) for i in range(0, 12, 4)], 1)
    outs = []
    for fuse in (True, False):
        eng = Engine(*sds, make_config(voc_hp=voc_hp, max_slots=8, max_ref_frames=64, voc_fuse_resblocks=fuse))
        outs.append(_run_vocoder(eng, mel, [5, 1, 2]))
        eng.close()
    s = min(snr_ac_db(ref[b].numpy(), outs[0][b].numpy()) for b in range(3))
    print("k = 5 / 9, dilations 1 2 4: worst SNR_ac", s, "fused vs per-conv SNR", snr_db(outs[1].numpy(), outs[0].numpy()))
    assert s >= SNR_MIN_DB
    assert snr_db(outs[1].numpy(), outs[0].numpy()) > 70.0
