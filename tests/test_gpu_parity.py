"""GPU parity tests (-m gpu): libwhisper_b200.so through its C ABI on a B200 against the oracle (compiled reference,
oracle/_ref) on the same inputs.  Tolerances are SURVEY.md §8c's: mel exact; embd_enc rel-L2 <= 2e-3 and max-abs <= 2e-2;
logits max-abs <= 5e-2 with identical argmax and top-5 set; greedy token ids, text and token timestamps exact."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, ids_of
from oracle import ref_lib
import whisper_b200 as wb

sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_model  # noqa: E402

pytestmark = pytest.mark.gpu

SOT, BEG = 50257, 50363


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# ---- kernels -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (384, 1500, 384), (64, 1500, 1504), (1500, 1500, 64), (384, 300, 240),
                                   (1536, 1500, 384), (384, 1500, 1536), (51864, 40, 384), (1152, 9, 384), (512, 3000, 1536),
                                   (200, 130, 72)])
def test_tcgen05_gemm_matches_numpy_and_simt(product, M, N, K):
    """UMMA/TMA contraction incl. every tail (K % 64, M % 128, N % 128) vs f32 numpy and vs the SIMT engine."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    B = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
    ref = B.astype(np.float32) @ A.astype(np.float32).T
    tc, _ = wb.gemm_f16(A, B, engine=0)
    simt, _ = wb.gemm_f16(A, B, engine=1)
    tol = 2e-3 * np.sqrt(K / 64.0)
    assert np.abs(tc - ref).max() <= tol
    assert np.abs(simt - ref).max() <= tol


def test_gemm_linearity_at_full_size(product):
    """Size-independent property at the encoder's largest shape: C(A, B1 + B2) == C(A, B1) + C(A, B2) for exactly representable sums."""
    rng = np.random.default_rng(5)
    A = rng.integers(-4, 5, size=(2048, 512)).astype(np.float16)
    B1 = rng.integers(-4, 5, size=(12000, 512)).astype(np.float16)
    B2 = rng.integers(-4, 5, size=(12000, 512)).astype(np.float16)
    c1, _ = wb.gemm_f16(A, B1)
    c2, _ = wb.gemm_f16(A, B2)
    c12, _ = wb.gemm_f16(A, (B1 + B2).astype(np.float16))
    assert np.array_equal(c12, c1 + c2)            # small integers: every partial sum is exact in f32
    assert np.array_equal(c1, B1.astype(np.float32) @ A.astype(np.float32).T)


def _gelu_table():
    import ctypes as C
    lib = wb.load_library()
    g, e = np.zeros(65536, np.uint16), np.zeros(65536, np.uint16)
    u16p = C.POINTER(C.c_uint16)
    lib.whisper_b200_f16_tables(g.ctypes.data_as(u16p), e.ctypes.data_as(u16p))
    return g.view(np.float16)


@pytest.mark.parametrize("N,M,K", [(128, 128, 64), (1500, 384, 384), (3000, 1536, 384), (1500, 384, 1536), (200, 256, 72), (24000, 1152, 384),
                                   (12000, 2048, 512), (77, 128, 1504)])
def test_encoder_gemm_tma_store_epilogues(product, N, M, K):
    """csrc/cuda/gemm_enc.cu (TMA -> tcgen05 -> TMEM -> tile assembled in shared memory -> TMA store) in every epilogue mode against
    numpy with the reference's rounding points (f32 accumulate, bias in f32, GELU through the f16 table, outputs rounded to f16):
    plain f16, GELU, transposed f16 (V^T), f32 with residual, three scaled segments; tile tails in N (clipped by the tensor map) and
    in K (zero-filled by TMA)."""
    rng = np.random.default_rng(N + 3 * M + 7 * K)
    act = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
    wgt = (rng.standard_normal((M, K)) * 0.25).astype(np.float16)
    bias = rng.standard_normal(M).astype(np.float32)
    res = rng.standard_normal((N, M)).astype(np.float32)
    acc = act.astype(np.float32) @ wgt.astype(np.float32).T
    tol = 2e-3 * np.sqrt(K / 64.0)
    half_tol = lambda ref: tol + np.abs(ref) * 2.0 ** -10          # one f16 rounding on top of the accumulation-order tolerance
    o0, _ = wb.gemm_enc_probe(act, wgt, 0, bias=bias)
    ref0 = acc + bias
    assert (np.abs(o0.astype(np.float32) - ref0) <= half_tol(ref0)).all()
    o1, _ = wb.gemm_enc_probe(act, wgt, 1, bias=bias)
    tab = _gelu_table()
    ref1 = tab[ref0.astype(np.float16).view(np.uint16)].astype(np.float32)
    # the table is evaluated at f16(acc + bias): where our accumulation rounds to a neighbouring f16 the GELU moves by about one input ulp
    assert (np.abs(o1.astype(np.float32) - ref1) <= half_tol(ref0) * 1.2 + 2.0 ** -10).all()
    exact = tab[o0.view(np.uint16)]                                # and it must be EXACTLY the table entry of our own pre-activation
    assert np.array_equal(o1.view(np.uint16), exact.view(np.uint16))
    o2, _ = wb.gemm_enc_probe(act, wgt, 2, bias=bias)
    assert np.array_equal(o2[:, :N], o0.T)
    # (the tensor map clips the innermost dimension at 16-byte granularity: the up to 7 padding columns behind token N - 1 may
    # receive the bias of zero-filled rows — finite values that every consumer multiplies by a zero probability)
    assert np.isfinite(o2.astype(np.float32)).all()
    o3, _ = wb.gemm_enc_probe(act, wgt, 3, bias=bias, res=res)
    ref3 = (acc + bias) + res
    assert (np.abs(o3 - ref3) <= tol).all()
    if M % 384 == 0:
        o4, _ = wb.gemm_enc_probe(act, wgt, 4, bias=bias)
        ref4 = (acc + bias) * 0.25
        assert (np.abs(o4.astype(np.float32) - ref4) <= half_tol(ref4)).all()


def _exp_table():
    import ctypes as C
    lib = wb.load_library()
    g, e = np.zeros(65536, np.uint16), np.zeros(65536, np.uint16)
    u16p = C.POINTER(C.c_uint16)
    lib.whisper_b200_f16_tables(g.ctypes.data_as(u16p), e.ctypes.data_as(u16p))
    return e.view(np.float16)


def attention_reference(q, k, vt, n_head, T):
    """ggml's arithmetic for one encoder attention (whisper.cpp:1880-1917, ggml.c:11116-11201) in numpy: f32 scores, scaled by 1/8, row
    maximum, exp through the f16 table at f16(s - max), f64 sum, p = f16(e * (float)(1 / sum)), f32 product with V, f16 output."""
    tab = _exp_table()
    B, _, d = q.shape
    out = np.zeros((B, T, d), np.float32)
    for b in range(B):
        for h in range(n_head):
            sl = slice(64 * h, 64 * h + 64)
            s = (q[b, :, sl].astype(np.float32) @ k[b, :, sl].astype(np.float32).T) * np.float32(0.125)
            x = (s - s.max(axis=1, keepdims=True)).astype(np.float16)
            e = tab[x.view(np.uint16)].astype(np.float32)
            inv = (1.0 / e.astype(np.float64).sum(axis=1, keepdims=True)).astype(np.float32)
            p = (e * inv).astype(np.float16).astype(np.float32)
            out[b, :, sl] = p @ vt[b, sl, :T].astype(np.float32).T
    return out


@pytest.mark.parametrize("B,T,n_head", [(2, 1500, 6), (1, 178, 6), (3, 678, 8), (1, 128, 6), (2, 1000, 6)])
def test_fused_encoder_attention_variants(product, B, T, n_head):
    """csrc/cuda/attn_enc.cu on host buffers: every kernel configuration (softmax warps per lane quadrant, ring depths, f16 or integer exp
    table, two or three score buffers, in-order or prefetching score pipeline, one work item per CTA or a persistent CTA per SM, three score tiles per item reused from registers) must produce the SAME BITS — they differ only in who reads which score columns and in how the exact sum is formed — and the
    result matches the numpy restatement of ggml's soft_max between two mul_mats up to the accumulation order of the tensor cores
    (a score that rounds to the neighbouring f16 argument moves its exponential by 2^-11 relative)."""
    d = 64 * n_head
    Tp = (T + 7) & ~7
    rng = np.random.default_rng(B * 1000 + T)
    q = (rng.standard_normal((B, T, d)) * 1.5).astype(np.float16)
    k = (rng.standard_normal((B, T, d)) * 1.5).astype(np.float16)
    vt = np.zeros((B, d, Tp), np.float16)
    vt[:, :, :T] = rng.standard_normal((B, d, T)).astype(np.float16)
    outs = [wb.attn_enc_probe(q, k, vt, n_head, variant=v)[0] for v in range(9)]
    for v in range(1, 9):
        assert np.array_equal(outs[0].view(np.uint16), outs[v].view(np.uint16)), f"variant {v} differs from variant 0"
    ref = attention_reference(q, k, vt, n_head, T)
    err = np.abs(outs[0].astype(np.float32) - ref)
    assert err.max() <= 2e-2 and np.sqrt((err ** 2).sum() / (ref ** 2).sum()) <= 2e-3, (err.max(),)


# ---- stages on real tiny.en weights ---------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def encoded(gpu_ctx, ref_session, jfk):
    assert ref_session.pcm_to_mel(jfk, 4) == 0 and gpu_ctx.pcm_to_mel(jfk, 4) == 0
    ref_session.lib.probe_set_audio_ctx(ref_session.ctx, 0)
    assert ref_session.encode(0, 8) == 0 and gpu_ctx.encode(0) == 0
    return gpu_ctx, ref_session


def test_mel_bit_exact(encoded):
    ctx, ref = encoded
    rmel, _ = ref.mel()
    assert np.array_equal(ctx.read_stage(wb.STAGE_HOST_MEL, np.float32).reshape(rmel.shape), rmel)


def test_conv_stem(encoded):
    ctx, ref = encoded
    conv_ref = ref.embd_conv().T                       # reference holds [d][T]
    conv = ctx.read_stage(wb.STAGE_EMBD_CONV, np.float32).reshape(conv_ref.shape)
    assert rel_l2(conv, conv_ref) <= 1e-3
    assert np.abs(conv - conv_ref).max() <= 1e-2        # one f16 ulp at |x| ~ 4 after the GELU table
    assert (conv == conv_ref).mean() > 0.9              # table-driven GELU: most outputs are bit-identical


def test_encoder_output(encoded):
    ctx, ref = encoded
    enc_ref = ref.embd_enc()
    enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(enc_ref.shape)
    assert not np.isnan(enc).any()
    assert rel_l2(enc, enc_ref) <= 2e-3
    assert np.abs(enc - enc_ref).max() <= 2e-2


def test_cross_kv(encoded):
    ctx, ref = encoded
    kr, vr = ref.kv_cross()
    k = ctx.read_stage(wb.STAGE_CROSS_K, np.float16)
    v = ctx.read_stage(wb.STAGE_CROSS_V, np.float16)
    assert k.size == kr.size and v.size == vr.size
    assert rel_l2(k, kr) <= 2e-3 and rel_l2(v, vr) <= 2e-3
    assert np.abs(k.astype(np.float32) - kr.astype(np.float32)).max() <= 1e-2
    assert np.abs(v.astype(np.float32) - vr.astype(np.float32)).max() <= 2e-2


def check_logits(mine, ref):
    assert np.abs(mine - ref).max() <= 5e-2
    assert int(mine.argmax()) == int(ref.argmax())
    # top-5 set, up to near-ties at its edge: whatever we rank in our top 5 must be within 2e-2 of the reference's 5th value
    fifth = np.sort(ref)[-5]
    assert (ref[np.argsort(mine)[-5:]] >= fifth - 2e-2).all()


def test_decoder_logits_skinny_and_tensor_core_paths(encoded):
    ctx, ref = encoded
    check_logits(ctx.decode([SOT], 0), ref.decode([SOT], 0, 4))                   # 1 row: skinny kernels
    check_logits(ctx.decode([BEG], 1), ref.decode([BEG], 1, 4))
    toks = [843, 523, 616, 5891, 3399, 1265, 407, 644, 534, 1499, 460, 466]
    check_logits(ctx.decode(toks, 2), ref.decode(toks, 2, 4))                     # 12 rows: mma.sync skinny path + causal mask
    check_logits(ctx.decode([329], 14), ref.decode([329], 14, 4))                 # reads the cache written by both paths
    more = [345, 1265, 644, 345, 460, 466, 329, 534, 1499, 13] * 4
    check_logits(ctx.decode(more, 15), ref.decode(more, 15, 4))                   # 40 rows: tcgen05 path
    check_logits(ctx.decode([50256], 55), ref.decode([50256], 55, 4))
    kr, vr = ref.kv_self()
    k = ctx.read_stage(wb.STAGE_SELF_K, np.float16).reshape(4, -1, 384)[:, :56]
    v = ctx.read_stage(wb.STAGE_SELF_V, np.float16).reshape(4, 384, -1)[:, :, :56]
    assert rel_l2(k, kr.reshape(4, -1, 384)[:, :56]) <= 2e-3
    assert rel_l2(v, vr.reshape(4, 384, -1)[:, :, :56]) <= 2e-3


def test_simt_engine_agrees_with_tensor_cores(gpu_ctx, jfk):
    gpu_ctx.pcm_to_mel(jfk, 4)
    gpu_ctx.set_gemm_engine(1)
    try:
        assert gpu_ctx.encode(0) == 0
        enc_simt = gpu_ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32)
    finally:
        gpu_ctx.set_gemm_engine(0)
    assert gpu_ctx.encode(0) == 0
    enc_tc = gpu_ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32)
    assert rel_l2(enc_tc, enc_simt) <= 2e-3
    assert gpu_ctx.encode(0) == 0                                                # determinism: same launch, same bits
    assert np.array_equal(gpu_ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32), enc_tc)


@pytest.mark.parametrize("audio_ctx", [178, 678, 1000])
def test_dynamic_audio_ctx(gpu_ctx, ref_session, jfk, audio_ctx):
    """CaptureStreamToText sets audio_ctx = t*50 + 128 (capture_stream_to_text.gd:84): arbitrary, not tile aligned."""
    pr = ref_lib.host_params(ref_session.lib, max_tokens=0, n_threads=4, temperature_inc=0.0, audio_ctx=audio_ctx)
    pm = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0, audio_ctx=audio_ctx)
    n = min(len(jfk), (audio_ctx - 128) * 320) if audio_ctx < 600 else len(jfk)
    assert ref_session.full(pr, jfk[:n]) == 0 and gpu_ctx.full(pm, jfk[:n]) == 0
    assert ids_of(gpu_ctx.result()) == ids_of(ref_session.result())
    ref_session.lib.probe_set_audio_ctx(ref_session.ctx, 0)


# ---- whisper_full ------------------------------------------------------------------------------------------------------------

def test_device_mel_equals_reference(gpu_ctx, ref_session, jfk):
    """whisper_full computes the log-mel spectrogram of a clip of up to 30 s on the device (cuda/mel_kernels.cu: the operations of
    csrc/mel.cpp in the same order with explicit roundings).  Stated tolerance: abs <= 1e-6; expected: bit-identical —
    the only operation that is not the same code as on the host is the f64 log10, and two correctly-behaved log10 implementations
    round to different f32 values only if the true value lies within ~1e-15 of a rounding boundary.  Speech, 30 s of it, noise, a clip
    that ends inside a frame, and near-silence."""
    rng = np.random.default_rng(11)
    clips = [jfk, ref_lib.jfk30(jfk), (rng.standard_normal(48000 + 123) * 0.1).astype(np.float32), jfk[:33333],
             (rng.standard_normal(32000) * 1e-4).astype(np.float32), np.roll(ref_lib.jfk30(jfk), 12345)]
    p = wb.host_params(gpu_ctx.lib, max_tokens=1, n_threads=4, temperature_inc=0.0)
    for c in clips:
        assert gpu_ctx.full(p, c) == 0
        assert ref_session.pcm_to_mel(c, 4) == 0
        rmel, _ = ref_session.mel()
        mine = gpu_ctx.read_stage(wb.STAGE_DEVICE_MEL, np.float32).reshape(rmel.shape)
        assert np.abs(mine - rmel).max() <= 1e-6
        assert int((mine != rmel).sum()) == 0


def test_host_mel_switch_gives_the_same_transcript(gpu_ctx, jfk, monkeypatch):
    """WHISPER_B200_HOST_MEL=1 keeps the spectrogram on the host (csrc/mel.cpp, the path longer audio always takes)."""
    audio = ref_lib.jfk30(jfk)
    p = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    assert gpu_ctx.full(p, audio) == 0
    dev = gpu_ctx.result()
    monkeypatch.setenv("WHISPER_B200_HOST_MEL", "1")
    assert gpu_ctx.full(p, audio) == 0
    assert_same_transcript(gpu_ctx.result(), dev)
    long = np.concatenate([audio, jfk])                  # 41 s: two windows, host spectrogram
    monkeypatch.delenv("WHISPER_B200_HOST_MEL")
    assert gpu_ctx.full(p, long) == 0
    assert len(ids_of(gpu_ctx.result())) > 80



def assert_same_transcript(rm, rr, eot=50256):
    """(A segment that ends in a text token: the reference clamps that token's t1 against the element one past the end of its token
    vector, whisper.cpp:6547 — heap garbage — so that one value is not compared; see tests/test_hostlogic.py::assert_same_result.)"""
    assert ids_of(rm) == ids_of(rr)
    assert rm["text"] == rr["text"]
    times = lambda r: [(t["t0"], None if (t is s["tokens"][-1] and t["id"] < eot) else t["t1"], t["tid"]) for s in r["segments"] for t in s["tokens"]]
    assert times(rm) == times(rr)
    for k in ("p", "pt", "ptsum"):
        a = np.array([t[k] for s in rm["segments"] for t in s["tokens"]])
        b = np.array([t[k] for s in rr["segments"] for t in s["tokens"]])
        assert np.abs(a - b).max() <= 5e-3, k


@pytest.mark.parametrize("max_tokens", [16, 0])
def test_greedy_transcript_exact(gpu_ctx, ref_session, jfk, max_tokens):
    pr = ref_lib.host_params(ref_session.lib, max_tokens=max_tokens, n_threads=4)
    pm = wb.host_params(gpu_ctx.lib, max_tokens=max_tokens, n_threads=4)
    assert ref_session.full(pr, jfk) == 0 and gpu_ctx.full(pm, jfk) == 0
    assert_same_transcript(gpu_ctx.result(), ref_session.result())
    if max_tokens == 0:
        assert gpu_ctx.result()["text"] == (b" And so my fellow Americans ask not what your country can do for you ask what you "
                                            b"can do for your country.")       # audio_transcribe.tscn:23


def test_thirty_second_chunk_greedy_exact(gpu_ctx, ref_session, jfk):
    audio = ref_lib.jfk30(jfk)
    pr = ref_lib.host_params(ref_session.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    pm = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    assert ref_session.full(pr, audio) == 0 and gpu_ctx.full(pm, audio) == 0
    assert_same_transcript(gpu_ctx.result(), ref_session.result())
    assert len(ids_of(gpu_ctx.result())) > 60


def test_host_logits_path_still_exact(gpu_ctx, ref_session, jfk, monkeypatch):
    """WHISPER_B200_DEVICE_SAMPLING=0: logits come back to the host and csrc/decode_host.cpp applies the rules (the path beam
    search and t > 0 sampling always take).  Same transcript, and the device sampler agrees with it token for token."""
    pm = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    assert gpu_ctx.full(pm, jfk) == 0
    dev = gpu_ctx.result()
    monkeypatch.setenv("WHISPER_B200_DEVICE_SAMPLING", "0")
    assert gpu_ctx.full(pm, jfk) == 0
    host = gpu_ctx.result()
    assert ids_of(dev) == ids_of(host)
    for k in ("p", "plog", "pt", "ptsum"):
        a = np.array([t[k] for s in dev["segments"] for t in s["tokens"]])
        b = np.array([t[k] for s in host["segments"] for t in s["tokens"]])
        # the device sums exp() in f64; the host path keeps the reference's sequential f32 sum, which drops terms below
        # ~6e-8 of the running sum — p / plog differ by a few 1e-4
        assert np.abs(a - b).max() <= 2e-3, k
    assert [(t["t0"], t["t1"], t["tid"]) for s in dev["segments"] for t in s["tokens"]] == \
           [(t["t0"], t["t1"], t["tid"]) for s in host["segments"] for t in s["tokens"]]
    pr = ref_lib.host_params(ref_session.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    assert ref_session.full(pr, jfk) == 0
    assert_same_transcript(host, ref_session.result())


def test_step_kernel_agrees_with_multi_kernel_path(gpu_ctx, jfk):
    """Engine 2 routes every decode step through the separate kernels (kernels.cu); engine 0 uses the persistent decode-step
    kernel (decode_step.cu).  Same token ids and timestamps; probabilities differ only by summation order."""
    audio = ref_lib.jfk30(jfk)
    p = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    c0 = gpu_ctx.counters()["launches"]
    assert gpu_ctx.full(p, audio) == 0
    step, l_step = gpu_ctx.result(), gpu_ctx.counters()["launches"] - c0
    gpu_ctx.set_gemm_engine(2)
    try:
        c0 = gpu_ctx.counters()["launches"]
        assert gpu_ctx.full(p, audio) == 0
        multi, l_multi = gpu_ctx.result(), gpu_ctx.counters()["launches"] - c0
        lm = gpu_ctx.decode([SOT], 0)
    finally:
        gpu_ctx.set_gemm_engine(0)
    ls = gpu_ctx.decode([SOT], 0)
    assert ids_of(step) == ids_of(multi) and len(ids_of(step)) > 60
    assert l_step < l_multi / 5                                  # one launch per token step instead of ~40
    for k in ("p", "plog", "pt", "ptsum"):
        a = np.array([t[k] for s in step["segments"] for t in s["tokens"]])
        b = np.array([t[k] for s in multi["segments"] for t in s["tokens"]])
        assert np.abs(a - b).max() <= 5e-3, k
    assert [(t["t0"], t["t1"], t["tid"]) for s in step["segments"] for t in s["tokens"]] == \
           [(t["t0"], t["t1"], t["tid"]) for s in multi["segments"] for t in s["tokens"]]
    assert np.abs(ls - lm).max() <= 1e-2 and int(ls.argmax()) == int(lm.argmax())    # tensor-core vs sequential accumulation order
    assert np.array_equal(gpu_ctx.decode([SOT], 0), ls)          # determinism of the step kernel


def test_golden_fixtures(gpu_ctx, jfk):
    """Against tests/golden/jfk_tiny_en.npz (tools/make_golden.py, generated from the compiled reference)."""
    import hashlib
    g = np.load(os.path.join(ROOT, "tests", "golden", "jfk_tiny_en.npz"))
    assert gpu_ctx.pcm_to_mel(jfk, 4) == 0
    mel = gpu_ctx.read_stage(wb.STAGE_HOST_MEL, np.float32)
    assert hashlib.sha1(mel.tobytes()).digest() == bytes(g["mel_sha1"])
    assert gpu_ctx.encode(0) == 0
    enc = gpu_ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(1500, 384)[::25, ::16]
    assert rel_l2(enc, g["enc_sample"]) <= 2e-3 and np.abs(enc - g["enc_sample"]).max() <= 2e-2
    ck = gpu_ctx.read_stage(wb.STAGE_CROSS_K, np.float16).reshape(4, 1500, 384)[:, ::50, ::16]
    assert rel_l2(ck, g["cross_k_sample"]) <= 2e-3
    check_logits(gpu_ctx.decode([SOT], 0), g["logits_sot"])
    check_logits(gpu_ctx.decode([BEG], 1), g["logits_beg"])
    for mt, key in ((16, "ids_maxtok16"), (0, "ids_full")):
        assert gpu_ctx.full(wb.host_params(gpu_ctx.lib, max_tokens=mt, n_threads=4), jfk) == 0
        assert ids_of(gpu_ctx.result()) == g[key].tolist()
    r = gpu_ctx.result()
    toks = [t for s in r["segments"] for t in s["tokens"]]
    assert r["text"] == bytes(g["text_full"])
    assert [t["t0"] for t in toks] == g["t0_full"].tolist() and [t["t1"] for t in toks] == g["t1_full"].tolist()
    assert np.abs(np.array([t["p"] for t in toks]) - g["p_full"]).max() <= 5e-3
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0), ref_lib.jfk30(jfk)) == 0
    assert ids_of(gpu_ctx.result()) == g["ids_jfk30"].tolist()


def test_sixteen_chunk_batch_equals_reference(gpu_ctx, ref_session, jfk):
    """The bench workload: 16 shifted 30 s chunks in one whisper_b200_full_batch; a sample of them against the reference."""
    chunks = [np.roll(ref_lib.jfk30(jfk), int(k * 1.7 * 16000)) for k in range(16)]
    p = wb.host_params(gpu_ctx.lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
    pr = ref_lib.host_params(ref_session.lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
    assert gpu_ctx.full_batch(p, chunks) == 0
    for i in (0, 5, 11, 15):
        assert ref_session.full(pr, chunks[i]) == 0
        assert ids_of(gpu_ctx.chunk_result(i)) == ids_of(ref_session.result()), i


@pytest.mark.parametrize("groups,wide_rows", [(1, 0), (2, 0), (2, 64), (2, 256)])
def test_many_live_sequences_per_decoder_pass(product, model_bytes, ref_session, jfk, groups, wide_rows, monkeypatch):
    """More than 16 live sequences per decoder pass.  wide_rows = 0: decode-step launches only, cut into independent row groups
    of 16 (WHISPER_B200_STEP_GROUPS).  wide_rows > 0 (the default is 256): passes of more than 32 rows take the multi-kernel
    path (tcgen05 GEMMs over all rows, one attention CTA per (row, head), device sampler).  72 chunks through
    whisper_b200_full_batch; a sample of the transcripts must equal the single-chunk reference token for token."""
    monkeypatch.setenv("WHISPER_B200_STEP_GROUPS", str(groups))
    monkeypatch.setenv("WHISPER_B200_DECODE_ROWS", str(wide_rows))
    ctx = wb.Context(model_bytes, lib=product)
    try:
        base = ref_lib.jfk30(jfk)
        chunks = [np.roll(base, int(k * 1.7 * 16000)) for k in range(72)]
        p = wb.host_params(product, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
        pr = ref_lib.host_params(ref_session.lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
        assert ctx.full_batch(p, chunks) == 0
        for i in (0, 40):                                 # the live oracle on two of them ...
            assert ref_session.full(pr, chunks[i]) == 0
            assert ids_of(ctx.chunk_result(i)) == ids_of(ref_session.result()), (groups, wide_rows, i)
        gold = bench_golden()                             # ... and EVERY chunk against the ids the oracle produced for it (tools/make_bench_golden.py)
        bad = [i for i in range(72) if ctx.chunk_ids(i) != gold[i % len(gold)]]
        assert bad == [], (groups, wide_rows, bad)
    finally:
        ctx.close()


def bench_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_chunks_tiny_en.npz"))
    off = g["offsets"]
    return [g["ids"][off[k]:off[k + 1]].tolist() for k in range(len(off) - 1)]


NON_SPEECH_SYMBOLS = ["\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`", "{", "|", "}", "~", "「", "」", "『", "』",
                      "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[", "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}", "♪♪", "♪♪♪", "♩",
                      "♪", "♫", "♬", "♭", "♮", "♯"]


def reference_logprobs_after_rules(rs, raw, prefix):
    """whisper_process_logits (whisper.cpp:4493-4720) for the greedy t = 0 host block on tiny.en, in numpy, applied to the REFERENCE's
    raw logits: returns (logprobs, timestamp_logprob, max_text_logprob).  Used only to measure how close a decision was."""
    EOT, NOT = 50256, 50362
    x = raw.astype(np.float32).copy()
    if len(prefix) == 0:
        x[EOT] = -np.inf
        x[rs.tokenize(b" ")[0]] = -np.inf
    x[[NOT, SOT, 50361, 50360, 50357, 50358, 50359]] = -np.inf       # not, sot, nosp, solm, translate, transcribe, prev (english vocabulary)
    for sym in NON_SPEECH_SYMBOLS:
        for t in (sym, " " + sym):
            ids = rs.tokenize(t.encode())
            if len(ids) == 1 and rs.lib.whisper_token_to_str(rs.ctx, ids[0]) == t.encode():
                x[ids[0]] = -np.inf
    for t in (b" -", b" '"):
        ids = rs.tokenize(t)
        if len(ids) == 1:
            x[ids[0]] = -np.inf
    last_ts = len(prefix) > 0 and prefix[-1] >= BEG
    pen_ts = len(prefix) < 2 or prefix[-2] >= BEG
    if last_ts:
        if pen_ts:
            x[BEG:] = -np.inf
        else:
            x[:EOT] = -np.inf
    if len(prefix) == 0:
        x[BEG + 50 + 1:] = -np.inf                                    # max_initial_ts = 1.0 s
    ts = [t for t in prefix if t > BEG]
    if ts:
        x[BEG:BEG + (2 * (ts[-1] - BEG)) // 2] = -np.inf              # timestamps do not decrease (seek_delta / 2)
    m = x.max()
    lp = x - (np.log(np.exp(x[x > -np.inf] - m).sum()) + m)
    tl = lp[BEG:]
    ts_lp = np.log(np.exp(tl[tl > -np.inf] - tl.max()).sum()) + tl.max() if (tl > -np.inf).any() else -np.inf
    return lp, float(ts_lp), float(lp[:BEG].max())


def assert_near_tie(rs, chunk, mine, want):
    """The stated rule for a transcript that differs from the oracle's: up to the first differing token both streams are equal, and
    at that token the reference's own decision margin — between its pick and ours in its log-probabilities, or between the timestamp
    mass and the best text token when the two picks are of different kinds (whisper.cpp:4659-4684) — is below the logits tolerance
    of 5e-2: the two paths sum in a different order, which only a near-tie can turn into another token."""
    n = next(i for i in range(min(len(mine), len(want)) + 1) if i >= len(mine) or i >= len(want) or mine[i] != want[i])
    assert n < len(mine) and n < len(want), "one transcript is a strict prefix of the other"
    prefix = want[:n]
    assert rs.pcm_to_mel(chunk, 4) == 0 and rs.encode(0, 8) == 0
    raw = rs.decode([SOT] + prefix, 0, 4)
    lp, ts_lp, text_lp = reference_logprobs_after_rules(rs, raw, prefix)
    a, b = want[n], mine[n]
    if (a >= BEG) != (b >= BEG):
        margin = abs(ts_lp - text_lp)
    else:
        margin = float(lp[a] - lp[b])
    assert margin <= 5e-2, (n, a, b, margin)
    return margin


def test_bench_workload_512_chunks_every_transcript(gpu_ctx, ref_session, jfk):
    """bench.py's step — 512 shifted 30 s chunks in one whisper_b200_full_batch — with EVERY chunk compared against the token ids the
    compiled reference produced for that audio (tests/golden/bench_chunks_tiny_en.npz; the shift k * 1.7 s has period 300).  At
    least 95 % of the transcripts must be identical; every other one must differ first at a near-tie of the reference (assert_near_tie)."""
    base = ref_lib.jfk30(jfk)
    chunks = [np.roll(base, int(k * 1.7 * 16000)) for k in range(512)]
    p = wb.host_params(gpu_ctx.lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
    gold = bench_golden()
    for _ in range(2):                                    # second call: warm graphs, reused slots
        assert gpu_ctx.full_batch(p, chunks) == 0
        bad = [i for i in range(512) if gpu_ctx.chunk_ids(i) != gold[i % len(gold)]]
        assert len(bad) <= 25, bad
        margins = [assert_near_tie(ref_session, chunks[i], gpu_ctx.chunk_ids(i), gold[i % len(gold)]) for i in bad]
        print("chunks differing at a near-tie:", list(zip(bad, [round(m, 4) for m in margins])))


@pytest.mark.parametrize("max_tokens", [0, 16])
def test_temperature_fallback_with_the_real_host_block(gpu_ctx, ref_session, jfk, max_tokens):
    """The block SpeechToText::transcribe really sets (src/speech_to_text.cpp:403-413): entropy_thold 2.8 and temperature_inc left at
    its 0.2 default.  On a 30 s window the t = 0 pass fails the entropy test and whisper_full falls back to t = 0.2 with best_of = 5
    decoders drawing from std::discrete_distribution (whisper.cpp:5187-5207, 5612-5668).  The host logic is proven draw for draw on the
    CPU (tests/test_hostlogic.py, same case); here the device feeds it.  Stated rule: the fallback decisions (n_fail_p / n_fail_h, made on
    the t = 0 pass, whose ids are exact) equal the reference's; the t > 0 draws consume probabilities that differ from the reference's
    in the last bits (summation order), so a draw can differ only where the uniform variate falls within that distance of a bucket
    edge — on this clip none does and the transcript is identical."""
    audio = ref_lib.jfk30(jfk)
    pr = ref_lib.host_params(ref_session.lib, max_tokens=max_tokens, entropy_thold=2.8, temperature_inc=0.2, n_threads=4)
    pm = wb.host_params(gpu_ctx.lib, max_tokens=max_tokens, entropy_thold=2.8, temperature_inc=0.2, n_threads=4)
    c0, r0 = gpu_ctx.counters(), ref_session.counters()
    assert ref_session.full(pr, audio) == 0 and gpu_ctx.full(pm, audio) == 0
    c1, r1 = gpu_ctx.counters(), ref_session.counters()
    fails = (c1["n_fail_p"] - c0["n_fail_p"], c1["n_fail_h"] - c0["n_fail_h"])
    assert fails == (r1["n_fail_p"] - r0["n_fail_p"], r1["n_fail_h"] - r0["n_fail_h"])
    if max_tokens == 0:
        assert fails[0] >= 1
    assert ids_of(gpu_ctx.result()) == ids_of(ref_session.result())
    assert gpu_ctx.result()["text"] == ref_session.result()["text"]


def test_multi_device_context(product, model_bytes, jfk):
    """whisper_b200_init_multi: one context over two GPUs of the box; full_batch deals chunk i to device i mod 2 and the results come
    back in the caller's order (compared with the oracle golden).  Needs two devices (gpurun --gpus 2); skipped on a one-GPU box."""
    import subprocess
    n_dev = len([l for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.startswith("GPU")])
    if n_dev < 2:
        pytest.skip("one GPU visible")
    ctx = wb.Context(model_bytes, lib=product, devices=[0, 1])
    try:
        assert product.whisper_b200_n_devices(ctx.ctx) == 2
        base = ref_lib.jfk30(jfk)
        chunks = [np.roll(base, int(k * 1.7 * 16000)) for k in range(37)] + [jfk]
        p = wb.host_params(product, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
        gold = bench_golden()
        for _ in range(2):
            assert ctx.full_batch(p, chunks) == 0
            bad = [i for i in range(37) if ctx.chunk_ids(i) != gold[i]]
            assert len(bad) <= 2, bad
            assert ctx.chunk_text(37).startswith(b" And so my fellow Americans")
        assert ctx.full(wb.host_params(product, max_tokens=0, n_threads=4), jfk) == 0      # plain whisper_full runs on the first device
        assert ctx.result()["text"].startswith(b" And so my fellow Americans")
    finally:
        ctx.close()


def test_beam_search_and_prompt(gpu_ctx, ref_session, jfk):
    kw = dict(max_tokens=0, n_threads=4, strategy=wb.WHISPER_SAMPLING_BEAM_SEARCH, initial_prompt=b"A speech by the president.")
    assert ref_session.full(ref_lib.host_params(ref_session.lib, **kw), jfk) == 0
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib, **kw), jfk) == 0
    assert gpu_ctx.result()["text"] == ref_session.result()["text"]


def test_error_codes_and_edge_inputs(gpu_ctx, jfk):
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib, audio_ctx=1501), jfk) == -5          # whisper.cpp:5098-5101
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib, speed_up=True), jfk) == -1           # whisper.cpp:4973-4976
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib), jfk[:8000]) == 0                    # < 1 s: no segments
    assert gpu_ctx.result()["segments"] == []
    assert gpu_ctx.full(wb.host_params(gpu_ctx.lib), np.zeros(32000, np.float32)) == 0   # silence must not crash
    assert gpu_ctx.transcribe(jfk)[0]["id"] == BEG


def test_full_batch_equals_single_calls(gpu_ctx, jfk):
    chunks = [ref_lib.jfk30(np.roll(jfk, int(k * 1.7 * 16000))) for k in range(4)] + [jfk, jfk[:40000]]
    p = wb.host_params(gpu_ctx.lib, max_tokens=0, n_threads=4, temperature_inc=0.0)
    singles = []
    for c in chunks:
        assert gpu_ctx.full(p, c) == 0
        singles.append(ids_of(gpu_ctx.result()))
    assert gpu_ctx.full_batch(p, chunks) == 0
    for i in range(len(chunks)):
        assert ids_of(gpu_ctx.chunk_result(i)) == singles[i]


# ---- base.en shapes on synthetic weights (BASELINE.json configs[2]) -------------------------------------------------------------

def test_base_en_shapes_synthetic_weights(product, ref, model_bytes, jfk):
    m = synth_model.make_model(model_bytes, "base.en", seed=1234)
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        assert rs.pcm_to_mel(jfk, 4) == 0 and ctx.pcm_to_mel(jfk, 4) == 0
        assert rs.encode(0, 8) == 0 and ctx.encode(0) == 0
        enc_ref = rs.embd_enc()
        enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(enc_ref.shape)
        assert rel_l2(enc, enc_ref) <= 2e-3
        lr, lm = rs.decode([SOT], 0, 4), ctx.decode([SOT], 0)
        assert np.abs(lm - lr).max() <= 5e-2
        toks = list(range(1000, 1020))
        lr, lm = rs.decode(toks, 1, 4), ctx.decode(toks, 1)
        assert np.abs(lm - lr).max() <= 5e-2
    finally:
        ctx.close(); rs.close()


def teacher_forced(ctx, rs, tokens, n_steps):
    """Feeds the same token stream to both decoders one token per step (KV cache growing) and returns the worst logits error."""
    worst, agree, decided = 0.0, 0, 0
    for i in range(n_steps):
        lr, lm = rs.decode([tokens[i]], i, 4), ctx.decode([tokens[i]], i)
        worst = max(worst, float(np.abs(lm - lr).max()))
        top2 = np.sort(lr)[-2:]
        if top2[1] - top2[0] > 1e-1:                      # the reference itself is decided: the argmax must agree
            decided += 1
            agree += int(lm.argmax()) == int(lr.argmax())
    return worst, agree, decided


def test_base_en_cross_kv_and_teacher_forced_decode(product, ref, model_bytes, jfk):
    """BASELINE.json configs[2] shapes (d = 512, 8 heads, 6 + 6 layers) on seeded synthetic weights: cross-attention K / V of every
    decoder layer and a 24-step KV-cached decode fed the golden tiny.en token stream, logits within 5e-2 of the reference at every
    step.  (Free-running greedy ids are not compared on random weights: their logits are near-uniform noise, so the arg-max flips on
    differences far below the stated logits tolerance — DESIGN.md §2.)"""
    m = synth_model.make_model(model_bytes, "base.en", seed=1234)
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        audio = ref_lib.jfk30(jfk)
        assert rs.pcm_to_mel(audio, 4) == 0 and ctx.pcm_to_mel(audio, 4) == 0
        assert rs.encode(0, 8) == 0 and ctx.encode(0) == 0
        kr, vr = rs.kv_cross()
        k = ctx.read_stage(wb.STAGE_CROSS_K, np.float16)
        v = ctx.read_stage(wb.STAGE_CROSS_V, np.float16)
        assert k.size == kr.size == 6 * 1500 * 512 and v.size == vr.size
        assert rel_l2(k, kr) <= 2e-3 and rel_l2(v, vr) <= 2e-3
        toks = [SOT] + bench_golden()[0][:23]
        worst, agree, decided = teacher_forced(ctx, rs, toks, 24)
        assert worst <= 5e-2
        assert agree == decided
        # beam-shaped pass: 5 rows at once behind the cache (prompt-style batch of one sequence)
        lr, lm = rs.decode(toks[1:6], 24, 4), ctx.decode(toks[1:6], 24)
        assert np.abs(lm - lr).max() <= 5e-2
    finally:
        ctx.close(); rs.close()


def multilingual_header():
    p = os.path.join(ROOT, "oracle", "_ref", "for-tests-ggml-multilingual.bin")
    if not os.path.exists(p):
        pytest.skip("multilingual test header not staged (make -C oracle stage)")
    return open(p, "rb").read()


def test_multilingual_language_detect_and_prompt(product, ref, jfk):
    """BASELINE.json configs[3] vocabulary (51 865 tokens, language / task tokens in the prompt) on a seeded synthetic tiny-shaped
    multilingual model: whisper_lang_auto_detect (whisper.cpp:3569-3642) probabilities within 1e-3 of the reference's for all 99
    languages (same winner whenever the reference's margin exceeds that), and the three-token prompt [sot, lang, transcribe]
    followed by a teacher-forced decode within the logits tolerance."""
    import ctypes as C
    m = synth_model.make_model(multilingual_header(), "tiny", seed=1235)
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        assert ctx.lib.whisper_is_multilingual(ctx.ctx) == 1 and ctx.lib.whisper_n_vocab(ctx.ctx) == 51865
        assert rs.pcm_to_mel(jfk, 4) == 0 and ctx.pcm_to_mel(jfk, 4) == 0
        n_lang = ctx.lib.whisper_lang_max_id() + 1
        pr, pm = (C.c_float * n_lang)(), (C.c_float * n_lang)()
        lid_r = ref.whisper_lang_auto_detect(rs.ctx, 0, 4, pr)
        lid_m = ctx.lib.whisper_lang_auto_detect(ctx.ctx, 0, 4, pm)
        pr, pm = np.array(pr[:]), np.array(pm[:])
        assert lid_r >= 0 and lid_m >= 0
        assert np.abs(pr - pm).max() <= 1e-3
        top2 = np.sort(pr)[-2:]
        if top2[1] - top2[0] > 2e-3:
            assert lid_m == lid_r
        sot = ctx.lib.whisper_token_sot(ctx.ctx)
        prompt = [sot, ctx.lib.whisper_token_lang(ctx.ctx, lid_r), ctx.lib.whisper_token_transcribe(ctx.ctx)]
        assert prompt[0] == ref.whisper_token_sot(rs.ctx) == 50258
        lr, lm = rs.decode(prompt, 0, 4), ctx.decode(prompt, 0)
        assert np.abs(lm - lr).max() <= 5e-2
        toks = [t + 1 if t >= 50257 else t for t in bench_golden()[0][:16]]    # the multilingual vocabulary shifts the specials by one
        for i, t in enumerate(toks):
            lr, lm = rs.decode([t], 3 + i, 4), ctx.decode([t], 3 + i)
            assert np.abs(lm - lr).max() <= 5e-2, i
        # whisper_full with language="auto" runs the detect + prompt path end to end (ids are not compared on random weights)
        p = wb.host_params(product, max_tokens=8, n_threads=4, language=b"auto", temperature_inc=0.0)
        assert ctx.full(p, jfk) == 0
        assert ctx.lib.whisper_full_lang_id(ctx.ctx) == lid_m
    finally:
        ctx.close(); rs.close()


def test_small_shapes_synthetic_weights(product, ref, model_bytes, jfk):
    """d = 768, 12 heads, 12 + 12 layers (BASELINE.json configs[3] shapes): encoder output and decoder logits within tolerance."""
    m = synth_model.make_model(model_bytes, "small.en", seed=1235)
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        assert rs.pcm_to_mel(jfk, 4) == 0 and ctx.pcm_to_mel(jfk, 4) == 0
        assert rs.encode(0, 16) == 0 and ctx.encode(0) == 0
        enc_ref = rs.embd_enc()
        enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(enc_ref.shape)
        assert rel_l2(enc, enc_ref) <= 2e-3
        worst, agree, decided = teacher_forced(ctx, rs, [SOT] + bench_golden()[0][:7], 8)
        assert worst <= 5e-2 and agree == decided
    finally:
        ctx.close(); rs.close()


def test_large_v3_shapes_synthetic_weights(product, ref, jfk):
    """SURVEY.md §8f.4, the large-v3 half: a model file with 128 mel bands and 51 866 tokens (whisper.cpp:1135, 1161-1163) at the large
    width (d = 1280, 20 heads; two layers each side, seeded weights, a generated 128-band filter bank).  The loader takes it, the device
    log-mel follows the file's filter bank, and encoder output and decoder logits stay within tolerance.  (The width is beyond the
    decode-step kernel's shared-memory plan: every decoder step takes the multi-kernel path.)"""
    m = synth_model.make_model(multilingual_header(), "tiny", seed=77, n_mels=128, n_vocab=51866, shape=(1280, 20, 2, 1280, 20, 2))
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        assert ctx.lib.whisper_n_vocab(ctx.ctx) == 51866 and ref.whisper_token_beg(rs.ctx) == ctx.lib.whisper_token_beg(ctx.ctx) == 50365
        assert rs.pcm_to_mel(jfk, 8) == 0 and ctx.pcm_to_mel(jfk, 4) == 0
        rmel, _ = rs.mel()
        assert rmel.shape[0] == 128
        assert np.array_equal(ctx.read_stage(wb.STAGE_HOST_MEL, np.float32).reshape(rmel.shape), rmel)          # host transform, 128 bands
        assert rs.encode(0, 16) == 0 and ctx.encode(0) == 0
        enc_ref = rs.embd_enc()
        enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(enc_ref.shape)
        assert rel_l2(enc, enc_ref) <= 2e-3
        sot = 50258
        lr, lm = rs.decode([sot, 50259, 50360], 0, 8), ctx.decode([sot, 50259, 50360], 0)
        assert np.abs(lm - lr).max() <= 5e-2
        lr, lm = rs.decode([1000], 3, 8), ctx.decode([1000], 3)
        assert np.abs(lm - lr).max() <= 5e-2
        # whisper_full: the device log-mel with the file's 128-band bank, bit for bit the reference's
        p = wb.host_params(product, max_tokens=4, n_threads=4, temperature_inc=0.0)
        assert ctx.full(p, jfk) == 0
        dmel = ctx.read_stage(wb.STAGE_DEVICE_MEL, np.float32).reshape(rmel.shape)
        assert np.abs(dmel - rmel).max() <= 1e-6 and int((dmel != rmel).sum()) == 0
    finally:
        ctx.close(); rs.close()


@pytest.mark.parametrize("qname,text_equal", [("q8_0", True), ("q5_0", False)])
def test_quantised_model_files(product, ref, model_bytes, jfk, qname, text_equal):
    """SURVEY.md §8f.4, the quantised half: the real tiny.en weights as whisper.cpp's quantize tool would write them (the reference's
    own ggml_quantize_chunk, tools/synth_model.py::quantize_model).  The loader expands the blocks to f16 (the values ggml's
    dequantize_row_* yields, tests/test_abi.py::test_dequantizer_equals_ggml); the reference additionally quantises the ACTIVATIONS to
    8 bits per 32-block inside its mat-muls, which this backend does not restate.  Stated tolerance against the reference's quantised
    path: encoder output rel-L2 <= 3e-2 (measured 1.1e-2 for Q8_0 and 1.0e-2 for Q5_0: the reference's 8-bit activations, not the weight
    format, set the distance); the Q8_0 transcript of jfk.wav is identical, Q5_0 must still be the sentence."""
    m = synth_model.quantize_model(model_bytes, qname, ref)
    assert len(m) < 0.62 * len(model_bytes)
    rs = ref_lib.RefSession(ref, m, use_gpu=False)
    ctx = wb.Context(m, lib=product)
    try:
        assert rs.pcm_to_mel(jfk, 4) == 0 and ctx.pcm_to_mel(jfk, 4) == 0
        assert rs.encode(0, 8) == 0 and ctx.encode(0) == 0
        enc_ref = rs.embd_enc()
        enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(enc_ref.shape)
        err = rel_l2(enc, enc_ref)
        print(qname, "encoder rel-L2 vs the reference's quantised path:", err)
        assert err <= 3e-2
        pr = ref_lib.host_params(ref, max_tokens=0, n_threads=4, temperature_inc=0.0)
        pm = wb.host_params(product, max_tokens=0, n_threads=4, temperature_inc=0.0)
        assert rs.full(pr, jfk) == 0 and ctx.full(pm, jfk) == 0
        if text_equal:
            assert ctx.result()["text"] == rs.result()["text"]
        assert b"ask not what your country can do for you" in ctx.result()["text"]
    finally:
        ctx.close(); rs.close()
