"""Host logic of the product on a CPU-only box.  libwhisper_hostlogic.so = the product's C++ driver (model parser, log-mel,
KV-cell table, logits rules, samplers, fallback loop, segment assembly, token timestamps) linked against a TEST-ONLY forward
that hands the tensor math to the compiled reference.  Whatever differs from the reference's whisper_full() here is a bug in
the product's host code (whisper.cpp:4960-5807 restated in csrc/full.cpp, csrc/decode_host.cpp, csrc/mel.cpp)."""
import ctypes as C

import numpy as np
import pytest

from conftest import ids_of
from oracle import ref_lib
import whisper_b200 as wb


@pytest.fixture(scope="module")
def host_ctx(hostlogic, model_bytes):
    ctx = wb.Context(model_bytes, lib=hostlogic)
    yield ctx
    ctx.close()


def both(ref, ref_session, host_ctx, pcm, **kw):
    pr = ref_lib.host_params(ref, n_threads=4, **kw)
    pm = wb.host_params(host_ctx.lib, n_threads=4, **kw)
    rc_r, rc_m = ref_session.full(pr, pcm), host_ctx.full(pm, pcm)
    return rc_r, rc_m, ref_session.result(), host_ctx.result()


def assert_same_result(rr, rm, last_t1=True, eot=50256):
    """last_t1=False: the end time of a segment's last token is not compared — the reference clamps it against tokens[j + 1] one
    past the end of the vector (whisper.cpp:6547, `j < ns - 1` with ns the sample count), i.e. against heap garbage, whenever the
    audio is still loud at that point; the product leaves it unclamped.  That read can only happen when the vector ends in a text
    token (timestamp / EOT tokens are skipped, whisper.cpp:6510), so such a token's t1 is never compared (`eot`: first special id)."""
    assert ids_of(rr) == ids_of(rm)
    assert rr["text"] == rm["text"]
    assert len(rr["segments"]) == len(rm["segments"])
    for sr, sm in zip(rr["segments"], rm["segments"]):
        assert (sr["t0"], sr["t1"]) == (sm["t0"], sm["t1"])
        for tr, tm in zip(sr["tokens"], sm["tokens"]):
            for k in ("id", "tid", "t0", "t1", "text"):
                if k == "t1" and tr is sr["tokens"][-1] and (not last_t1 or tr["id"] < eot):
                    continue
                assert tr[k] == tm[k], k
            for k in ("p", "plog", "pt", "ptsum", "vlen"):
                assert tr[k] == tm[k] or (np.isnan(tr[k]) and np.isnan(tm[k])), k


def test_mel_is_bit_exact(ref_session, host_ctx, jfk):
    for n_threads in (1, 3, 4):
        assert ref_session.pcm_to_mel(jfk, n_threads) == 0 and host_ctx.pcm_to_mel(jfk, n_threads) == 0
        rmel, _ = ref_session.mel()
        mine = host_ctx.read_stage(wb.STAGE_HOST_MEL, np.float32).reshape(rmel.shape)
        assert np.array_equal(rmel, mine)


def test_mel_avx2_path_is_bit_exact_too(ref, model_bytes, jfk):
    """The leaf DFT has an AVX-512 and an AVX2 body (csrc/mel.cpp); a fresh process with WHISPER_B200_NO_AVX512=1 takes the latter."""
    import os, subprocess, sys, textwrap
    from conftest import ROOT, PKG, build_hostlogic
    code = textwrap.dedent(f"""
        import os, sys, hashlib
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {PKG!r})
        import numpy as np
        from oracle import ref_lib
        import whisper_b200 as wb
        os.environ["WHISPER_HOSTLOGIC_REF_LIB"] = ref_lib.ref_lib_path()
        lib = wb.load_library({build_hostlogic()!r}); wb.set_log_sink(lib, None)
        ctx = wb.Context(open(ref_lib.tiny_en_model_path(), "rb").read(), lib=lib)
        pcm = ref_lib.read_wav_f32(os.path.join({ROOT!r}, "tests", "golden", "jfk.wav"))
        assert ctx.pcm_to_mel(pcm, 2) == 0
        print("SHA", hashlib.sha1(ctx.read_stage(wb.STAGE_HOST_MEL, np.float32).tobytes()).hexdigest())
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, WHISPER_B200_NO_AVX512="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    import hashlib
    g = np.load(os.path.join(ROOT, "tests", "golden", "jfk_tiny_en.npz"))
    assert out.stdout.strip().split()[-1] == bytes(g["mel_sha1"]).hex()


def test_mel_on_noise_and_short_input(ref_session, host_ctx):
    rng = np.random.default_rng(7)
    for n in (16000, 17001, 48000 + 123):
        pcm = (rng.standard_normal(n) * 0.1).astype(np.float32)
        assert ref_session.pcm_to_mel(pcm, 2) == 0 and host_ctx.pcm_to_mel(pcm, 2) == 0
        rmel, _ = ref_session.mel()
        mine = host_ctx.read_stage(wb.STAGE_HOST_MEL, np.float32).reshape(rmel.shape)
        assert np.array_equal(rmel, mine)


def test_greedy_host_defaults(ref, ref_session, host_ctx, jfk):
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk, max_tokens=16)
    assert rc_r == rc_m == 0
    assert_same_result(rr, rm)


def test_greedy_full_sentence(ref, ref_session, host_ctx, jfk):
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk, max_tokens=0)
    assert rc_r == rc_m == 0
    assert_same_result(rr, rm)


def test_beam_search(ref, ref_session, host_ctx, jfk):
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk, max_tokens=0, strategy=wb.WHISPER_SAMPLING_BEAM_SEARCH)
    assert rc_r == rc_m == 0
    assert_same_result(rr, rm)


def test_initial_prompt_and_dynamic_audio_ctx(ref, ref_session, host_ctx, jfk):
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk, max_tokens=0, audio_ctx=11 * 50 + 128,
                              initial_prompt=b"A speech by the president.")
    assert rc_r == rc_m == 0
    assert_same_result(rr, rm)


def test_error_codes(ref, ref_session, host_ctx, jfk):
    # audio_ctx beyond the model's 1500 frames -> -5 (whisper.cpp:5098-5101)
    rc_r, rc_m, _, _ = both(ref, ref_session, host_ctx, jfk, audio_ctx=1501)
    assert rc_r == rc_m == -5
    # speed_up -> -1 (whisper.cpp:4973-4976)
    rc_r, rc_m, _, _ = both(ref, ref_session, host_ctx, jfk, speed_up=True)
    assert rc_r == rc_m == -1
    # < 1 s of audio -> 0 with no segments (whisper.cpp:5015-5021)
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk[:8000])
    assert rc_r == rc_m == 0 and rr["segments"] == rm["segments"] == []


def test_tokenizer_and_vocab(ref, ref_session, host_ctx):
    for text in (b" And so my fellow Americans", b"hello world", b" A speech by the president.", b""):
        assert ref_session.tokenize(text) == [int(x) for x in _tok(host_ctx, text)]
    for tid in (0, 50256, 50257, 50362, 50363, 50364, 51863):
        assert ref.whisper_token_to_str(ref_session.ctx, tid) == host_ctx.lib.whisper_token_to_str(host_ctx.ctx, tid)


def _tok(ctx, text, cap=1024):
    buf = (C.c_int32 * cap)()
    n = ctx.lib.whisper_tokenize(ctx.ctx, text, buf, cap)
    return list(buf[:max(n, 0)])


def test_threaded_full_batch_matches_single_calls(host_ctx, jfk):
    """whisper_b200_full_batch: one worker thread per chunk, device passes merged by the Batcher (csrc/batcher.cpp).  With the
    checker forward every slot is its own reference context, so per-chunk results must equal plain whisper_full() calls."""
    chunks = [jfk, jfk[:60000], np.roll(jfk, 16000), jfk[:100000], jfk[20000:]]
    p = wb.host_params(host_ctx.lib, max_tokens=0, n_threads=2)
    singles = []
    for c in chunks:
        assert host_ctx.full(p, c) == 0
        singles.append(host_ctx.result())
    import os
    os.environ["WHISPER_B200_MAX_WORKERS"] = "3"          # fewer workers than chunks: slots are reused
    try:
        assert host_ctx.full_batch(p, chunks) == 0
    finally:
        del os.environ["WHISPER_B200_MAX_WORKERS"]
    for i, want in enumerate(singles):
        got = host_ctx.chunk_result(i)
        assert ids_of(got) == ids_of(want)
        assert got["text"] == want["text"]


@pytest.mark.parametrize("runs", [1, 0])
def test_encoder_driver_thread_next_to_decoder_passes(hostlogic, model_bytes, host_ctx, jfk, monkeypatch, runs):
    """The Batcher's second driver: when the forward pass can run encoder passes on their own stream (Forward::encoder_concurrent,
    the CUDA forward outside profiling), encode requests are served by an encoder thread while the decoder driver keeps serving
    decode passes.  The checker forward takes that role with WHISPER_HOSTLOGIC_CONCURRENT_ENC=1 (every job only touches its own
    slot's reference context); more chunks than workers, so encodes of later chunks overlap with decodes of earlier ones.
    runs = 1: every greedy pass is one device-resident run (run_state.h; the checker carries 3 runs per step at depth 2, so runs queue,
    join and leave mid-flight); runs = 0: the per-token request path (WHISPER_HOSTLOGIC_RUNS=0)."""
    chunks = [jfk, jfk[:60000], np.roll(jfk, 16000), jfk[:100000], jfk[20000:], np.roll(jfk, 40000)[:90000], jfk[8000:150000]]
    p = wb.host_params(host_ctx.lib, max_tokens=0, n_threads=2)
    singles = []
    for c in chunks:
        assert host_ctx.full(p, c) == 0
        singles.append(host_ctx.result())
    monkeypatch.setenv("WHISPER_HOSTLOGIC_CONCURRENT_ENC", "1")
    monkeypatch.setenv("WHISPER_B200_MAX_WORKERS", "4")
    monkeypatch.setenv("WHISPER_HOSTLOGIC_RUNS", str(runs))
    ctx = wb.Context(model_bytes, lib=hostlogic)
    try:
        for _ in range(2):
            assert ctx.full_batch(p, chunks) == 0
            for i, want in enumerate(singles):
                got = ctx.chunk_result(i)
                assert ids_of(got) == ids_of(want), i
                assert got["text"] == want["text"]
    finally:
        ctx.close()


def test_sampling_from_the_distribution_on_the_device_side(hostlogic, model_bytes, ref, jfk):
    """Passes that draw from the distribution (best-of decoders at t > 0, beam search) send the uniform variates of the decoders'
    generators to the forward pass and get tokens back (Forward::can_sample_dist; here the checker inverts libstdc++'s
    std::discrete_distribution on the reference's logits).  Same generators, same order: results equal the reference's, including the
    generator state carried from one call to the next."""
    ctx = wb.Context(model_bytes, lib=hostlogic)
    fresh = ref_lib.RefSession(ref, model_bytes, use_gpu=False)
    try:
        for kw in (dict(max_tokens=24, temperature=0.4), dict(max_tokens=0, temperature=0.2, temperature_inc=0.4),
                   dict(max_tokens=0, strategy=wb.WHISPER_SAMPLING_BEAM_SEARCH), dict(max_tokens=12, temperature=1.0, **{"greedy.best_of": 3}),
                   dict(max_tokens=0, strategy=wb.WHISPER_SAMPLING_BEAM_SEARCH, temperature=0.2, **{"beam_search.beam_size": 3})):
            rc_r, rc_m, rr, rm = both(ref, fresh, ctx, jfk, **kw)
            assert rc_r == rc_m == 0, kw
            assert_same_result(rr, rm, last_t1=False)
    finally:
        ctx.close(); fresh.close()


@pytest.mark.parametrize("max_tokens", [0, 16])
def test_temperature_fallback_with_the_real_host_block(ref, ref_session, host_ctx, jfk, max_tokens):
    """The parameter block SpeechToText::transcribe really sets (src/speech_to_text.cpp:403-413: entropy_thold 2.8, temperature_inc
    left at 0.2) on a 30 s window: the t = 0 pass fails its entropy test, the loop falls back to t = 0.2 with best_of 5 decoders drawing
    from std::discrete_distribution (whisper.cpp:5187-5207, 5612-5668).  Same ids, same fallback counts as the reference."""
    audio = ref_lib.jfk30(jfk)
    c0, r0 = host_ctx.counters(), ref_session.counters()
    rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, audio, max_tokens=max_tokens, entropy_thold=2.8, temperature_inc=0.2)
    assert rc_r == rc_m == 0
    assert_same_result(rr, rm)
    c1, r1 = host_ctx.counters(), ref_session.counters()
    fails = (c1["n_fail_p"] - c0["n_fail_p"], c1["n_fail_h"] - c0["n_fail_h"])
    assert fails == (r1["n_fail_p"] - r0["n_fail_p"], r1["n_fail_h"] - r0["n_fail_h"])
    if max_tokens == 0:
        assert fails[0] >= 1 and fails[1] >= 1            # the case exists to exercise the loop: it must actually fall back


def test_per_token_host_path_without_runs(hostlogic, model_bytes, ref, ref_session, jfk, monkeypatch):
    """WHISPER_HOSTLOGIC_RUNS=0: the checker forward offers no runs, so whisper_full takes the per-token loop of csrc/full.cpp
    (sample on the host, one decoder request per token) — the path beam search and t > 0 always take.  Same result as the reference."""
    monkeypatch.setenv("WHISPER_HOSTLOGIC_RUNS", "0")
    monkeypatch.setenv("WHISPER_HOSTLOGIC_DIST", "0")     # ... and no sampling from the distribution on the "device" either: logits rows come back
    ctx = wb.Context(model_bytes, lib=hostlogic)
    fresh = ref_lib.RefSession(ref, model_bytes, use_gpu=False)      # (decoder 0's generator is seeded once per state and never again: both sides start fresh)
    try:
        for kw in (dict(max_tokens=0), dict(max_tokens=16), dict(max_tokens=0, initial_prompt=b"A speech by the president."),
                   dict(max_tokens=0, strategy=wb.WHISPER_SAMPLING_BEAM_SEARCH), dict(max_tokens=24, temperature=0.4)):
            rc_r, rc_m, rr, rm = both(ref, fresh, ctx, jfk, **kw)
            assert rc_r == rc_m == 0
            assert_same_result(rr, rm)
    finally:
        ctx.close(); fresh.close()


def test_run_state_machine_edge_cases(ref, ref_session, host_ctx, jfk):
    """The device-resident run (run_state.h, executed here by the checker forward on the reference's logits) against the reference's
    token loop where its integer rules bite: max_tokens cut-offs (whisper.cpp:5467-5490), a window that ends inside the audio
    (seek + seek_delta + 100 >= seek_end), several 30 s windows with context carried over (no_context = false: the prompt is prefilled
    in one pass, then the run starts), and a clip too short for a second window."""
    long = np.concatenate([jfk, jfk[:100000], jfk])               # 28.3 s ... one window; doubled below for two
    cases = [dict(max_tokens=1), dict(max_tokens=3), dict(max_tokens=0, single_segment=False), dict(max_tokens=0, duration_ms=5000),
             dict(max_tokens=0, offset_ms=3000, duration_ms=6000)]
    for kw in cases:
        rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, jfk, **kw)
        assert rc_r == rc_m == 0, kw
        assert_same_result(rr, rm, last_t1=False)
    two = np.concatenate([long, long])
    for kw in (dict(max_tokens=0, no_context=False, single_segment=False), dict(max_tokens=0, no_context=False)):
        rc_r, rc_m, rr, rm = both(ref, ref_session, host_ctx, two, **kw)
        assert rc_r == rc_m == 0, kw
        assert_same_result(rr, rm, last_t1=False)
