"""ctypes binding of libwhisper_b200.so — the C ABI declared in include/whisper_b200.h.

This is what a host written in Python binds; it mirrors the GDExtension host's use of the library
(/root/reference/src/speech_to_text.cpp:331-447): `Context` = one whisper_context, `host_params` = the parameter block
SpeechToText::transcribe builds, `transcribe` = its result marshalling.  No CPU fallback: if the library or a B200 is
missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwhisper_b200.so")

WHISPER_SAMPLE_RATE = 16000
WHISPER_SAMPLING_GREEDY = 0
WHISPER_SAMPLING_BEAM_SEARCH = 1


class WhisperContextParams(C.Structure):               # whisper.h:87-89
    _fields_ = [("use_gpu", C.c_bool)]


class WhisperTokenData(C.Structure):                   # whisper.h:91-106
    _fields_ = [
        ("id", C.c_int32), ("tid", C.c_int32),
        ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float), ("ptsum", C.c_float),
        ("t0", C.c_int64), ("t1", C.c_int64),
        ("vlen", C.c_float),
    ]


class _Greedy(C.Structure):
    _fields_ = [("best_of", C.c_int)]


class _Beam(C.Structure):
    _fields_ = [("beam_size", C.c_int), ("patience", C.c_float)]


class WhisperFullParams(C.Structure):                  # whisper.h:433-526
    _fields_ = [
        ("strategy", C.c_int),
        ("n_threads", C.c_int), ("n_max_text_ctx", C.c_int), ("offset_ms", C.c_int), ("duration_ms", C.c_int),
        ("translate", C.c_bool), ("no_context", C.c_bool), ("no_timestamps", C.c_bool), ("single_segment", C.c_bool),
        ("print_special", C.c_bool), ("print_progress", C.c_bool), ("print_realtime", C.c_bool),
        ("print_timestamps", C.c_bool),
        ("token_timestamps", C.c_bool), ("thold_pt", C.c_float), ("thold_ptsum", C.c_float), ("max_len", C.c_int),
        ("split_on_word", C.c_bool), ("max_tokens", C.c_int),
        ("speed_up", C.c_bool), ("debug_mode", C.c_bool), ("audio_ctx", C.c_int),
        ("tdrz_enable", C.c_bool),
        ("initial_prompt", C.c_char_p), ("prompt_tokens", C.POINTER(C.c_int32)), ("prompt_n_tokens", C.c_int),
        ("language", C.c_char_p), ("detect_language", C.c_bool),
        ("suppress_blank", C.c_bool), ("suppress_non_speech_tokens", C.c_bool),
        ("temperature", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
        ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float),
        ("no_speech_thold", C.c_float),
        ("greedy", _Greedy),
        ("beam_search", _Beam),
        ("new_segment_callback", C.c_void_p), ("new_segment_callback_user_data", C.c_void_p),
        ("progress_callback", C.c_void_p), ("progress_callback_user_data", C.c_void_p),
        ("encoder_begin_callback", C.c_void_p), ("encoder_begin_callback_user_data", C.c_void_p),
        ("abort_callback", C.c_void_p), ("abort_callback_user_data", C.c_void_p),
        ("logits_filter_callback", C.c_void_p), ("logits_filter_callback_user_data", C.c_void_p),
        ("grammar_rules", C.c_void_p), ("n_grammar_rules", C.c_size_t), ("i_start_rule", C.c_size_t),
        ("grammar_penalty", C.c_float),
    ]


assert C.sizeof(WhisperFullParams) == 256 and C.sizeof(WhisperTokenData) == 48

LOG_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)

STAGE_MEL_WINDOW, STAGE_EMBD_CONV, STAGE_EMBD_ENC, STAGE_CROSS_K, STAGE_CROSS_V, STAGE_SELF_K, STAGE_SELF_V, STAGE_HOST_MEL = range(8)
STAGE_DEVICE_MEL = 9

# every symbol include/whisper_b200.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = """
whisper_init_from_buffer_with_params whisper_free whisper_print_system_info whisper_full_default_params whisper_full
whisper_full_n_segments whisper_full_n_tokens whisper_full_get_segment_text whisper_full_get_token_text
whisper_full_get_token_data whisper_log_set whisper_context_default_params whisper_full_get_segment_t0
whisper_full_get_segment_t1 whisper_full_get_token_id whisper_full_lang_id whisper_pcm_to_mel whisper_set_mel
whisper_encode whisper_decode whisper_get_logits whisper_tokenize whisper_lang_max_id whisper_lang_id whisper_lang_str
whisper_lang_auto_detect whisper_n_len whisper_n_vocab whisper_n_text_ctx whisper_n_audio_ctx whisper_is_multilingual
whisper_model_n_audio_state whisper_model_n_audio_head whisper_model_n_audio_layer whisper_model_n_text_layer
whisper_token_to_str whisper_token_eot whisper_token_sot whisper_token_solm whisper_token_prev whisper_token_nosp
whisper_token_not whisper_token_beg whisper_token_lang whisper_token_translate whisper_token_transcribe
whisper_print_timings whisper_reset_timings whisper_b200_full_batch whisper_b200_chunk_n_segments
whisper_b200_chunk_n_tokens whisper_b200_chunk_segment_text whisper_b200_chunk_token_data whisper_b200_chunk_token_ids whisper_b200_init_multi whisper_b200_n_devices whisper_b200_host_alloc whisper_b200_host_free whisper_b200_dequantize whisper_b200_set_device
whisper_b200_counters whisper_b200_timings_us whisper_b200_read_stage whisper_b200_set_gemm_engine whisper_b200_gemm_f16 whisper_b200_gemm_enc_probe whisper_b200_attn_enc_probe whisper_b200_f16_tables
whisper_b200_gpu_times whisper_b200_gpu_busy_ms whisper_b200_gpu_mel_ms whisper_b200_high_pass_filter whisper_b200_vad_simple whisper_b200_set_profiling whisper_b200_profile
""".split()


def build(verbose: bool = False) -> str:
    """Compile libwhisper_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libwhisper_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


_LIB = None
_KEEP = []


def load_library(path: str | None = None) -> C.CDLL:
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not built: run `make -C godot-whisper_b200` (there is no CPU fallback)")
    lib = C.CDLL(p, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    sig = {
        "whisper_init_from_buffer_with_params": ([vp, C.c_size_t, WhisperContextParams], vp),
        "whisper_free": ([vp], None),
        "whisper_print_system_info": ([], C.c_char_p),
        "whisper_full_default_params": ([C.c_int], WhisperFullParams),
        "whisper_full": ([vp, WhisperFullParams, fp, C.c_int], C.c_int),
        "whisper_full_n_segments": ([vp], C.c_int),
        "whisper_full_n_tokens": ([vp, C.c_int], C.c_int),
        "whisper_full_get_segment_text": ([vp, C.c_int], C.c_char_p),
        "whisper_full_get_segment_t0": ([vp, C.c_int], C.c_int64),
        "whisper_full_get_segment_t1": ([vp, C.c_int], C.c_int64),
        "whisper_full_get_token_text": ([vp, C.c_int, C.c_int], C.c_char_p),
        "whisper_full_get_token_data": ([vp, C.c_int, C.c_int], WhisperTokenData),
        "whisper_full_get_token_id": ([vp, C.c_int, C.c_int], C.c_int32),
        "whisper_log_set": ([LOG_CALLBACK, vp], None),
        "whisper_pcm_to_mel": ([vp, fp, C.c_int, C.c_int], C.c_int),
        "whisper_set_mel": ([vp, fp, C.c_int, C.c_int], C.c_int),
        "whisper_encode": ([vp, C.c_int, C.c_int], C.c_int),
        "whisper_decode": ([vp, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int], C.c_int),
        "whisper_get_logits": ([vp], fp),
        "whisper_tokenize": ([vp, C.c_char_p, C.POINTER(C.c_int32), C.c_int], C.c_int),
        "whisper_n_vocab": ([vp], C.c_int),
        "whisper_n_len": ([vp], C.c_int),
        "whisper_n_audio_ctx": ([vp], C.c_int),
        "whisper_n_text_ctx": ([vp], C.c_int),
        "whisper_model_n_audio_state": ([vp], C.c_int),
        "whisper_model_n_text_layer": ([vp], C.c_int),
        "whisper_token_to_str": ([vp, C.c_int32], C.c_char_p),
        "whisper_print_timings": ([vp], None),
        "whisper_reset_timings": ([vp], None),
        "whisper_b200_full_batch": ([vp, WhisperFullParams, C.POINTER(fp), ip, C.c_int], C.c_int),
        "whisper_b200_chunk_n_segments": ([vp, C.c_int], C.c_int),
        "whisper_b200_chunk_n_tokens": ([vp, C.c_int, C.c_int], C.c_int),
        "whisper_b200_chunk_segment_text": ([vp, C.c_int, C.c_int], C.c_char_p),
        "whisper_b200_chunk_token_data": ([vp, C.c_int, C.c_int, C.c_int], WhisperTokenData),
        "whisper_b200_chunk_token_ids": ([vp, C.c_int, C.POINTER(C.c_int32), C.c_int], C.c_int),
        "whisper_b200_init_multi": ([vp, C.c_size_t, WhisperContextParams, C.POINTER(C.c_int), C.c_int], vp),
        "whisper_b200_n_devices": ([vp], C.c_int),
        "whisper_b200_host_alloc": ([C.c_size_t], vp),
        "whisper_b200_host_free": ([vp], None),
        "whisper_b200_dequantize": ([C.c_int, vp, C.c_longlong, fp], C.c_int),
        "whisper_b200_set_device": ([C.c_int], None),
        "whisper_b200_counters": ([vp, C.POINTER(C.c_int64)], None),
        "whisper_b200_timings_us": ([vp, C.POINTER(C.c_int64)], None),
        "whisper_b200_read_stage": ([vp, C.c_int, vp, C.c_longlong], C.c_longlong),
        "whisper_b200_set_gemm_engine": ([vp, C.c_int], None),
        "whisper_b200_gpu_times": ([vp, C.POINTER(C.c_double)], None),
        "whisper_b200_gpu_busy_ms": ([vp], C.c_double),
        "whisper_b200_gpu_mel_ms": ([vp], C.c_double),
        "whisper_b200_high_pass_filter": ([fp, C.c_int, C.c_float, C.c_float], None),
        "whisper_b200_vad_simple": ([fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float], C.c_int),
        "whisper_b200_set_profiling": ([vp, C.c_int], None),
        "whisper_b200_profile": ([vp, C.POINTER(C.c_double)], None),
        "whisper_b200_f16_tables": ([C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)], None),
        "whisper_b200_gemm_f16": ([vp, vp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "whisper_b200_gemm_enc_probe": ([vp, vp, fp, fp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp], C.c_int),
        "whisper_b200_attn_enc_probe": ([vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp], C.c_int),
    }
    for name, (args, res) in sig.items():
        if path is not None and not hasattr(lib, name):
            continue    # an explicitly named library (the CPU-only host-logic test build) may lack the CUDA-only entries
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    for name in ("eot", "sot", "solm", "prev", "nosp", "not", "beg", "translate", "transcribe"):
        fn = getattr(lib, f"whisper_token_{name}")
        fn.argtypes = [vp]
        fn.restype = C.c_int32
    lang = {"whisper_lang_auto_detect": ([vp, C.c_int, C.c_int, fp], C.c_int), "whisper_token_lang": ([vp, C.c_int], C.c_int32),
            "whisper_is_multilingual": ([vp], C.c_int), "whisper_full_lang_id": ([vp], C.c_int), "whisper_lang_max_id": ([], C.c_int),
            "whisper_lang_id": ([C.c_char_p], C.c_int), "whisper_lang_str": ([C.c_int], C.c_char_p)}
    for name, (args, res) in lang.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if path is None:
        _LIB = lib
    return lib


def set_log_sink(lib: C.CDLL, sink: list | None) -> None:
    """Route library log lines into `sink` (None drops them)."""
    def _cb(level, text, _ud):
        if sink is not None:
            sink.append((level, text.decode("utf-8", "replace")))
    cb = LOG_CALLBACK(_cb)
    _KEEP.append(cb)
    lib.whisper_log_set(cb, None)


def read_wav_f32(path: str) -> np.ndarray:
    """16-bit PCM WAV at 16 kHz -> mono float32 in [-1, 1)."""
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getframerate() == WHISPER_SAMPLE_RATE
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        if w.getnchannels() == 2:
            raw = raw.reshape(-1, 2)
            return ((raw[:, 0].astype(np.float32) + raw[:, 1].astype(np.float32)) / 65536.0).astype(np.float32)
        return (raw.astype(np.float32) / 32768.0).astype(np.float32)


def host_params(lib: C.CDLL, *, strategy: int = WHISPER_SAMPLING_GREEDY, language: bytes = b"en", audio_ctx: int = 0,
                max_tokens: int = 16, entropy_thold: float = 2.8, initial_prompt: bytes = b"",
                n_threads: int | None = None, **overrides) -> WhisperFullParams:
    """The parameter block SpeechToText::transcribe builds (src/speech_to_text.cpp:403-413) with the project-setting
    defaults of src/register_types.cpp:64-69."""
    p = lib.whisper_full_default_params(strategy)
    p.language = language
    p.audio_ctx = audio_ctx
    p.speed_up = False
    p.split_on_word = True
    p.token_timestamps = True
    p.suppress_non_speech_tokens = True
    p.single_segment = True
    p.max_tokens = max_tokens
    p.entropy_thold = entropy_thold
    p.initial_prompt = initial_prompt
    if n_threads is not None:
        p.n_threads = n_threads
    for k, v in overrides.items():
        if "." in k:
            a, b = k.split(".")
            setattr(getattr(p, a), b, v)
        else:
            setattr(p, k, v)
    return p


class Context:
    """One whisper_context on one B200 (model resident in HBM)."""

    def __init__(self, model_bytes: bytes, device: int | None = None, lib: C.CDLL | None = None, devices: list | None = None):
        self.lib = lib or load_library()
        buf = (C.c_char * len(model_bytes)).from_buffer_copy(model_bytes)
        if devices is not None:
            # one context over several GPUs of the box (whisper_b200_init_multi): full_batch deals the chunks over them
            arr = (C.c_int * len(devices))(*devices)
            self.ctx = self.lib.whisper_b200_init_multi(C.cast(buf, C.c_void_p), len(model_bytes), WhisperContextParams(True), arr, len(devices))
        else:
            if device is not None:
                self.lib.whisper_b200_set_device(device)
            self.ctx = self.lib.whisper_init_from_buffer_with_params(C.cast(buf, C.c_void_p), len(model_bytes),
                                                                     WhisperContextParams(True))
        del buf   # the contract says the buffer may die right after init (src/speech_to_text.cpp:342-346)
        if not self.ctx:
            raise RuntimeError("whisper_init_from_buffer_with_params returned NULL (no B200 / bad model); no CPU fallback")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.whisper_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- whisper_full + result marshalling (src/speech_to_text.cpp:419-447) --
    def full(self, params: WhisperFullParams, pcm: np.ndarray) -> int:
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.whisper_full(self.ctx, params, pcm.ctypes.data_as(C.POINTER(C.c_float)), int(pcm.size))

    def result(self) -> dict:
        lib, ctx = self.lib, self.ctx
        segs = []
        for i in range(lib.whisper_full_n_segments(ctx)):
            toks = []
            for j in range(lib.whisper_full_n_tokens(ctx, i)):
                d = lib.whisper_full_get_token_data(ctx, i, j)
                toks.append(dict(id=d.id, tid=d.tid, p=d.p, plog=d.plog, pt=d.pt, ptsum=d.ptsum, t0=d.t0, t1=d.t1,
                                 vlen=d.vlen, text=lib.whisper_full_get_token_text(ctx, i, j)))
            segs.append(dict(text=lib.whisper_full_get_segment_text(ctx, i),
                             t0=lib.whisper_full_get_segment_t0(ctx, i), t1=lib.whisper_full_get_segment_t1(ctx, i),
                             tokens=toks))
        return dict(segments=segs, text=b"".join(s["text"] for s in segs))

    def transcribe(self, pcm: np.ndarray, initial_prompt: bytes = b"", audio_ctx: int = 0, **kw) -> list:
        """SpeechToText::transcribe: returns the token dictionaries the GDExtension hands to GDScript."""
        rc = self.full(host_params(self.lib, initial_prompt=initial_prompt, audio_ctx=audio_ctx, **kw), pcm)
        if rc != 0:
            return []
        out = []
        for seg in self.result()["segments"]:
            for t in seg["tokens"]:
                out.append(dict(id=t["id"], text=t["text"], p=t["p"], plog=t["plog"], pt=t["pt"], ptsum=t["ptsum"],
                                t0=t["t0"], t1=t["t1"], vlen=t["vlen"]))
        return out

    def full_batch(self, params: WhisperFullParams, chunks: list) -> int:
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in chunks]
        fp = C.POINTER(C.c_float)
        ptrs = (fp * len(arrs))(*[a.ctypes.data_as(fp) for a in arrs])
        lens = (C.c_int * len(arrs))(*[int(a.size) for a in arrs])
        return self.lib.whisper_b200_full_batch(self.ctx, params, ptrs, lens, len(arrs))

    def chunk_text(self, c: int) -> bytes:
        lib, ctx = self.lib, self.ctx
        return b"".join(lib.whisper_b200_chunk_segment_text(ctx, c, i) for i in range(lib.whisper_b200_chunk_n_segments(ctx, c)))

    def chunk_ids(self, c: int) -> list:
        buf = (C.c_int32 * 1024)()
        n = self.lib.whisper_b200_chunk_token_ids(self.ctx, c, buf, 1024)
        return list(buf[:max(0, min(n, 1024))])

    def chunk_result(self, c: int) -> dict:
        lib, ctx = self.lib, self.ctx
        segs = []
        for i in range(lib.whisper_b200_chunk_n_segments(ctx, c)):
            toks = []
            for j in range(lib.whisper_b200_chunk_n_tokens(ctx, c, i)):
                d = lib.whisper_b200_chunk_token_data(ctx, c, i, j)
                toks.append(dict(id=d.id, tid=d.tid, p=d.p, plog=d.plog, pt=d.pt, ptsum=d.ptsum, t0=d.t0, t1=d.t1, vlen=d.vlen))
            segs.append(dict(text=lib.whisper_b200_chunk_segment_text(ctx, c, i), tokens=toks))
        return dict(segments=segs, text=b"".join(s["text"] for s in segs))

    # -- stage API --
    def pcm_to_mel(self, pcm: np.ndarray, n_threads: int = 1) -> int:
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.whisper_pcm_to_mel(self.ctx, pcm.ctypes.data_as(C.POINTER(C.c_float)), int(pcm.size), n_threads)

    def set_mel(self, mel: np.ndarray) -> int:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        n_mel, n_len = mel.shape
        return self.lib.whisper_set_mel(self.ctx, mel.ctypes.data_as(C.POINTER(C.c_float)), n_len, n_mel)

    def encode(self, offset: int = 0) -> int:
        return self.lib.whisper_encode(self.ctx, offset, 1)

    def decode(self, tokens, n_past: int) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        rc = self.lib.whisper_decode(self.ctx, toks.ctypes.data_as(C.POINTER(C.c_int32)), int(toks.size), n_past, 1)
        if rc != 0:
            raise RuntimeError(f"whisper_decode -> {rc}")
        nv = self.lib.whisper_n_vocab(self.ctx)
        rows = np.ctypeslib.as_array(self.lib.whisper_get_logits(self.ctx), shape=(int(toks.size), nv))
        return np.array(rows[-1], dtype=np.float32)

    def read_stage(self, what: int, dtype) -> np.ndarray:
        nbytes = self.lib.whisper_b200_read_stage(self.ctx, what, None, 0)
        if nbytes < 0:
            raise RuntimeError(f"stage {what} not available")
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self.lib.whisper_b200_read_stage(self.ctx, what, out.ctypes.data_as(C.c_void_p), nbytes)
        return out

    def set_gemm_engine(self, engine: int) -> None:
        self.lib.whisper_b200_set_gemm_engine(self.ctx, engine)

    def counters(self) -> dict:
        out = (C.c_int64 * 8)()
        self.lib.whisper_b200_counters(self.ctx, out)
        return dict(zip(("n_sample", "n_encode", "n_decode", "n_batchd", "n_prompt", "n_fail_p", "n_fail_h", "launches"), out))

    def gpu_times(self) -> dict:
        out = (C.c_double * 8)()
        self.lib.whisper_b200_gpu_times(self.ctx, out)
        return dict(encode_ms=out[0], decode_ms=out[1], n_encode=int(out[2]), n_decode=int(out[3]), h2d_bytes=out[4], d2h_bytes=out[5],
                    step_launches=int(out[6]), step_bytes=out[7], mel_ms=float(self.lib.whisper_b200_gpu_mel_ms(self.ctx)))

    def gpu_busy_ms(self) -> float:
        return float(self.lib.whisper_b200_gpu_busy_ms(self.ctx))

    PROF_KINDS = ("gemm_enc", "gemm_attn", "softmax", "layernorm", "skinny", "dec_attn", "misc", "gemm_dec", "decode_step")

    def set_profiling(self, on: bool) -> None:
        self.lib.whisper_b200_set_profiling(self.ctx, 1 if on else 0)

    def profile(self) -> dict:
        out = (C.c_double * 36)()
        self.lib.whisper_b200_profile(self.ctx, out)
        return {k: dict(launches=int(out[4 * i]), ms=out[4 * i + 1], flop=out[4 * i + 2], bytes=out[4 * i + 3])
                for i, k in enumerate(self.PROF_KINDS)}

    def timings_us(self) -> dict:
        out = (C.c_int64 * 6)()
        self.lib.whisper_b200_timings_us(self.ctx, out)
        return dict(zip(("mel", "sample", "encode", "decode", "batchd", "prompt"), out))


def gemm_f16(A: np.ndarray, B: np.ndarray, engine: int = 0, iters: int = 0):
    """C[n][m] = sum_k A[m][k] * B[n][k] on the GPU (A, B float16).  Returns (C float32 [N][M], ms_per_iter)."""
    lib = load_library()
    A = np.ascontiguousarray(A, dtype=np.float16)
    B = np.ascontiguousarray(B, dtype=np.float16)
    M, K = A.shape
    N, K2 = B.shape
    assert K == K2
    out = np.empty((N, M), dtype=np.float32)
    ms = C.c_float(0.0)
    rc = lib.whisper_b200_gemm_f16(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.POINTER(C.c_float)), M, N, K, engine, iters, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"whisper_b200_gemm_f16 -> {rc}")
    return out, ms.value


def gemm_enc_probe(act: np.ndarray, wgt: np.ndarray, mode: int, bias: np.ndarray | None = None, res: np.ndarray | None = None, iters: int = 0):
    """The encoder GEMM with the TMA-store epilogue on host buffers (whisper_b200_gemm_enc_probe).  act f16 [N][K], wgt f16 [M][K]."""
    lib = load_library()
    act = np.ascontiguousarray(act, dtype=np.float16)
    wgt = np.ascontiguousarray(wgt, dtype=np.float16)
    N, K = act.shape
    M, K2 = wgt.shape
    assert K == K2
    ldt = (N + 7) & ~7
    out = np.zeros((N, M), np.float32) if mode == 3 else np.zeros((M, ldt), np.float16) if mode == 2 else np.zeros((N, M), np.float16)
    fp = C.POINTER(C.c_float)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    r = None if res is None else np.ascontiguousarray(res, dtype=np.float32)
    ms = C.c_float(0.0)
    rc = lib.whisper_b200_gemm_enc_probe(act.ctypes.data_as(C.c_void_p), wgt.ctypes.data_as(C.c_void_p),
                                         None if b is None else b.ctypes.data_as(fp), None if r is None else r.ctypes.data_as(fp),
                                         out.ctypes.data_as(C.c_void_p), N, M, K, mode, iters, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"whisper_b200_gemm_enc_probe -> {rc}")
    return out, ms.value


def vad_simple(pcm: np.ndarray, sample_rate: int = 16000, last_ms: int = 500, vad_thold: float = 0.3, freq_thold: float = 200.0, lib: C.CDLL | None = None):
    """whisper_b200_vad_simple on a copy of `pcm`: returns (decision, the filtered window)."""
    lib = lib or load_library()
    a = np.array(pcm, dtype=np.float32, copy=True)
    rc = lib.whisper_b200_vad_simple(a.ctypes.data_as(C.POINTER(C.c_float)), a.size, sample_rate, last_ms, vad_thold, freq_thold)
    return int(rc), a


def attn_enc_probe(q: np.ndarray, k: np.ndarray, vt: np.ndarray, n_head: int, variant: int = -1, iters: int = 0):
    """Fused encoder attention on host buffers (whisper_b200_attn_enc_probe).  q, k f16 [B][T][d]; vt f16 [B][d][Tp], Tp = T rounded up to 8.
    Returns (out f16 [B][T][d], ms_per_iter)."""
    lib = load_library()
    q = np.ascontiguousarray(q, dtype=np.float16)
    k = np.ascontiguousarray(k, dtype=np.float16)
    vt = np.ascontiguousarray(vt, dtype=np.float16)
    B, T, d = q.shape
    assert k.shape == q.shape and vt.shape == (B, d, (T + 7) & ~7) and d == 64 * n_head
    out = np.zeros((B, T, d), np.float16)
    ms = C.c_float(0.0)
    vp = C.c_void_p
    rc = lib.whisper_b200_attn_enc_probe(q.ctypes.data_as(vp), k.ctypes.data_as(vp), vt.ctypes.data_as(vp), out.ctypes.data_as(vp), B, T, d, n_head,
                                         variant, iters, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"whisper_b200_attn_enc_probe -> {rc}")
    return out, ms.value


def pinned_copy(arr: np.ndarray, lib: C.CDLL | None = None) -> np.ndarray:
    """A float32 copy of `arr` in page-locked host memory (whisper_b200_host_alloc): the library uploads such PCM from where it lies.
    The memory is never freed (bench / test inputs live as long as the process)."""
    lib = lib or load_library()
    a = np.ascontiguousarray(arr, dtype=np.float32)
    p = lib.whisper_b200_host_alloc(a.nbytes)
    if not p:
        return a
    out = np.ctypeslib.as_array((C.c_float * a.size).from_address(p)).reshape(a.shape)
    out[...] = a
    return out
