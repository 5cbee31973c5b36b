"""TEST INFRASTRUCTURE — ctypes binding of the compiled reference (oracle/_ref/libwhisper_ref_*.so).

The library is the UNMODIFIED whisper.cpp v1.5.4 CPU path of /root/reference (ggml, BLAS off) built by
oracle/Makefile, plus the probe_* accessors of oracle/ref_probe.cpp.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Struct layouts follow /root/reference/thirdparty/whisper.cpp/whisper.h:87-106 (context params, token data) and
whisper.h:433-526 (whisper_full_params, 256 bytes).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

WHISPER_SAMPLING_GREEDY = 0
WHISPER_SAMPLING_BEAM_SEARCH = 1


class WhisperContextParams(C.Structure):
    _fields_ = [("use_gpu", C.c_bool)]


class WhisperTokenData(C.Structure):
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


class WhisperFullParams(C.Structure):
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


assert C.sizeof(WhisperFullParams) == 256, C.sizeof(WhisperFullParams)
assert C.sizeof(WhisperTokenData) == 48

LOG_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)


def cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
    except OSError:
        return False
    flags = set()
    for line in txt.splitlines():
        if line.startswith("flags"):
            flags = set(line.split(":", 1)[1].split())
            break
    return {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags


def ref_lib_path(variant: str | None = None) -> str:
    if variant is None:
        variant = os.environ.get("WHISPER_REF_VARIANT") or ("v4" if cpu_has_avx512() else "v3")
    return os.path.join(REF_DIR, f"libwhisper_ref_{variant}.so")


def available() -> bool:
    return os.path.exists(ref_lib_path())


def bind_whisper_api(lib: C.CDLL) -> None:
    """Declare argtypes/restype of the whisper.h functions shared by the reference and by the drop-in."""
    vp = C.c_void_p
    lib.whisper_init_from_buffer_with_params.argtypes = [vp, C.c_size_t, WhisperContextParams]
    lib.whisper_init_from_buffer_with_params.restype = vp
    lib.whisper_free.argtypes = [vp]
    lib.whisper_free.restype = None
    lib.whisper_print_system_info.argtypes = []
    lib.whisper_print_system_info.restype = C.c_char_p
    lib.whisper_full_default_params.argtypes = [C.c_int]
    lib.whisper_full_default_params.restype = WhisperFullParams
    lib.whisper_full.argtypes = [vp, WhisperFullParams, C.POINTER(C.c_float), C.c_int]
    lib.whisper_full.restype = C.c_int
    lib.whisper_full_n_segments.argtypes = [vp]
    lib.whisper_full_n_segments.restype = C.c_int
    lib.whisper_full_n_tokens.argtypes = [vp, C.c_int]
    lib.whisper_full_n_tokens.restype = C.c_int
    lib.whisper_full_get_segment_text.argtypes = [vp, C.c_int]
    lib.whisper_full_get_segment_text.restype = C.c_char_p
    lib.whisper_full_get_segment_t0.argtypes = [vp, C.c_int]
    lib.whisper_full_get_segment_t0.restype = C.c_int64
    lib.whisper_full_get_segment_t1.argtypes = [vp, C.c_int]
    lib.whisper_full_get_segment_t1.restype = C.c_int64
    lib.whisper_full_get_token_text.argtypes = [vp, C.c_int, C.c_int]
    lib.whisper_full_get_token_text.restype = C.c_char_p
    lib.whisper_full_get_token_data.argtypes = [vp, C.c_int, C.c_int]
    lib.whisper_full_get_token_data.restype = WhisperTokenData
    lib.whisper_log_set.argtypes = [LOG_CALLBACK, vp]
    lib.whisper_log_set.restype = None
    # stage API (whisper.h:223-290,364) — the per-stage parity hooks
    lib.whisper_pcm_to_mel.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_int]
    lib.whisper_pcm_to_mel.restype = C.c_int
    lib.whisper_set_mel.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_int]
    lib.whisper_set_mel.restype = C.c_int
    lib.whisper_encode.argtypes = [vp, C.c_int, C.c_int]
    lib.whisper_encode.restype = C.c_int
    lib.whisper_decode.argtypes = [vp, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int]
    lib.whisper_decode.restype = C.c_int
    lib.whisper_get_logits.argtypes = [vp]
    lib.whisper_get_logits.restype = C.POINTER(C.c_float)
    lib.whisper_tokenize.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32), C.c_int]
    lib.whisper_tokenize.restype = C.c_int
    lib.whisper_n_vocab.argtypes = [vp]
    lib.whisper_n_vocab.restype = C.c_int
    lib.whisper_n_len.argtypes = [vp]
    lib.whisper_n_len.restype = C.c_int
    lib.whisper_token_to_str.argtypes = [vp, C.c_int32]
    lib.whisper_token_to_str.restype = C.c_char_p
    for name in ("eot", "sot", "solm", "prev", "nosp", "not", "beg", "translate", "transcribe"):
        fn = getattr(lib, f"whisper_token_{name}")
        fn.argtypes = [vp]
        fn.restype = C.c_int32


def bind_lang_api(lib: C.CDLL) -> None:
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    for name, (args, res) in {"whisper_lang_auto_detect": ([vp, C.c_int, C.c_int, fp], C.c_int), "whisper_token_lang": ([vp, C.c_int], C.c_int32),
                              "whisper_is_multilingual": ([vp], C.c_int), "whisper_full_lang_id": ([vp], C.c_int),
                              "whisper_lang_max_id": ([], C.c_int)}.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res


def bind_probe_api(lib: C.CDLL) -> None:
    vp = C.c_void_p
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    u16p = C.POINTER(C.c_uint16)
    lib.probe_mel.argtypes = [vp, fp, C.c_int, ip]
    lib.probe_mel.restype = C.c_int
    lib.probe_embd_conv.argtypes = [vp, fp, C.c_int, ip, ip]
    lib.probe_embd_conv.restype = C.c_int
    lib.probe_embd_enc.argtypes = [vp, fp, C.c_int, ip, ip]
    lib.probe_embd_enc.restype = C.c_int
    lib.probe_kv_cross.argtypes = [vp, u16p, u16p, C.c_int]
    lib.probe_kv_cross.restype = C.c_int
    lib.probe_kv_self.argtypes = [vp, u16p, u16p, C.c_int]
    lib.probe_kv_self.restype = C.c_int
    lib.probe_set_audio_ctx.argtypes = [vp, C.c_int]
    lib.probe_set_audio_ctx.restype = None
    lib.probe_kv_self_clear.argtypes = [vp]
    lib.probe_kv_self_clear.restype = None
    lib.probe_kv_self_seq_cp.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.probe_kv_self_seq_cp.restype = None
    lib.probe_kv_self_seq_rm.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.probe_kv_self_seq_rm.restype = None
    lib.probe_kv_self_cells.argtypes = [vp, ip, C.POINTER(C.c_uint), C.c_int, ip]
    lib.probe_kv_self_cells.restype = C.c_int
    lib.probe_kv_self_set_cells.argtypes = [vp, ip, C.POINTER(C.c_uint), C.c_int, C.c_int]
    lib.probe_kv_self_set_cells.restype = None
    lib.probe_decode_batch.argtypes = [vp, ip, ip, ip, C.POINTER(C.c_int8), C.c_int, C.c_int]
    lib.probe_decode_batch.restype = C.c_int
    lib.probe_process_logits.argtypes = [vp, WhisperFullParams, fp, ip, C.c_int, C.c_int, C.c_int, C.c_float,
                                         fp, fp, fp]
    lib.probe_process_logits.restype = None
    lib.probe_sample_token.argtypes = [vp, C.c_int]
    lib.probe_sample_token.restype = WhisperTokenData
    lib.probe_sample_token_topk.argtypes = [vp, C.c_int, C.POINTER(WhisperTokenData)]
    lib.probe_sample_token_topk.restype = C.c_int
    lib.probe_seed_rng.argtypes = [vp, C.c_int, C.c_uint]
    lib.probe_seed_rng.restype = None
    lib.probe_counters.argtypes = [vp, ip]
    lib.probe_counters.restype = None
    lib.probe_timings_us.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.probe_timings_us.restype = None
    lib.probe_sizeof_full_params.restype = C.c_int
    lib.probe_sizeof_token_data.restype = C.c_int
    lib.probe_f16_tables.argtypes = [u16p, u16p]
    lib.probe_f16_tables.restype = None


_LIB = None
_LOG_KEEPALIVE = []


def load(quiet: bool = True) -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = ref_lib_path()
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference is mounted")
        lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        bind_whisper_api(lib)
        bind_probe_api(lib)
        bind_lang_api(lib)
        _LIB = lib
    if quiet:
        set_quiet(_LIB)
    return _LIB


def set_quiet(lib: C.CDLL, sink: list | None = None) -> None:
    """Route the library's log lines into `sink` (or drop them)."""
    def _cb(level, text, _ud):
        if sink is not None:
            sink.append((level, text.decode("utf-8", "replace")))
    cb = LOG_CALLBACK(_cb)
    _LOG_KEEPALIVE.append(cb)
    lib.whisper_log_set(cb, None)


# ---- convenience wrappers used by tests and bench ---------------------------------------------------------------------

def read_wav_f32(path: str) -> np.ndarray:
    """16-bit PCM mono/stereo WAV -> float32 mono in [-1, 1) (same scaling as examples/common.cpp:668-676)."""
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getframerate() == 16000
        n = w.getnframes()
        raw = np.frombuffer(w.readframes(n), dtype="<i2")
        if w.getnchannels() == 2:
            raw = raw.reshape(-1, 2)
            return ((raw[:, 0].astype(np.float32) + raw[:, 1].astype(np.float32)) / 65536.0).astype(np.float32)
        return (raw.astype(np.float32) / 32768.0).astype(np.float32)


def host_params(lib: C.CDLL, *, strategy: int = WHISPER_SAMPLING_GREEDY, language: bytes = b"en", audio_ctx: int = 0,
                max_tokens: int = 16, entropy_thold: float = 2.8, initial_prompt: bytes = b"",
                n_threads: int | None = None, **overrides) -> WhisperFullParams:
    """The exact parameter block SpeechToText::transcribe builds (/root/reference/src/speech_to_text.cpp:403-413),
    with the project-setting defaults of src/register_types.cpp:64-69."""
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


class Session:
    """One whisper_context of either library (same ABI)."""

    def __init__(self, lib: C.CDLL, model_bytes: bytes, use_gpu: bool = True):
        self.lib = lib
        self._buf = (C.c_char * len(model_bytes)).from_buffer_copy(model_bytes)
        self.ctx = lib.whisper_init_from_buffer_with_params(C.cast(self._buf, C.c_void_p), len(model_bytes),
                                                            WhisperContextParams(use_gpu))
        self._buf = None  # the API contract says the buffer may die right after init
        if not self.ctx:
            raise RuntimeError("whisper_init_from_buffer_with_params returned NULL")

    def close(self):
        if self.ctx:
            self.lib.whisper_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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

    # stage API
    def pcm_to_mel(self, pcm: np.ndarray, n_threads: int = 1) -> int:
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.whisper_pcm_to_mel(self.ctx, pcm.ctypes.data_as(C.POINTER(C.c_float)), int(pcm.size), n_threads)

    def set_mel(self, mel: np.ndarray) -> int:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        n_mel, n_len = mel.shape
        return self.lib.whisper_set_mel(self.ctx, mel.ctypes.data_as(C.POINTER(C.c_float)), n_len, n_mel)

    def encode(self, offset: int = 0, n_threads: int = 1) -> int:
        return self.lib.whisper_encode(self.ctx, offset, n_threads)

    def decode(self, tokens, n_past: int, n_threads: int = 1) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        rc = self.lib.whisper_decode(self.ctx, toks.ctypes.data_as(C.POINTER(C.c_int32)), int(toks.size), n_past,
                                     n_threads)
        if rc != 0:
            raise RuntimeError(f"whisper_decode -> {rc}")
        nv = self.lib.whisper_n_vocab(self.ctx)
        ptr = self.lib.whisper_get_logits(self.ctx)
        all_rows = np.ctypeslib.as_array(ptr, shape=(int(toks.size), nv))
        return np.array(all_rows[-1], dtype=np.float32)  # only the last row is defined (whisper.cpp:457,2566-2572)

    def tokenize(self, text: bytes, cap: int = 1024) -> list:
        buf = (C.c_int32 * cap)()
        n = self.lib.whisper_tokenize(self.ctx, text, buf, cap)
        return list(buf[:max(n, 0)])


class RefSession(Session):
    """Reference context + probe accessors."""

    def mel(self):
        n_org = C.c_int()
        n_len = self.lib.probe_mel(self.ctx, None, 0, C.byref(n_org))
        self.lib.whisper_model_n_mels.argtypes = [C.c_void_p]
        self.lib.whisper_model_n_mels.restype = C.c_int
        out = np.empty((self.lib.whisper_model_n_mels(self.ctx), n_len), dtype=np.float32)
        self.lib.probe_mel(self.ctx, out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(n_org))
        return out, n_org.value

    def _tensor(self, fn):
        ne0, ne1 = C.c_int(), C.c_int()
        n = fn(self.ctx, None, 0, C.byref(ne0), C.byref(ne1))
        out = np.empty((ne1.value, ne0.value), dtype=np.float32)
        fn(self.ctx, out.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(ne0), C.byref(ne1))
        return out

    def embd_conv(self):
        """[n_state][n_ctx] (ggml ne0 = n_ctx fastest)"""
        return self._tensor(self.lib.probe_embd_conv)

    def embd_enc(self):
        """[n_ctx][n_state] (ggml ne0 = n_state fastest)"""
        return self._tensor(self.lib.probe_embd_enc)

    def kv_cross(self):
        n = self.lib.probe_kv_cross(self.ctx, None, None, 0)
        k = np.empty(n, dtype=np.uint16)
        v = np.empty(n, dtype=np.uint16)
        self.lib.probe_kv_cross(self.ctx, k.ctypes.data_as(C.POINTER(C.c_uint16)),
                                v.ctypes.data_as(C.POINTER(C.c_uint16)), n)
        return k.view(np.float16), v.view(np.float16)

    def kv_self(self):
        n = self.lib.probe_kv_self(self.ctx, None, None, 0)
        k = np.empty(n, dtype=np.uint16)
        v = np.empty(n, dtype=np.uint16)
        self.lib.probe_kv_self(self.ctx, k.ctypes.data_as(C.POINTER(C.c_uint16)),
                               v.ctypes.data_as(C.POINTER(C.c_uint16)), n)
        return k.view(np.float16), v.view(np.float16)

    def decode_batch(self, tokens, pos, seq, want, n_threads: int = 1) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        p = np.ascontiguousarray(pos, dtype=np.int32)
        s = np.ascontiguousarray(seq, dtype=np.int32)
        w = np.ascontiguousarray(want, dtype=np.int8)
        ip = C.POINTER(C.c_int)
        rc = self.lib.probe_decode_batch(self.ctx, t.ctypes.data_as(ip), p.ctypes.data_as(ip), s.ctypes.data_as(ip),
                                         w.ctypes.data_as(C.POINTER(C.c_int8)), int(t.size), n_threads)
        if rc != 0:
            raise RuntimeError("probe_decode_batch failed")
        nv = self.lib.whisper_n_vocab(self.ctx)
        rows = np.ctypeslib.as_array(self.lib.whisper_get_logits(self.ctx), shape=(int(t.size), nv))
        out = np.full((int(t.size), nv), np.nan, dtype=np.float32)
        for i in range(int(t.size)):
            if w[i]:
                out[i] = rows[i]
        return out

    def counters(self) -> dict:
        out = (C.c_int * 7)()
        self.lib.probe_counters(self.ctx, out)
        return dict(zip(("n_sample", "n_encode", "n_decode", "n_batchd", "n_prompt", "n_fail_p", "n_fail_h"), out))

    def timings_us(self) -> dict:
        out = (C.c_longlong * 6)()
        self.lib.probe_timings_us(self.ctx, out)
        return dict(zip(("mel", "sample", "encode", "decode", "batchd", "prompt"), out))


def tiny_en_model_path() -> str | None:
    """Real tiny.en weights: staged copy first (travels to the GPU box), the reference mount second."""
    for p in (os.path.join(REF_DIR, "ggml-tiny.en.bin"),
              "/root/reference/bin/addons/godot_whisper/models/gglm-tiny.en.bin"):
        if os.path.exists(p):
            return p
    return None


def jfk_wav_path() -> str | None:
    for p in (os.path.join(HERE, "..", "tests", "golden", "jfk.wav"), os.path.join(REF_DIR, "jfk.wav"),
              "/root/reference/thirdparty/whisper.cpp/samples/jfk.wav"):
        if os.path.exists(p):
            return os.path.abspath(p)
    return None


def jfk30(pcm: np.ndarray) -> np.ndarray:
    """jfk.wav tiled x3 and truncated to 480 000 samples (30 s)."""
    return np.tile(pcm, 3)[:480000].copy()
