"""C-ABI checks that need no GPU: the library loads, exports every symbol include/whisper_b200.h declares, and its by-value
structs / default parameter block are byte-identical to the compiled reference's (whisper.h:87-106, 433-526; whisper.cpp:4311-4410)."""
import ctypes as C
import os
import re

import numpy as np

from conftest import ROOT
import whisper_b200 as wb


def declared_in_header():
    txt = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"WHISPER_B200_API[^;(]*?\b(whisper_\w+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_in_header() == sorted(wb.DECLARED_SYMBOLS)


def test_library_exports_every_declared_symbol(product):
    missing = [s for s in declared_in_header() if not hasattr(product, s)]
    assert missing == []


def test_struct_sizes_match_reference(ref):
    assert ref.probe_sizeof_full_params() == C.sizeof(wb.WhisperFullParams) == 256
    assert ref.probe_sizeof_token_data() == C.sizeof(wb.WhisperTokenData) == 48


def _param_bytes(p):
    raw = bytearray(C.string_at(C.addressof(p), C.sizeof(p)))
    for name in ("initial_prompt", "language"):           # pointers into each library's own rodata
        off = getattr(wb.WhisperFullParams, name).offset
        raw[off:off + 8] = b"\0" * 8
    return bytes(raw)


def test_default_params_match_reference(ref, product):
    for strategy in (0, 1):
        a, b = ref.whisper_full_default_params(strategy), product.whisper_full_default_params(strategy)
        assert a.language == b.language == b"en"
        assert _param_bytes(a) == _param_bytes(b)


def test_init_fails_loudly_without_usable_input(product):
    """NULL on garbage (src/speech_to_text.cpp:346-349 expects that), never a silent fallback."""
    log = []
    wb.set_log_sink(product, log)
    junk = (C.c_char * 64)(*b"not a ggml file" + b"\0" * 49)
    ctx = product.whisper_init_from_buffer_with_params(C.cast(junk, C.c_void_p), 64, wb.WhisperContextParams(True))
    assert not ctx
    assert any(level == 2 for level, _ in log)
    product.whisper_free(None)                            # NULL-safe (src/speech_to_text.cpp:332)
    assert b"CPU_FALLBACK = 0" in product.whisper_print_system_info()


def test_activation_tables_equal_reference(ref, ref_session, product):
    """GELU / exp f16 tables (ggml.c:2218-2236): the kernels' lookups are bit-identical to the reference's."""
    u16p = C.POINTER(C.c_uint16)
    g, e = np.zeros(65536, np.uint16), np.zeros(65536, np.uint16)
    rg, re_ = np.zeros(65536, np.uint16), np.zeros(65536, np.uint16)
    product.whisper_b200_f16_tables(g.ctypes.data_as(u16p), e.ctypes.data_as(u16p))
    ref.probe_f16_tables(rg.ctypes.data_as(u16p), re_.ctypes.data_as(u16p))
    assert np.array_equal(g, rg)
    assert np.array_equal(e, re_)


def _init(product, blob):
    buf = (C.c_char * len(blob)).from_buffer_copy(blob)
    return product.whisper_init_from_buffer_with_params(C.cast(buf, C.c_void_p), len(blob), wb.WhisperContextParams(True))


def test_crafted_headers_return_null(product, model_bytes):
    """A hostile ggml header must end in NULL (src/speech_to_text.cpp:346-349 checks for it), never in SIGFPE / a 2^31-iteration
    loop / bad_alloc across the C ABI.  hparams are 11 int32 behind the magic (whisper.cpp:1127-1138): n_vocab, n_audio_ctx,
    n_audio_state, n_audio_head, n_audio_layer, n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype."""
    import struct
    log = []
    wb.set_log_sink(product, log)
    head = bytearray(model_bytes[:1 << 16])               # header + filters + part of the vocabulary is all the parser may touch

    def patched(index, value):
        b = bytearray(head)
        struct.pack_into("<i", b, 4 + 4 * index, value)
        return bytes(b)

    cases = [(7, 0), (3, 0), (7, -6), (4, 1 << 30), (8, 1 << 30), (8, -1), (0, 1 << 30), (0, -5), (1, 1 << 30), (5, 1 << 30),
             (9, 1 << 30), (2, 0), (2, 1 << 30), (10, 0), (10, 5), (10, 12), (10, 7)]
    for index, value in cases:
        del log[:]
        assert not _init(product, patched(index, value)), (index, value)
        # rejected by the header check itself, not further down the road (vocabulary, tensors, device)
        assert any("invalid model hyper-parameters" in m or "unsupported model ftype" in m or "unsupported quantisation format" in m
                   for _, m in log), (index, value, log[-3:])
    # a tensor record whose name / dims run past the end of the buffer
    n_mel, n_fft = struct.unpack_from("<ii", model_bytes, 48)
    off = 56 + 4 * n_mel * n_fft
    n_vocab_file, = struct.unpack_from("<i", model_bytes, off)
    off += 4
    for _ in range(n_vocab_file):
        ln, = struct.unpack_from("<I", model_bytes, off)
        off += 4 + ln
    for rec in (struct.pack("<iii", 4, 64, 1) + b"\0" * 10, struct.pack("<iii", 2, 1 << 20, 1) + b"\0" * 40,
                struct.pack("<iiiii", 2, 20, 1, -3, 7) + b"encoder.conv1.weight", struct.pack("<iii", 3, 4, 1) + b"\0\0"):
        del log[:]
        assert not _init(product, model_bytes[:off] + rec)
        assert any("corrupt tensor record" in m for _, m in log), log[-3:]


def test_dequantizer_equals_ggml(ref, product):
    """The loader expands block-quantised matrices (Q4_0 / Q4_1 / Q5_0 / Q5_1 / Q8_0 files of whisper.cpp's quantize tool) with
    csrc/model.cpp::dequantize_blocks: bit for bit what ggml-quants.c dequantize_row_* yields on blocks the reference itself produced."""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(32 * 4096) * 0.05).astype(np.float32)
    x[:64] = 0.0
    hist = (C.c_int64 * 16)()
    ref.ggml_quantize_chunk.restype = C.c_size_t
    ref.ggml_quantize_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for name, gtype, bsz in (("q4_0", 2, 18), ("q4_1", 3, 20), ("q5_0", 6, 22), ("q5_1", 7, 24), ("q8_0", 8, 34)):
        blocks = np.empty(x.size // 32 * bsz, np.uint8)
        assert ref.ggml_quantize_chunk(gtype, x.ctypes.data, blocks.ctypes.data, 0, x.size, C.addressof(hist)) == blocks.size
        want, mine = np.empty_like(x), np.empty_like(x)
        fn = getattr(ref, "dequantize_row_" + name)
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        fn.restype = None
        fn(blocks.ctypes.data, want.ctypes.data, x.size)
        assert product.whisper_b200_dequantize(gtype, blocks.ctypes.data, x.size, mine.ctypes.data_as(C.POINTER(C.c_float))) == 0
        # (the reference compiles x * d + m with its own contraction choices: compare after the f16 rounding the loader applies)
        assert np.array_equal(mine.astype(np.float16), want.astype(np.float16)), name
        assert np.abs(mine - want).max() <= 1e-6
    assert product.whisper_b200_dequantize(12, blocks.ctypes.data, 64, mine.ctypes.data_as(C.POINTER(C.c_float))) == -1      # k-quants: refused
