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
