"""Host front of the realtime path: whisper_b200_vad_simple / whisper_b200_high_pass_filter against the reference's
own _vad_simple / _high_pass_filter (src/speech_to_text.cpp:53-104), whose function text oracle/Makefile cuts out of the reference file
and compiles against stand-ins for the godot-cpp names it uses (oracle/ref_vad.cpp).  Bit-exact: the filtered window and the decision."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import ref_lib
import whisper_b200 as wb

VAD_SO = os.path.join(ROOT, "oracle", "_ref", "libvad_ref.so")


@pytest.fixture(scope="module")
def ref_vad():
    if not os.path.exists(VAD_SO):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/libvad_ref.so not built and /root/reference not mounted")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libvad_ref.so"], check=True, capture_output=True)
    lib = C.CDLL(VAD_SO)
    fp = C.POINTER(C.c_float)
    lib.ref_high_pass_filter.argtypes = [fp, C.c_int, C.c_float, C.c_float]
    lib.ref_high_pass_filter.restype = None
    lib.ref_vad_simple.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    lib.ref_vad_simple.restype = C.c_int
    return lib


def windows():
    """Three-second windows as SpeechToText::voice_activity_detection cuts them (src/speech_to_text.cpp:378-399): speech that ends, speech
    that goes on, noise, near-silence, silence, a window shorter than the 500 ms tail, an empty tail."""
    jfk = ref_lib.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
    rng = np.random.default_rng(5)
    n = 48000
    out = [("jfk %d" % s, jfk[s:s + n]) for s in range(0, len(jfk) - n, 16000)]
    out.append(("jfk tail then silence", np.concatenate([jfk[-32000:], np.zeros(16000, np.float32)])))
    out.append(("noise", (rng.standard_normal(n) * 0.05).astype(np.float32)))
    out.append(("faint noise", (rng.standard_normal(n) * 2e-5).astype(np.float32)))
    out.append(("faint noise, louder tail", np.concatenate([(rng.standard_normal(n - 8000) * 2e-5), rng.standard_normal(8000) * 8e-5]).astype(np.float32)))
    out.append(("silence", np.zeros(n, np.float32)))
    out.append(("short", jfk[:4000]))
    return out


@pytest.mark.parametrize("vad_thold,freq_thold", [(0.3, 200.0), (0.6, 100.0), (0.3, 0.0), (2.0, 200.0)])
def test_vad_simple_equals_reference(ref_vad, vad_thold, freq_thold):
    fp = C.POINTER(C.c_float)
    decisions = []
    for name, w in windows():
        for last_ms in (500, 0, 1000):
            r = np.array(w, dtype=np.float32, copy=True)
            want = ref_vad.ref_vad_simple(r.ctypes.data_as(fp), r.size, 16000, last_ms, vad_thold, freq_thold)
            got, mine = wb.vad_simple(w, 16000, last_ms, vad_thold, freq_thold)
            assert got == want, (name, last_ms)
            assert np.array_equal(mine.view(np.uint32), r.view(np.uint32)), (name, last_ms)     # the in-place high-pass, bit for bit
            decisions.append(got)
    assert 0 in decisions
    if vad_thold < 1.0:
        assert 1 in decisions           # (faint windows: the only case in which the host's variant of the test answers true)


def test_high_pass_filter_equals_reference(ref_vad):
    fp = C.POINTER(C.c_float)
    lib = wb.load_library()
    rng = np.random.default_rng(6)
    for n in (1, 2, 1000, 48000):
        for cutoff in (50.0, 200.0, 3000.0):
            x = (rng.standard_normal(n) * 0.3).astype(np.float32)
            a, b = x.copy(), x.copy()
            ref_vad.ref_high_pass_filter(a.ctypes.data_as(fp), n, cutoff, 16000.0)
            lib.whisper_b200_high_pass_filter(b.ctypes.data_as(fp), n, cutoff, 16000.0)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
