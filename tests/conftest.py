"""Shared fixtures.  Two kinds of tests live here:

  -m "not gpu"  run on a CPU-only box: the oracle (compiled reference, oracle/_ref) against the reference's known
                answers, the host logic of the product (its C++ driver linked against a TEST-ONLY checker forward that
                delegates tensor math to the compiled reference), ABI/exports of libwhisper_b200.so, gloo sharding.
  -m gpu        the parity tests proper: libwhisper_b200.so (CUDA, sm_100a) through its C ABI against the oracle.

/root/reference is only needed to BUILD oracle/_ref (done by __graft_entry__.build() / `make -C oracle`); nothing here reads
it at run time.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "godot-whisper_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)

from oracle import ref_lib  # noqa: E402  (test infrastructure)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _run(cmd, **kw):
    res = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if res.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{res.stdout[-3000:]}\n{res.stderr[-3000:]}")


def build_hostlogic() -> str:
    """Product host sources + tests/hostlogic/forward_checker.cpp -> tests/_build/libwhisper_hostlogic.so (no CUDA)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libwhisper_hostlogic.so")
    srcs = [os.path.join(PKG, "csrc", f) for f in ("api.cpp", "model.cpp", "mel.cpp", "decode_host.cpp", "full.cpp", "tables.cpp",
                                                   "batcher.cpp")]
    srcs = [s for s in srcs if os.path.exists(s)]
    srcs.append(os.path.join(ROOT, "tests", "hostlogic", "forward_checker.cpp"))
    deps = srcs + [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith(".h")]
    deps.append(os.path.join(ROOT, "include", "whisper_b200.h"))
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    _run(["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-o", out] + srcs + ["-ldl"])
    return out


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref) with probe accessors."""
    if not ref_lib.available():
        if os.path.isdir("/root/reference"):
            _run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "all", "stage"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference not mounted")
    return ref_lib.load()


@pytest.fixture(scope="session")
def model_bytes():
    p = ref_lib.tiny_en_model_path()
    if p is None:
        pytest.skip("tiny.en weights not staged (make -C oracle stage)")
    with open(p, "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def jfk():
    return ref_lib.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))


@pytest.fixture(scope="session")
def hostlogic(ref):
    """Product C++ driver linked to the checker forward (tensor math = compiled reference)."""
    import whisper_b200 as wb
    os.environ["WHISPER_HOSTLOGIC_REF_LIB"] = ref_lib.ref_lib_path()
    lib = wb.load_library(build_hostlogic())
    wb.set_log_sink(lib, None)
    return lib


@pytest.fixture(scope="session")
def product():
    """libwhisper_b200.so (built in-tree by __graft_entry__.build())."""
    import whisper_b200 as wb
    if not os.path.exists(wb.LIB_PATH):
        wb.build()
    lib = wb.load_library()
    return lib


def have_gpu() -> bool:
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx(product, model_bytes):
    """One product context with real tiny.en weights on cuda:0.  Fails (not skips) when the CUDA path cannot start."""
    import whisper_b200 as wb
    log = []
    wb.set_log_sink(product, log)
    try:
        ctx = wb.Context(model_bytes, lib=product)
    except Exception as e:  # pragma: no cover
        pytest.fail(f"CUDA path failed to start: {e}; log tail: {log[-5:]}")
    ctx.log = log
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def ref_session(ref, model_bytes):
    s = ref_lib.RefSession(ref, model_bytes, use_gpu=False)
    yield s
    s.close()


def ids_of(result):
    return [t["id"] for s in result["segments"] for t in s["tokens"]]
