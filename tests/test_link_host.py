"""Link test of the UNTOUCHED GDExtension host against the drop-in: /root/reference/src/*.cpp are compiled as they
lie — against godot-cpp bindings generated on the spot, the reference's libsamplerate header, and include/whisper_b200.h standing in
for <whisper.cpp/whisper.h> — and every whisper_* / ggml_* symbol the resulting objects import must be exported by
libwhisper_b200.so with default visibility.  Running inside Godot is not possible here (no godot binary); this is the proof that the
boundary is complete: same names, same by-value struct layouts (the host's by-value calls compile against our declarations), nothing
missing at link time.  Needs /root/reference (skipped on the GPU box)."""
import glob
import os
import subprocess
import sys

import pytest

from conftest import ROOT
import whisper_b200 as wb

REF = "/root/reference"


@pytest.fixture(scope="module")
def host_objects(tmp_path_factory):
    if not os.path.isdir(os.path.join(REF, "src")) or not os.path.isdir(os.path.join(REF, "thirdparty", "godot-cpp")):
        pytest.skip("/root/reference not mounted")
    work = tmp_path_factory.mktemp("linkhost")
    gcpp = os.path.join(REF, "thirdparty", "godot-cpp")
    gen = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); import binding_generator as b; "
                          "b.generate_bindings(%r, True, '64', 'single', %r)" % (gcpp, os.path.join(gcpp, "gdextension", "extension_api.json"), str(work))],
                         capture_output=True, text=True)
    assert gen.returncode == 0, gen.stderr[-2000:]
    # <whisper.cpp/whisper.h> resolves to OUR header (the reference's header is not on the include path at all)
    shim = work / "shim" / "whisper.cpp"
    shim.mkdir(parents=True)
    (shim / "whisper.h").write_text('#include "%s"\n' % os.path.join(ROOT, "include", "whisper_b200.h"))
    inc = ["-I" + str(work / "shim"), "-I" + os.path.join(gcpp, "include"), "-I" + os.path.join(gcpp, "gdextension"), "-I" + str(work / "gen" / "include"),
           "-I" + os.path.join(REF, "src"),
           "-I" + os.path.join(REF, "thirdparty")]        # <libsamplerate/src/samplerate.h>; the shim directory comes first, so whisper.h is ours
    objs = []
    for src in sorted(glob.glob(os.path.join(REF, "src", "*.cpp"))):
        obj = str(work / (os.path.basename(src) + ".o"))
        res = subprocess.run(["g++", "-std=c++17", "-fPIC", "-O1", "-DWHISPER_SHARED", "-DGGML_SHARED", "-c", src, "-o", obj] + inc, capture_output=True, text=True)
        assert res.returncode == 0, (src, res.stderr[-3000:])
        objs.append(obj)
    return work, objs


def undefined(objs):
    out = subprocess.run(["nm", "-u"] + objs, capture_output=True, text=True, check=True).stdout
    return sorted({l.split()[-1] for l in out.splitlines() if l.strip().startswith("U ")})


def test_untouched_host_compiles_against_our_header_and_links(host_objects, product):
    work, objs = host_objects
    assert len(objs) == 4
    need = [s for s in undefined(objs) if s.startswith(("whisper_", "ggml_"))]
    # exactly the boundary SURVEY.md §8b lists: eleven functions
    assert need == sorted(["whisper_init_from_buffer_with_params", "whisper_free", "whisper_print_system_info", "whisper_full_default_params",
                           "whisper_full", "whisper_full_n_segments", "whisper_full_n_tokens", "whisper_full_get_segment_text",
                           "whisper_full_get_token_text", "whisper_full_get_token_data", "whisper_log_set"])
    exported = subprocess.run(["nm", "-D", "--defined-only", wb.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in exported.splitlines() if l.strip()}
    assert [s for s in need if s not in exported] == []
    # and an actual link: the host objects + our library, everything that is not ours (godot-cpp runtime, libsamplerate) left open
    so = str(work / "libgodot_whisper_host.so")
    res = subprocess.run(["g++", "-shared", "-o", so] + objs + ["-L" + os.path.dirname(wb.LIB_PATH), "-lwhisper_b200", "-Wl,--unresolved-symbols=ignore-all",
                          "-Wl,-rpath," + os.path.dirname(wb.LIB_PATH)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True, check=True).stdout
    assert "libwhisper_b200.so" in dyn
    # none of the whisper symbols is left dangling in the linked host: the dynamic linker finds each in libwhisper_b200.so
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "libwhisper_b200.so" in ldd and "not found" not in [l for l in ldd.splitlines() if "whisper_b200" in l][0]
