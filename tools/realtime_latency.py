#!/usr/bin/env python
"""BASELINE.json configs[4]: the CaptureStreamToText loop (bin/addons/godot_whisper/capture_stream_to_text.gd:63-120 of the
reference) against this backend — a growing 16 kHz buffer, one whisper_full() per transcribe_interval (1 s) with the host's
parameter block and `audio_ctx = seconds * 50 + 128` — and the per-call wall time (p50 / max).  With --reference the same calls
also go to the compiled reference (oracle/_ref, 4 threads): its latency, and whether the token ids agree call by call.
  python tools/realtime_latency.py [--reference] [--repeat 3]"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))

import whisper_b200 as wb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    model = open(os.path.join(ROOT, "oracle", "_ref", "ggml-tiny.en.bin"), "rb").read()
    pcm = wb.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
    lib = wb.load_library()
    ctx = wb.Context(model, device=0)
    n_ticks = len(pcm) // 16000

    def ids(res):
        return [t["id"] for s in res["segments"] for t in s["tokens"]]

    def run(full, params_of, result):
        lat, out = [], []
        for t in range(1, n_ticks + 1):
            buf = pcm[: t * 16000]
            p = params_of(audio_ctx=min(1500, t * 50 + 128))
            t0 = time.perf_counter()
            assert full(p, buf) == 0
            lat.append((time.perf_counter() - t0) * 1e3)
            out.append(ids(result()))
        return lat, out

    ours = None
    for _ in range(a.repeat):        # the first sweep also warms the CUDA graphs / tensor maps of every audio_ctx
        ours = run(ctx.full, lambda **kw: wb.host_params(lib, n_threads=4, **kw), ctx.result)
    res = {"config": "CaptureStreamToText realtime: jfk.wav streamed in 1 s increments, tiny.en, host parameter block (max_tokens 16, entropy_thold 2.8), "
                     "audio_ctx = seconds*50+128", "calls": n_ticks, "ours_ms_p50": statistics.median(ours[0]), "ours_ms_max": max(ours[0]),
           "ours_ms": [round(x, 2) for x in ours[0]]}
    if a.reference:
        from oracle import ref_lib       # checker / CPU baseline only
        rlib = ref_lib.load()
        rs = ref_lib.RefSession(rlib, model, use_gpu=False)
        ref = run(rs.full, lambda **kw: ref_lib.host_params(rlib, n_threads=4, **kw), rs.result)
        res.update({"reference_ms_p50": statistics.median(ref[0]), "reference_ms_max": max(ref[0]),
                    "token_ids_equal_per_call": [x == y for x, y in zip(ours[1], ref[1])]})
        rs.close()
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
