#!/usr/bin/env python
"""End-to-end audio-s/s of ONE process driving N GPUs through the C ABI (whisper_b200_init_multi + whisper_b200_full_batch) — what a
single-process host such as the GDExtension gets without torchrun.  python tools/multi_device_bench.py [N] [chunks_per_gpu] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
import bench  # noqa: E402  (inputs only)
import whisper_b200 as wb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    blob, _ = bench.model_bytes_for("tiny.en")
    lib = wb.load_library()
    ctx = wb.Context(blob, devices=list(range(n)))
    assert lib.whisper_b200_n_devices(ctx.ctx) == n
    chunks = bench.load_inputs(per * n)
    p = wb.host_params(lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
    gold = bench.load_golden()
    import hashlib
    for _ in range(2):
        assert ctx.full_batch(p, chunks) == 0
    t0 = time.perf_counter()
    for _ in range(steps):
        assert ctx.full_batch(p, chunks) == 0
        texts = [ctx.chunk_text(i) for i in range(len(chunks))]
    dt = time.perf_counter() - t0
    same = sum(hashlib.sha1(t).digest() == gold[i % len(gold)] for i, t in enumerate(texts)) if gold else None
    print({"n_gpus": n, "chunks_per_step": len(chunks), "e2e_audio_s_per_s": 30.0 * len(chunks) * steps / dt, "ms_per_step": dt * 1e3 / steps,
           "transcripts_identical_to_oracle": same})
    ctx.close()


if __name__ == "__main__":
    main()
