#!/usr/bin/env python
"""A short, fixed workload for `ncu`: one warm step + one step of whisper_b200_full_batch (B chunks of 30 s, greedy) with the
same parameter block bench.py uses.  Prints the kernel launch count of the second step so `-s/-c` can be chosen.
  python tools/ncu_workload.py [--model tiny.en|base.en] [--batch B] [--steps K]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import bench  # noqa: E402  (input / model helpers only)
import whisper_b200 as wb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny.en")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--encode-only", action="store_true", help="only the encoder pass of one batch (whisper_encode semantics)")
    a = ap.parse_args()
    blob, note = bench.model_bytes_for(a.model)
    lib = wb.load_library()
    ctx = wb.Context(blob, device=0)
    chunks = bench.load_inputs(a.batch)
    p = wb.host_params(lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
    for s in range(a.steps):
        c0 = ctx.counters()["launches"]
        assert ctx.full_batch(p, chunks) == 0
        print(f"step {s}: launches {ctx.counters()['launches'] - c0}", flush=True)
    print("gpu ms:", ctx.gpu_times(), note)
    ctx.close()


if __name__ == "__main__":
    main()
