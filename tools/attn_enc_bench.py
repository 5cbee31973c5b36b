#!/usr/bin/env python
"""Times the fused encoder attention (csrc/cuda/attn_enc.cu) in isolation for every kernel configuration: CUDA events around `iters`
back-to-back launches of one layer of a `chunks`-chunk encoder pass.  flop = 4 T^2 64 per head (the reference's count)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
import whisper_b200 as wb  # noqa: E402


def main():
    n_head = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
    d, Tp = 64 * n_head, (T + 7) & ~7
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    rng = np.random.default_rng(0)
    q = (rng.standard_normal((chunks, T, d)) * 1.5).astype(np.float16)
    k = (rng.standard_normal((chunks, T, d)) * 1.5).astype(np.float16)
    vt = np.zeros((chunks, d, Tp), np.float16)
    vt[:, :, :T] = rng.standard_normal((chunks, d, T)).astype(np.float16)
    fl = 4.0 * chunks * n_head * T * T * 64.0
    print("| variant | us per launch (%d chunks, %d heads, T = %d) | TFLOP/s (reference's flop count) | frac of %.0f | same bits as variant 0 |" % (chunks, n_head, T, peak))
    print("|---|---|---|---|---|")
    base = None
    for v in (int(x) for x in os.environ.get('ATTN_VARIANTS', '0,1,2,3,4').split(',')):
        out, ms = wb.attn_enc_probe(q, k, vt, n_head, variant=v, iters=20)
        if base is None:
            base = out
        print("| %d | %.1f | %.0f | %.3f | %s |" % (v, ms * 1e3, fl / (ms * 1e-3) / 1e12, fl / (ms * 1e-3) / 1e12 / peak, np.array_equal(base.view(np.uint16), out.view(np.uint16))))


if __name__ == "__main__":
    main()
