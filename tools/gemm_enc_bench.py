#!/usr/bin/env python
"""Times the encoder GEMM shapes of one 16-chunk pass in isolation (CUDA events around `iters` back-to-back launches):
gemm_enc.cu in each epilogue mode next to the first-generation kernel (plain f32 output) — TFLOP/s and fraction of the measured peak."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
import whisper_b200 as wb  # noqa: E402


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    N = 1500 * chunks
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    rng = np.random.default_rng(0)
    rows = []
    for name, M, K, mode in (("qkv (3 segments)", 3 * d, d, 4), ("qkv as plain f16", 3 * d, d, 0), ("out-proj + residual", d, d, 3), ("fc1 + gelu", 4 * d, d, 1),
                             ("fc1 plain f16", 4 * d, d, 0), ("fc2 + residual", d, 4 * d, 3), ("cross kv transposed", 2 * d, d, 2)):
        act = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        wgt = (rng.standard_normal((M, K)) * 0.05).astype(np.float16)
        bias = rng.standard_normal(M).astype(np.float32)
        res = rng.standard_normal((N, M)).astype(np.float32) if mode == 3 else None
        _, ms = wb.gemm_enc_probe(act, wgt, mode, bias=bias, res=res, iters=20)
        _, ms_old = wb.gemm_f16(wgt, act, engine=0, iters=20)
        fl = 2.0 * N * M * K
        rows.append((name, N, M, K, ms * 1e3, fl / (ms * 1e-3) / 1e12, fl / (ms * 1e-3) / 1e12 / peak, ms_old * 1e3))
    print("| shape | N | M | K | v2 us | v2 TFLOP/s | frac of %.0f | v1 (f32 out) us |" % peak)
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %s | %d | %d | %d | %.1f | %.0f | %.3f | %.1f |" % r)


if __name__ == "__main__":
    main()
