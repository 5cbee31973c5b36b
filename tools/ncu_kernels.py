#!/usr/bin/env python
"""One launch of each encoder tensor-core kernel at the shapes of a 16-chunk tiny.en encoder pass, for `ncu --set full`:
k_gemm_enc in its four modes (QKV three segments, out-proj + residual by TMA, FC1 + GELU, FC2 + residual read by the epilogue,
cross K / V^T) and the fused attention kernel in its default configuration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
import whisper_b200 as wb  # noqa: E402


def main():
    d, chunks, T, n_head = 384, 16, 1500, 6
    N = T * chunks
    rng = np.random.default_rng(0)
    for name, M, K, mode in (("qkv (3 segments)", 3 * d, d, 4), ("out-proj + residual", d, d, 3), ("fc1 + gelu", 4 * d, d, 1),
                             ("fc2 + residual", d, 4 * d, 3), ("cross kv transposed", 2 * d, d, 2)):
        act = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        wgt = (rng.standard_normal((M, K)) * 0.05).astype(np.float16)
        bias = rng.standard_normal(M).astype(np.float32)
        res = rng.standard_normal((N, M)).astype(np.float32) if mode == 3 else None
        wb.gemm_enc_probe(act, wgt, mode, bias=bias, res=res, iters=0)
        print("launched", name, flush=True)
    Tp = (T + 7) & ~7
    q = (rng.standard_normal((chunks, T, d)) * 1.5).astype(np.float16)
    k = (rng.standard_normal((chunks, T, d)) * 1.5).astype(np.float16)
    vt = np.zeros((chunks, d, Tp), np.float16)
    vt[:, :, :T] = rng.standard_normal((chunks, d, T)).astype(np.float16)
    wb.attn_enc_probe(q, k, vt, n_head, variant=-1, iters=0)
    print("launched attention", flush=True)


if __name__ == "__main__":
    main()
