#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for dbg in 0 1 2 3; do echo "== WHISPER_B200_GEMM_DBG=$dbg"; WHISPER_B200_GEMM_DBG=$dbg timeout 300 python tools/gemm_enc_bench.py 384 16; done > $O/gemm_enc_dbg.md 2>&1; cat $O/gemm_enc_dbg.md
