#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for g in 2 1 2 1; do echo "== WHISPER_B200_GEMM_GROUPS=$g (32 chunks)"; WHISPER_B200_GEMM_GROUPS=$g timeout 300 python tools/gemm_enc_bench.py 384 32 2>&1 | cut -d'|' -f2,6,7; done > $O/gemm_enc_groups.md 2>&1; cat $O/gemm_enc_groups.md
