#!/bin/bash
# attention variants (bit identity + timing) and the encoder GEMM shapes outside ncu
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused_encoder_attention_variants or encoder_gemm_tma_store" > $O/pytest_probe_a.log 2>&1; tail -5 $O/pytest_probe_a.log
for args in "6 16 1500" "6 32 1500" "8 8 1500" "6 16 628"; do timeout 300 python tools/attn_enc_bench.py $args >> $O/attn_enc_bench.md 2>&1; done; cat $O/attn_enc_bench.md
timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_384.md 2>&1; cat $O/gemm_enc_bench_384.md
timeout 300 python tools/gemm_enc_bench.py 384 32 > $O/gemm_enc_bench_384_32.md 2>&1; cat $O/gemm_enc_bench_384_32.md
timeout 300 python tools/gemm_enc_bench.py 512 8 > $O/gemm_enc_bench_512.md 2>&1; cat $O/gemm_enc_bench_512.md
