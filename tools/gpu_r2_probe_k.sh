#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "encoder_gemm_tma_store or conv_stem or encoder_output or greedy_transcript" > $O/pytest_probe_k.log 2>&1; tail -3 $O/pytest_probe_k.log
timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_k.md 2>&1; cat $O/gemm_enc_bench_k.md
WHISPER_B200_GEMM_DBG=11 timeout 300 python tools/gemm_enc_bench.py 384 16 2>&1 | grep -E "fc1|shape"
