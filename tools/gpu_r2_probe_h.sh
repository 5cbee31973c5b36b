#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "encoder_gemm_tma_store or conv_stem or encoder_output or cross_kv or greedy_transcript or sixteen_chunk" > $O/pytest_probe_h.log 2>&1; tail -4 $O/pytest_probe_h.log

timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_h.md 2>&1; cat $O/gemm_enc_bench_h.md
