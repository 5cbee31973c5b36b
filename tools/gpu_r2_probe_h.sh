#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_h.log 2>&1; tail -4 $O/pytest_probe_h.log
ATTN_VARIANTS=0,1,4 timeout 300 python tools/attn_enc_bench.py 6 16 1500 > $O/attn_enc_bench_h.md 2>&1; cat $O/attn_enc_bench_h.md
timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_h.md 2>&1; cat $O/gemm_enc_bench_h.md
