#!/bin/bash
# attention pipeline restructure + branchless GELU epilogue: parity suite, then the two micro-benchmarks
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_b.log 2>&1; tail -5 $O/pytest_probe_b.log
rm -f $O/attn_enc_bench_b.md
for args in "6 16 1500" "6 32 1500" "8 8 1500"; do timeout 300 python tools/attn_enc_bench.py $args >> $O/attn_enc_bench_b.md 2>&1; done; cat $O/attn_enc_bench_b.md
timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_384_b.md 2>&1; cat $O/gemm_enc_bench_384_b.md
