#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused_encoder_attention_variants" > $O/pytest_probe_j.log 2>&1; tail -4 $O/pytest_probe_j.log
rm -f $O/attn_enc_bench_j.md
for args in "6 16 1500" "6 32 1500" "8 8 1500"; do ATTN_VARIANTS=0,5,7,8 timeout 300 python tools/attn_enc_bench.py $args >> $O/attn_enc_bench_j.md 2>&1; done; cat $O/attn_enc_bench_j.md
