#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
ATTN_VARIANTS=10,18,19,20,21 timeout 300 python tools/attn_enc_bench.py 6 16 1500 > $O/attn_enc_bench_c.md 2>&1; cat $O/attn_enc_bench_c.md
