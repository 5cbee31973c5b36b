#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_enc -s 2 -c 12 -f -o $O/prof_gemm_enc python tools/gemm_enc_bench.py 384 16 > $O/ncu_gemm_enc.log 2>&1
tail -3 $O/ncu_gemm_enc.log; ls -la $O/prof_gemm_enc.ncu-rep
