#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "encoder_gemm_tma_store or conv_stem or encoder_output or cross_kv or greedy_transcript" > $O/pytest_probe_g.log 2>&1; tail -4 $O/pytest_probe_g.log
for ew in 16 8; do echo "== WHISPER_B200_GEMM_EPI_WARPS=$ew"; WHISPER_B200_GEMM_EPI_WARPS=$ew timeout 300 python tools/gemm_enc_bench.py 384 16; done > $O/gemm_enc_epi_warps.md 2>&1; cat $O/gemm_enc_epi_warps.md
echo "== epilogue only (DBG 11), 16 warps"; WHISPER_B200_GEMM_DBG=11 timeout 300 python tools/gemm_enc_bench.py 384 16
