#!/bin/bash
# weight-stationary encoder GEMM: epilogue tests, then timings with and without it
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "encoder_gemm_tma_store or conv_stem or encoder_output or cross_kv or greedy_transcript" > $O/pytest_probe_f.log 2>&1; tail -4 $O/pytest_probe_f.log
for ws in 1 0; do echo "== WHISPER_B200_GEMM_WS=$ws"; WHISPER_B200_GEMM_WS=$ws timeout 300 python tools/gemm_enc_bench.py 384 16; WHISPER_B200_GEMM_WS=$ws timeout 300 python tools/gemm_enc_bench.py 512 8; done > $O/gemm_enc_ws.md 2>&1; cat $O/gemm_enc_ws.md
for dbg in 4 7 8 11 15; do echo "== WHISPER_B200_GEMM_WS=0 WHISPER_B200_GEMM_DBG=$dbg"; WHISPER_B200_GEMM_WS=0 WHISPER_B200_GEMM_DBG=$dbg timeout 300 python tools/gemm_enc_bench.py 384 16; done > $O/gemm_enc_dbg2.md 2>&1; cat $O/gemm_enc_dbg2.md
