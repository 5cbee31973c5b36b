#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_d.log 2>&1; tail -4 $O/pytest_probe_d.log
timeout 900 python bench.py > $O/bench_d.json 2> $O/bench_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','kernel_classes','encoder_gemm_roofline','base_en_b8_beam5','gpu_launches'):
    print(k, json.dumps(d.get(k))[:900])
PY
mkdir -p gpurun_out; O=gpurun_out
for dbg in 0 1 2 3; do echo "== WHISPER_B200_GEMM_DBG=$dbg"; WHISPER_B200_GEMM_DBG=$dbg timeout 300 python tools/gemm_enc_bench.py 384 16; done > $O/gemm_enc_dbg.md 2>&1; cat $O/gemm_enc_dbg.md
