#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_i.log 2>&1; tail -4 $O/pytest_probe_i.log
timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench_i.md 2>&1; cat $O/gemm_enc_bench_i.md
timeout 300 python tools/gemm_enc_bench.py 512 8 > $O/gemm_enc_bench_i512.md 2>&1; cat $O/gemm_enc_bench_i512.md
timeout 900 python bench.py --no-host-block > $O/bench_i.json 2> $O/bench_i.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','kernel_classes','encoder_gemm_roofline','base_en_b8_beam5','gpu_launches'):
    print(k, json.dumps(d.get(k))[:1200])
PY
