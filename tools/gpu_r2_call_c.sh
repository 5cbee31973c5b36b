#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tiny.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','e2e_host_block','base_en_b8_beam5','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:700])
PY
tail -3 $O/bench_tiny.err
