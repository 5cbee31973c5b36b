#!/bin/bash
# tests + bench + ncu launch list (128 chunks)
mkdir -p gpurun_out; O=gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py --no-base-en > $O/bench_tiny.json 2> $O/bench_tiny.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tiny.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','kernel_classes','transcripts_vs_oracle'):
    print(k, json.dumps(d.get(k))[:900])
PY
grep "full_batch: 512\|run steps\|^device" $O/bench_tiny.err | tail -6
bash tools/gpu_r2_ncu_list.sh 2>&1 | head -24
