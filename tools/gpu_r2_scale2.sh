#!/bin/bash
# two ranks on one box with the final build (the driver's own run does 1, 2, 4, 8)
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-base-en --no-cpu-baseline --no-host-block > $O/scale2_n2.json 2> $O/scale2_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/scale2_n2.json').read().strip().splitlines()[-1])
    print(2, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), d['device_passes_per_step'], d.get('transcripts_vs_oracle',{}).get('identical'), d['roofline']['frac'])
except Exception as e:
    print(2, 'failed', e); print(open('gpurun_out/scale2_n2.err').read()[-1500:])
PY
