#!/bin/bash
# the very last check of the round: GPU suite + default bench line of the build as committed
mkdir -p gpurun_out/final4; O=gpurun_out/final4
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final4/bench_final.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(r['frac'],4), 'excl', round(r['exclusive']['frac'],4), 'cpu', round(d['cpu_baseline']['value'],1), 'identical', d['transcripts_vs_oracle']['identical'], d['transcripts_vs_oracle']['compared'], 'clocks', d['clocks'])
PY
