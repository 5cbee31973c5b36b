#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_l.log 2>&1; tail -3 $O/pytest_probe_l.log
timeout 900 python bench.py --no-host-block > $O/bench_l.json 2> $O/bench_l.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_l.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', r['frac'], 'excl', r['exclusive']['frac'], 't_enc', r['t_encoder_ms_per_step'], r['exclusive']['t_encoder_ms'])
print('classes', json.dumps(d['kernel_classes']))
print('base', json.dumps(d['base_en_b8_beam5'])[:700])
print('tv', d['transcripts_vs_oracle']['identical'], d['transcripts_vs_oracle']['compared'])
PY
