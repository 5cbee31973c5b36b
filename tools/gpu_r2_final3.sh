#!/bin/bash
# last refresh of the round: smoke + the default bench line of the build as committed
mkdir -p gpurun_out/final3; O=gpurun_out/final3
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log | cut -c1-200
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final3/bench_final.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(r['frac'],4), 'excl', round(r['exclusive']['frac'],4), 'cpu', round(d['cpu_baseline']['value'],1), 'identical', d['transcripts_vs_oracle'], 'clocks', d['clocks'])
print('base', json.dumps(d['base_en_b8_beam5'])[:400])
PY
