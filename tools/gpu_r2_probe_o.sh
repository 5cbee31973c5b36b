#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for w in 1 0 1 0; do WHISPER_B200_WORKER_POOL=$w timeout 300 python bench.py --steps 6 --warmup 3 --no-base-en --no-cpu-baseline --no-host-block > $O/bench_o$w.json 2> $O/bench_o$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_o$w.json').read().strip().splitlines()[-1])
print('pool $w: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'dev ms', round(d['ms_per_step'],1), 'identical', d['transcripts_vs_oracle']['identical'], '/', d['transcripts_vs_oracle']['compared'])
PY
done
