#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_probe_n.log 2>&1; tail -3 $O/pytest_probe_n.log
for w in 1 0; do WHISPER_B200_SELF_ATTN_WARP=$w timeout 300 python bench.py --steps 4 --warmup 3 --no-base-en --no-cpu-baseline --no-host-block > $O/bench_n$w.json 2> $O/bench_n$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$w.json').read().strip().splitlines()[-1])
print('warp path $w: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dec_ms', round(d['device_passes_per_step']['decoder_ms'],1), 'dec_attn', d['kernel_classes']['dec_attn'], 'identical', d['transcripts_vs_oracle']['identical'], '/', d['transcripts_vs_oracle']['compared'])
PY
done
