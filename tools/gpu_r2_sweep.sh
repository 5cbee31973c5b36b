#!/bin/bash
# tuning sweep of the batch scheduler knobs (bench value / e2e per variant) + the large-v3 shape test + realtime latency
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-base-en --no-cpu-baseline --no-host-block > $O/sweep_$name.json 2> $O/sweep_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'enc_ms', round(d['device_passes_per_step']['encoder_ms'],1), 'dec_ms', round(d['device_passes_per_step']['decoder_ms'],1), 'dec passes', d['device_passes_per_step']['decoder'])
except Exception as e:
    print('$name failed', e)
PY
}
run base X=1
run enc32 WHISPER_B200_ENC_BATCH=32
run min128 WHISPER_B200_RUN_MIN_ROWS=128
run min320 WHISPER_B200_RUN_MIN_ROWS=320
run depth2 WHISPER_B200_RUN_DEPTH=2
run enc32min320 WHISPER_B200_ENC_BATCH=32 WHISPER_B200_RUN_MIN_ROWS=320
echo "== large-v3 shapes"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k large_v3 2>&1 | tail -5
echo "== realtime"; timeout 300 python tools/realtime_latency.py --reference > $O/realtime_latency.json 2> $O/realtime.err; cat $O/realtime_latency.json | cut -c1-700
