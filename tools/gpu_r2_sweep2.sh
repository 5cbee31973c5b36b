#!/bin/bash
# scheduler knobs after the encoder got faster (bench value / e2e per variant, same box)
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-base-en --no-cpu-baseline --no-host-block > $O/sweep2_$name.json 2> $O/sweep2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep2_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'enc_ms', round(d['device_passes_per_step']['encoder_ms'],1), 'dec_ms', round(d['device_passes_per_step']['decoder_ms'],1), 'dec passes', d['device_passes_per_step']['decoder'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['exclusive']['frac'],3))
except Exception as e:
    print('$name failed', e)
PY
}
run base X=1
run enc64 WHISPER_B200_ENC_BATCH=64
run enc48 WHISPER_B200_ENC_BATCH=48
run min448 WHISPER_B200_RUN_MIN_ROWS=448
run min256 WHISPER_B200_RUN_MIN_ROWS=256
run base2 X=1
