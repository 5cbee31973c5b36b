#!/bin/bash
# decoder pass width sweep: WHISPER_B200_DECODE_ROWS x WHISPER_B200_PASS_SPLIT on the default bench line
mkdir -p gpurun_out; O=gpurun_out
for cfg in "$@"; do
  IFS=: read rows split smax batch <<< "$cfg"
  batch=${batch:-256}; tag="r${rows}_s${split}_m${smax}_b${batch}"
  WHISPER_B200_DECODE_ROWS=$rows WHISPER_B200_PASS_SPLIT=$split WHISPER_B200_STEP_MAX_ROWS=${smax:-32} timeout 600 python bench.py --batch $batch --no-cpu-baseline --steps 4 --warmup 3 > $O/sweep_$tag.json 2> $O/sweep_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/sweep_$tag.json"))
    print("$tag", "value %.0f" % d["value"], "ms %.1f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "ms %.1f" % d["e2e"]["ms_per_step"], d["device_passes_per_step"], {k:(v["launches"],round(v["ms"],1)) for k,v in d["kernel_classes"].items()})
except Exception as e:
    print("$tag failed", e)
PY
  tail -2 $O/sweep_$tag.err
done
