#!/bin/bash
# bench lines only: tiny.en at several settings (+ the parity tests first, so a broken build never produces numbers)
mkdir -p gpurun_out; O=gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
for G in 1 2 3 4; do
  B=256
  echo "== bench tiny b$B groups $G"; WHISPER_B200_STEP_GROUPS=$G timeout 600 python bench.py --batch $B --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_tiny_b${B}_g$G.json 2> $O/bench_tiny_b${B}_g$G.err
  python - <<PY
import json
d=json.load(open("$O/bench_tiny_b${B}_g$G.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_phase_ms_per_chunk"], d["device_passes_per_step"], {k:(v["launches"],v["ms"]) for k,v in d["kernel_classes"].items()})
PY
  tail -2 $O/bench_tiny_b${B}_g$G.err
done
