#!/bin/bash
# parity tests + the default bench line
mkdir -p gpurun_out; O=gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $O/bench_tiny.json 2> $O/bench_tiny.err
python - <<PY
import json
d=json.load(open("$O/bench_tiny.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_phase_ms_per_chunk"], d["device_passes_per_step"], {k:(v["launches"],v["ms"]) for k,v in d["kernel_classes"].items()}, d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
PY
tail -2 $O/bench_tiny.err
