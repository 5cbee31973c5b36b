#!/bin/bash
# weak-scaling check on one box: N = 1 and N = 8 (the driver's own run does 1, 2, 4, 8)
mkdir -p gpurun_out; O=gpurun_out
nproc > $O/nproc_scale.txt
for n in 1 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-base-en --no-cpu-baseline > $O/scale_n$n.json 2> $O/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-base-en --no-cpu-baseline > $O/scale_n$n.json 2> $O/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
    print($n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), d['device_passes_per_step'], d.get('transcripts_vs_oracle',{}).get('identical'))
except Exception as e:
    print($n, 'failed', e); print(open('gpurun_out/scale_n$n.err').read()[-1500:])
PY
done
echo "== single process, 8 GPUs through whisper_b200_init_multi"; timeout 600 python tools/multi_device_bench.py 8 512 3 2>&1 | tail -2
echo "== multi-device test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_device 2>&1 | tail -3
