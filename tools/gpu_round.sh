#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, the ncu launch list.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
echo "== bench tiny"; timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_tiny.json 2> $O/bench_tiny.err; tail -c 3000 $O/bench_tiny.json
echo "== bench base"; timeout 600 python bench.py --model base.en --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_base.json 2> $O/bench_base.err; tail -c 2500 $O/bench_base.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 1500 $O/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_tiny.csv \
    python tools/ncu_workload.py --batch 4 --steps 2 > $O/ncu_workload.log 2>&1
tail -3 $O/ncu_workload.log; wc -l $O/launches_tiny.csv
