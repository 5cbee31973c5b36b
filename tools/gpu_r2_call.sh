#!/bin/bash
# Round 2 standard call: GPU parity tests, smoke, both bench arms.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; echo "rc=$?"; tail -c 2500 $O/bench_tiny.json; grep -v "^full_batch: process\|^device:" $O/bench_tiny.err | tail -8
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
