#!/bin/bash
# Round 2, call 1: the round-1 build with the new parity tests (fallback, 512-chunk golden, base.en / multilingual / small shapes) + baseline bench.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -40 $O/pytest_gpu.log
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; tail -c 1200 $O/bench_tiny.json; tail -4 $O/bench_tiny.err
