#!/bin/bash
# Round-end style call: GPU parity tests, smoke, both bench arms, ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
echo "== bench (default)"; timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; tail -c 1500 $O/bench_tiny.json; tail -2 $O/bench_tiny.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 700 $O/bench_ref.json
echo "== bench b16"; timeout 600 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_tiny_b16.json 2> $O/bench_tiny_b16.err; tail -c 400 $O/bench_tiny_b16.json
echo "== bench b1"; timeout 600 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_tiny_b1.json 2> $O/bench_tiny_b1.err; tail -c 400 $O/bench_tiny_b1.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_tiny.csv \
    python tools/ncu_workload.py --batch 64 --steps 2 > $O/ncu_workload.log 2>&1
tail -3 $O/ncu_workload.log; wc -l $O/launches_tiny.csv
echo "== ncu full k_decode_step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_step -s 60 -c 2 -f -o $O/prof_decode_step \
    python tools/ncu_workload.py --batch 64 --steps 1 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
echo "== ncu full encoder kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_softmax_rows" -s 30 -c 12 -f -o $O/prof_encoder \
    python tools/ncu_workload.py --batch 16 --steps 1 > $O/ncu_full_enc.log 2>&1
tail -2 $O/ncu_full_enc.log; ls -la $O/*.ncu-rep
