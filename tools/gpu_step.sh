#!/bin/bash
# GPU call focused on the decode-step kernel: parity tests, step trace, bench, launch list, one full ncu capture.
mkdir -p gpurun_out; O=gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
echo "== trace"; timeout 300 python tools/step_trace.py 16 > $O/trace_b16.txt 2>&1; grep -v "slowest" $O/trace_b16.txt | tail -38; timeout 300 python tools/step_trace.py 1 > $O/trace_b1.txt 2>&1; grep "step total" $O/trace_b1.txt
echo "== bench tiny"; timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_tiny.json 2> $O/bench_tiny.err; tail -c 2800 $O/bench_tiny.json; tail -3 $O/bench_tiny.err
if [ "$1" == "full" ]; then
echo "== bench tiny b1"; timeout 600 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_tiny_b1.json 2> $O/bench_tiny_b1.err; tail -c 1500 $O/bench_tiny_b1.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_tiny.csv \
    python tools/ncu_workload.py --batch 16 --steps 2 > $O/ncu_workload.log 2>&1
tail -3 $O/ncu_workload.log; wc -l $O/launches_tiny.csv
echo "== ncu full k_decode_step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_step -s 40 -c 2 -f -o $O/prof_decode_step \
    python tools/ncu_workload.py --batch 16 --steps 1 > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log; ls -la $O/*.ncu-rep
fi
