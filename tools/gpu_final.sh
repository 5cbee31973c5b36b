#!/bin/bash
# Round-end style call: GPU parity tests, smoke, both bench arms, ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; tail -c 1500 $O/bench_tiny.json; tail -2 $O/bench_tiny.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 700 $O/bench_ref.json
for b in 256 16 1; do
echo "== bench b$b"; timeout 600 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_tiny_b$b.json 2> $O/bench_tiny_b$b.err; tail -c 300 $O/bench_tiny_b$b.json
done
echo "== bench base.en b8"; timeout 600 python bench.py --model base.en --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_base.json 2> $O/bench_base.err; tail -c 300 $O/bench_base.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_tiny.csv \
    python tools/ncu_workload.py --batch 256 --steps 1 > $O/ncu_workload.log 2>&1
tail -3 $O/ncu_workload.log; python tools/launch_summary.py $O/launches_tiny.csv > $O/launches_tiny.md; head -16 $O/launches_tiny.md
echo "== ncu full: decoder attention of a 256-row pass (4 self + 4 cross launches)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_attention -s 400 -c 8 -f -o $O/prof_dec_attn \
    python tools/ncu_workload.py --batch 256 --steps 1 > $O/ncu_full_dec_attn.log 2>&1
tail -1 $O/ncu_full_dec_attn.log
echo "== ncu full: fused encoder attention + encoder GEMMs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_attn_enc|k_gemm_tc" -s 4 -c 12 -f -o $O/prof_encoder \
    python tools/ncu_workload.py --batch 16 --steps 1 > $O/ncu_full_enc.log 2>&1
tail -1 $O/ncu_full_enc.log
echo "== ncu full: decode-step kernel (passes of up to 32 rows)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_step -s 20 -c 2 -f -o $O/prof_decode_step \
    python tools/ncu_workload.py --batch 16 --steps 1 > $O/ncu_full_step.log 2>&1
tail -1 $O/ncu_full_step.log; ls -la $O/*.ncu-rep
