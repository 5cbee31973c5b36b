#!/bin/bash
# Round-2 evidence run: GPU parity tests, smoke, both bench arms, single-chunk line, micro-benchmarks, ncu launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log | cut -c1-300
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','e2e_host_block','base_en_b8_beam5','cpu_baseline','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:600])
PY
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
echo "== bench b1"; timeout 300 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-base-en --no-host-block > $O/bench_tiny_b1.json 2> $O/bench_tiny_b1.err; tail -c 300 $O/bench_tiny_b1.json
echo "== realtime"; timeout 300 python tools/realtime_latency.py --reference > $O/realtime_latency.json 2> $O/realtime.err; cat $O/realtime_latency.json | cut -c1-400
echo "== micro-benchmarks"; timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench.md 2>&1; timeout 300 python tools/attn_enc_bench.py 6 16 1500 > $O/attn_enc_bench.md 2>&1; cat $O/gemm_enc_bench.md $O/attn_enc_bench.md
echo "== ncu launch list (128-chunk step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/launches_b128.csv python tools/ncu_workload.py --batch 128 --steps 2 > $O/ncu_workload.log 2>&1
python tools/launch_summary.py $O/launches_b128.csv > $O/launches_b128.md; head -24 $O/launches_b128.md
python tools/launch_summary.py $O/launches_b128.csv --by-grid > $O/launches_b128_grid.md
gzip -f $O/launches_b128.csv
echo "== ncu full: encoder tensor-core kernels, one launch each"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_enc|k_attn_enc" -f -o $O/prof_encoder python tools/ncu_kernels.py > $O/ncu_full_enc.log 2>&1; tail -1 $O/ncu_full_enc.log
echo "== ncu full: wide decoder step + spectrogram kernels (64 chunks)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_attention|k_gemm_tc<32|k_sample_greedy|k_logmel_frames|k_layernorm_vec|k_run_" -s 700 -c 30 -f -o $O/prof_decoder \
    python tools/ncu_workload.py --batch 64 --steps 2 > $O/ncu_full_dec.log 2>&1; tail -1 $O/ncu_full_dec.log
echo "== ncu full: decode-step kernel (single sequence)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_step -s 20 -c 1 -f -o $O/prof_decode_step python tools/ncu_workload.py --batch 1 --steps 2 > $O/ncu_full_step.log 2>&1; tail -1 $O/ncu_full_step.log
# the reports stay on the box (gpurun brings back at most 64 MiB): raw-metric tables and the per-instruction view of the hottest kernels come home instead
for r in prof_encoder prof_decoder prof_decode_step; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.summary.md 2>/dev/null
done
ncu -i $O/prof_encoder.ncu-rep --page source --csv --kernel-id ::regex:k_attn_enc:1 2>/dev/null | gzip > $O/prof_attn_enc.source.csv.gz
ncu -i $O/prof_encoder.ncu-rep --page source --csv --kernel-id ::regex:k_gemm_enc:3 2>/dev/null | gzip > $O/prof_gemm_enc_gelu.source.csv.gz
ls -la $O/*.ncu-rep; rm -f $O/*.ncu-rep; gzip -f $O/*.raw.csv; du -sh $O; cat $O/prof_encoder.summary.md
