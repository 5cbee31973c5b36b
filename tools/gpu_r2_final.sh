#!/bin/bash
# Round-2 evidence run: GPU parity tests, smoke, both bench arms, single-chunk line, ncu launch list, ncu --set full captures of the top kernels.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log; grep "rel-L2\|near-tie" $O/pytest_gpu.log | head
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log | cut -c1-300
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_tiny.json 2> $O/bench_tiny.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tiny.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','e2e_host_block','base_en_b8_beam5','cpu_baseline','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:500])
PY
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
echo "== bench b1"; timeout 300 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-base-en --no-host-block > $O/bench_tiny_b1.json 2> $O/bench_tiny_b1.err; tail -c 300 $O/bench_tiny_b1.json
echo "== ncu launch list"; bash tools/gpu_r2_ncu_list.sh 2>&1 | head -26
echo "== ncu full: encoder pass kernels (16 chunks)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_enc|k_attn_enc|k_logmel_frames|k_gemm_tc_persistent" -s 27 -c 27 -f -o $O/prof_encoder \
    python tools/ncu_workload.py --batch 16 --steps 2 > $O/ncu_full_enc.log 2>&1; tail -1 $O/ncu_full_enc.log
echo "== ncu full: wide decoder step kernels (256 chunks)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_attention|k_gemm_tc<32|k_sample_greedy|k_run_|k_layernorm_vec" -s 2600 -c 48 -f -o $O/prof_decoder \
    python tools/ncu_workload.py --batch 256 --steps 2 > $O/ncu_full_dec.log 2>&1; tail -1 $O/ncu_full_dec.log
echo "== ncu full: decode-step kernel (single sequence)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_step -s 20 -c 2 -f -o $O/prof_decode_step \
    python tools/ncu_workload.py --batch 1 --steps 2 > $O/ncu_full_step.log 2>&1; tail -1 $O/ncu_full_step.log
# the reports stay on the box (gpurun brings back at most 64 MiB): raw-metric tables and the per-instruction view of the hottest kernels come home instead
for r in prof_encoder prof_decoder prof_decode_step; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.summary.md 2>/dev/null
done
ncu -i $O/prof_encoder.ncu-rep --page source --csv --kernel-id ::regex:k_attn_enc:1 2>/dev/null | gzip > $O/prof_attn_enc.source.csv.gz
ncu -i $O/prof_encoder.ncu-rep --page source --csv --kernel-id ::regex:k_gemm_enc:2 2>/dev/null | gzip > $O/prof_gemm_enc.source.csv.gz
ncu -i $O/prof_decoder.ncu-rep --page source --csv --kernel-id ::regex:k_decode_attention:2 2>/dev/null | gzip > $O/prof_decode_attention.source.csv.gz
ls -la $O/*.ncu-rep; rm -f $O/*.ncu-rep; gzip -f $O/*.raw.csv; du -sh $O
