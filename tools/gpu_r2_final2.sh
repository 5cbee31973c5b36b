#!/bin/bash
# Round-2 evidence refresh on the final build: GPU parity tests, smoke, both bench arms, micro-benchmarks, ncu launch list, ncu --set full of the encoder kernels and of a 256-row decoder attention pass.
mkdir -p gpurun_out; O=gpurun_out/final2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $O/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log | cut -c1-300
echo "== bench (default)"; WHISPER_B200_HOST_TRACE=1 timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2/bench_final.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','roofline','device_passes_per_step','transcripts_vs_oracle','e2e_host_block','base_en_b8_beam5','cpu_baseline','clocks','gpu_launches','kernel_classes'):
    print(k, json.dumps(d.get(k))[:700])
PY
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
echo "== micro-benchmarks"; timeout 300 python tools/gemm_enc_bench.py 384 16 > $O/gemm_enc_bench.md 2>&1; timeout 300 python tools/gemm_enc_bench.py 384 64 > $O/gemm_enc_bench_64chunks.md 2>&1; ATTN_VARIANTS=0,1,5 timeout 300 python tools/attn_enc_bench.py 6 64 1500 > $O/attn_enc_bench_64chunks.md 2>&1; cat $O/gemm_enc_bench.md $O/gemm_enc_bench_64chunks.md $O/attn_enc_bench_64chunks.md
echo "== ncu full: encoder tensor-core kernels, one launch each"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_enc|k_attn_enc" -f -o $O/prof_encoder python tools/ncu_kernels.py > $O/ncu_full_enc.log 2>&1; tail -1 $O/ncu_full_enc.log
echo "== ncu full: decoder attention of a 256-row pass"
timeout 600 ncu --set full --clock-control none -k regex:k_decode_attention -s 16 -c 4 -f -o $O/prof_dec_attn_256 python tools/ncu_workload.py --batch 256 --steps 1 > $O/ncu_full_dec256.log 2>&1; tail -1 $O/ncu_full_dec256.log
for r in prof_encoder prof_dec_attn_256; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.summary.md 2>/dev/null
done
rm -f $O/*.ncu-rep; gzip -f $O/*.raw.csv
echo "== ncu launch list (128-chunk step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $O/launches_b128.csv python tools/ncu_workload.py --batch 128 --steps 2 > $O/ncu_workload.log 2>&1
python tools/launch_summary.py $O/launches_b128.csv > $O/launches_b128.md; head -24 $O/launches_b128.md
python tools/launch_summary.py $O/launches_b128.csv --by-grid > $O/launches_b128_grid.md
gzip -f $O/launches_b128.csv
du -sh $O; cat $O/prof_encoder.summary.md $O/prof_dec_attn_256.summary.md | cut -d'|' -f2-13
