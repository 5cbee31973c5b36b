export WHISPER_B200_HOST_TRACE=1
for cfg in 384:1024:384 384:1024:512 384:768:512 512:1024:640; do
  IFS=: read rows batch workers <<< "$cfg"
  WHISPER_B200_MAX_WORKERS=$workers bash tools/gpu_sweep.sh $rows:2:32:$batch 2>&1 | grep -v "^full_batch" ; grep full_batch gpurun_out/sweep_r${rows}_s2_m32_b${batch}.err | head -1
done
