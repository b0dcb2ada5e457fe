# compute-sanitizer over small renders of every kernel family.
# Outputs -> gpurun_out/sanitize_*.txt (copied to profiles/ by hand when clean).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  SANITIZE_RES=16 PBRTGPU_SCRATCH_GB=2 timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
      python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.txt
  tail -4 gpurun_out/sanitize_$tool.txt
done
