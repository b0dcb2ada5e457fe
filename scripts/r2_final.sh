# Round-2 final single-GPU visit: the full GPU suite, the default bench (both arms), the launch list of the bench command,
# ncu --set full of the vertex wavefront's kernels, compute-sanitizer over small renders of every kernel family.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_reference.json
cut -c1-500 gpurun_out/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
for k in k_wshade k_wtrace k_weval; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
for tool in memcheck racecheck; do
  SANITIZE_RES=10 PBRTGPU_SCRATCH_GB=2 timeout 240 compute-sanitizer --tool $tool --print-limit 20 \
      python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.txt
  tail -3 gpurun_out/sanitize_$tool.txt
done
ls -la gpurun_out | tail -30
