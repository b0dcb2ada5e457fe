# Round-2 GPU visit 3 (session re-entry: the earlier visits' gpurun_out/ was lost with the container): suite, default
# bench (both arms), launch list, A/B of env switches and traversal variants, ncu full of the hot kernels, sanitizer.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_reference.json
cut -c1-600 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
timeout 1200 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_VERTEX=0 --env PBRTGPU_VERTEX=3 \
    --env PBRTGPU_LANES=1 --env PBRTGPU_LANES=3 --env PBRTGPU_LANES=4 \
    ab/libpbrtgpu_top10.so ab/libpbrtgpu_sstack.so ab/libpbrtgpu_top9_sstack.so ab/libpbrtgpu_pairs.so \
    2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_vertex.json
cut -c1-8000 gpurun_out/ab_vertex.json
for k in k_trace k_vertex k_logic k_tables; do
  skip=200; [ $k = k_vertex ] && skip=2; [ $k = k_tables ] && skip=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
bash scripts/gpu_sanitize.sh
ls -la gpurun_out | tail -30
