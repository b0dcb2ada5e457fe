# Round-2 GPU visit 5: local-memory diet of the vertex wavefront (grid-constant kernel parameters, scalarised BxSet /
# Vertex, branch-free k_weval step): quick parity subset, A/B of the register bounds, default bench, launch list, ncu.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_specular.py -q -m gpu 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/pytest_gpu_subset.txt
timeout 900 python scripts/ab_variants.py --spp 256 --steps 2 ab/libpbrtgpu_ws5.so ab/libpbrtgpu_ws6.so ab/libpbrtgpu_we10.so \
    --env PBRTGPU_LANES=1 --env PBRTGPU_LANES=3 --env PBRTGPU_POOL=2097152 --env PBRTGPU_POOL=8388608 \
    2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_wave2.json
cut -c1-6000 gpurun_out/ab_wave2.json
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
for k in k_wshade k_weval; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
ls -la gpurun_out | tail -30
