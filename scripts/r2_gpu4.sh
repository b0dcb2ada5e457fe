# Round-2 GPU visit 4: the vertex wavefront (kwave.cuh) as the default phase 1 -- full GPU suite (no -x: every failure
# is wanted), A/B against the request-granular wavefront and the traversal variants, default bench, launch list.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 | cut -c1-900 | tee gpurun_out/pytest_gpu.txt
timeout 900 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_WAVE=0 \
    ab/libpbrtgpu_ww.so ab/libpbrtgpu_ww_sstack.so ab/libpbrtgpu_sstack.so ab/libpbrtgpu_top9_sstack.so ab/libpbrtgpu_pairs.so \
    2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_wave.json
cut -c1-6000 gpurun_out/ab_wave.json
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
for k in k_wshade k_weval k_wtrace; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
ls -la gpurun_out | tail -30
