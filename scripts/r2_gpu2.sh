# Round-2 GPU visit 2: suite, A/B of the phase-1/phase-2 kernel choices and of the north_star traversal variants
# (TMA-staged BVH top, shared-memory stack, sibling pairs), default bench, ncu of the reworked k_trace and k_vertex<2>, sanitizer.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 1200 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_VERTEX=0 --env PBRTGPU_VERTEX=3 \
    --env PBRTGPU_TAIL_SORT=0 --env PBRTGPU_VERTEX_BLOCKS=3 --env PBRTGPU_LANES=1 --env PBRTGPU_LANES=3 \
    ab/libpbrtgpu_top10.so ab/libpbrtgpu_sstack.so ab/libpbrtgpu_top9_sstack.so ab/libpbrtgpu_pairs.so \
    2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_vertex.json
cut -c1-6000 gpurun_out/ab_vertex.json
timeout 600 python bench.py --steps 2 --warmup 2 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-1500 gpurun_out/bench_ours.json
for k in k_trace k_vertex k_logic; do
  skip=200; [ $k = k_vertex ] && skip=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
bash scripts/gpu_sanitize.sh
ls -la gpurun_out | tail -12
