# Round-2 GPU visit 1: full GPU suite (incl. BASELINE-scale sub-range parity), vertex-synchronous kernels vs the round-1
# kernels (A/B), ncu captures of k_tables / k_vertex, compute-sanitizer.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 900 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_VERTEX=0 --env PBRTGPU_VERTEX=1 --env PBRTGPU_VERTEX=2 \
    --env PBRTGPU_TAIL_SORT=0 --env PBRTGPU_VERTEX_BLOCKS=3 2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_vertex.json
cut -c1-3000 gpurun_out/ab_vertex.json
timeout 600 python bench.py --steps 2 --warmup 2 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-1500 gpurun_out/bench_ours.json
for k in k_tables k_vertex; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -o gpurun_out/full_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
bash scripts/gpu_sanitize.sh
ls -la gpurun_out | tail -30
