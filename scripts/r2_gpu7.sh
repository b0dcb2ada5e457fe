# Round-2 GPU visit 7: chain wavefront with compacted live lists and live-sized grids; transform cache on the animated scene.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_specular.py tests/test_gpu_fullscale.py::test_c4_anim_nlevels3_1080p_1024spp -q -m gpu 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/pytest_gpu_subset.txt
timeout 900 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_WAVE=1 ab/libpbrtgpu_ws6.so --env PBRTGPU_SCRATCH_GB=100 \
    --env PBRTGPU_LANES=1 2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_wave4.json
cut -c1-6000 gpurun_out/ab_wave4.json
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 600 python bench.py --workload anim-c4 --spp 64 --steps 1 --warmup 1 --economy --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c4_dev.json
cut -c1-1500 gpurun_out/bench_c4_dev.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
ls -la gpurun_out | tail -30
