# Round-2 GPU visit 6: phase 2 as a chain wavefront (k_wshade<2>), k_wshade at 6 blocks/SM, leaf tests in batches.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_specular.py tests/test_gpu_fullscale.py::test_c2_bunny_1024sq_1024spp tests/test_gpu_fullscale.py::test_c5_synthetic_4096sq_512spp -q -m gpu 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/pytest_gpu_subset.txt
timeout 900 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_WAVE=1 ab/libpbrtgpu_ws6.so ab/libpbrtgpu_ws6_lb8.so \
    ab/libpbrtgpu_ws6_lb12.so ab/libpbrtgpu_ws6_lb20.so 2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_wave3.json
cut -c1-6000 gpurun_out/ab_wave3.json
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-2500 gpurun_out/bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
ls -la gpurun_out | tail -30
