# Round-2 starter for the scene-preparation kernels: device BVH build and device LoopSubdiv (CPU-verified in round 1, tests/test_bvhbuild_on_host.py):
# first runs on hardware.  On the GPU box:  bash scripts/r2_bvhbuild.sh
set -x
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_bvhbuild.py tests/test_zz_gpu_loopsubdiv.py -q -m gpu -s 2>&1 | tail -15 | tee gpurun_out/r2_bvhbuild_tests.txt
timeout 900 python scripts/bench_bvhbuild.py --n 10000000 2>&1 | tee gpurun_out/r2_bvhbuild_bench.txt
# launch list of one 1 M-primitive build (which kernels dominate a level)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_bvhbuild_launches.csv \
    python scripts/bench_bvhbuild.py --n 1000000 --reps 1 > /dev/null 2>&1
