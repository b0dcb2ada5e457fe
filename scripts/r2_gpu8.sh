# Round-2 GPU visit 8: k_wshade<1> with its slots sorted by next action (default) against thread i = slot i (PBRTGPU_WSORT=0).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_specular.py tests/test_gpu_fullscale.py::test_c2_bunny_1024sq_1024spp -q -m gpu 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/pytest_gpu_subset.txt
timeout 600 python scripts/ab_variants.py --spp 256 --steps 2 --env PBRTGPU_WSORT=0 2>gpurun_out/ab_err.txt | tail -1 > gpurun_out/ab_wave6.json
cut -c1-3000 gpurun_out/ab_wave6.json
ls -la gpurun_out | tail
