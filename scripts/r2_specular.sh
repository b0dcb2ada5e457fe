# Round-2 starter: build the library WITH the specular-material code (kept out of the shipped kernels until it has run
# on a GPU), check its parity and what it costs the headline.  Run on the GPU box:  bash scripts/r2_specular.sh
set -x
mkdir -p pbrt_v2_b200/ab gpurun_out
# (built by `make variants` / __graft_entry__.build() and shipped with the snapshot; compiled here only if missing)
[ -f pbrt_v2_b200/ab/libpbrtgpu_specular.so ] || make variants
PBRTGPU_LIB=ab/libpbrtgpu_specular.so python -m pytest tests/test_gpu_specular.py tests/test_gpu_parity.py -q -m gpu \
    --deselect tests/test_gpu_parity.py::test_specular_materials_are_refused_loudly 2>&1 | tail -5 | tee gpurun_out/r2_specular_tests.txt
for lib in libpbrtgpu.so ab/libpbrtgpu_specular.so; do
  PBRTGPU_LIB=$lib python bench.py --spp 256 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['value'],1), d['kernel_ms'])" | tee -a gpurun_out/r2_specular_bench.txt
done
