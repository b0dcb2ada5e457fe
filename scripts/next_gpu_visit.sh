# The first hardware run of everything written after round 2's last GPU visit.  One B200:
#     gpurun --timeout 1500 -- 'bash scripts/next_gpu_visit.sh > gpurun_out/next_visit.log 2>&1'
# 1. the shipped kernels are the last visit's (scripts/sass_fingerprint.sh): the suite must be what it was, plus the late
#    front-end features' first renders (tests/test_zz_gpu_late.py: object instances, named materials, constant textures)
# 2. the f4 library = metal + infinite lights / radiance maps + pixel filters + MatInverse's zero-initialised pivot arrays
#    (make variants VARIANT_NAMES=f4 before the visit: the .so travels): parity of the old suite under it (nothing may move),
#    then the gated features' own tests
# 3. what the gates cost: throughput of f4 against the shipped library on configs[1] and on the animated scene
# 4. compute-sanitizer over the f4 kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
if [ -f pbrt_v2_b200/ab/libpbrtgpu_f4.so ]; then
    PBRTGPU_LIB=ab/libpbrtgpu_f4.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_specular.py tests/test_gpu_refusals.py tests/test_zz_gpu_late.py -m gpu -q 2>&1 | tail -8
    python scripts/ab_variants.py ab/libpbrtgpu_f4.so --workload bunny-c2 --spp 128 > gpurun_out/ab_f4_c2.json 2> gpurun_out/ab_f4_c2.err || true
    python scripts/ab_variants.py ab/libpbrtgpu_f4.so --workload anim-c4 --spp 64 > gpurun_out/ab_f4_c4.json 2> gpurun_out/ab_f4_c4.err || true
    PBRTGPU_LIB=ab/libpbrtgpu_f4.so bash scripts/gpu_sanitize.sh > gpurun_out/sanitize_f4.log 2>&1 || true
else
    echo "no f4 library: run  make variants VARIANT_NAMES=f4  first"
fi
