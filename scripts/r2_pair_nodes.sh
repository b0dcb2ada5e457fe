# Round-2 A/B: the sibling-pair node layout (-DPBRT_PAIR_NODES=1, device/dscene.cuh), CPU-verified in round 1
# (tests/test_device_on_host.py::test_pair_node_walk_on_host) but never timed.  On the GPU box:  bash scripts/r2_pair_nodes.sh
set -x
mkdir -p pbrt_v2_b200/ab gpurun_out
# (built by `make variants` / __graft_entry__.build() and shipped with the snapshot; compiled here only if missing)
[ -f pbrt_v2_b200/ab/libpbrtgpu_pairs.so ] || make variants
# parity first: the whole GPU suite against the variant library
PBRTGPU_LIB=ab/libpbrtgpu_pairs.so python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_pairs_tests.txt
# then the headline, the denser mesh and the animated scene (nested walks), both libraries back to back
for wl in bunny-c2 killeroo-c1 anim-c4; do
  for lib in libpbrtgpu.so ab/libpbrtgpu_pairs.so; do
    PBRTGPU_LIB=$lib python bench.py --workload $wl --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', '$lib', round(d['value'],1), d.get('kernel_ms'))" | tee -a gpurun_out/r2_pairs_bench.txt
  done
done
