for v in trace8 trace7only eval10; do
  lib="libpbrtgpu.so"; [ $v != base ] && lib="ab/libpbrtgpu_$v.so"
  PBRTGPU_LIB=$lib python bench.py --spp 256 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, 'ktrace', round(d['roofline']['k_trace_ms_per_step'],1), 'tail', round(d['roofline']['k_path2_ms_per_step'],1), 'serial', round(d['roofline']['serial_step_ms'],1))"
done
