for lanes in 1 2; do
  PBRTGPU_LANES=$lanes python bench.py --workload anim-c4 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('anim-c4 lanes=$lanes', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
done
