for lanes in 1 2 3 4; do
  for wl in bunny-c2 killeroo-c1; do
    extra=""; [ $wl = bunny-c2 ] && extra="--spp 128"
    PBRTGPU_LANES=$lanes python bench.py --workload $wl $extra --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes=$lanes','$wl', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
  done
done
for lanes in 1 2 4; do
  PBRTGPU_LANES=$lanes python bench.py --workload synthetic-c5 --spp 32 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes=$lanes','synthetic-c5 10M', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
done
