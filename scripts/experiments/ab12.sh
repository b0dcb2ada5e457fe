# lanes / scratch at the full 1024 spp of C2
for cfg in "1 40" "2 40" "4 40" "4 120" "2 120"; do
  set -- $cfg
  PBRTGPU_LANES=$1 PBRTGPU_SCRATCH_GB=$2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes=$1 scratch=$2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
done
