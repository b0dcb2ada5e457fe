# A/B of kernel build variants: prints Msamples/s and per-kernel ms for the 64-spp development run
for lib in "$@"; do
  for wl in bunny-c2 killeroo-c1; do
    extra=""; [ $wl = bunny-c2 ] && extra="--spp 64"
    PBRTGPU_LIB=$lib python bench.py --workload $wl $extra --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib','$wl', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()})"
  done
done
