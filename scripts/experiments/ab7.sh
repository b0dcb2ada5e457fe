for wl in bunny-c2 killeroo-c1; do
  extra=""; [ $wl = bunny-c2 ] && extra="--spp 128"
  python bench.py --workload $wl $extra --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
done
