for w2 in 0 1; do
  for wl in bunny-c2 killeroo-c1 synthetic-c5; do
    extra=""; [ $wl = bunny-c2 ] && extra="--spp 128"; [ $wl = synthetic-c5 ] && extra="--spp 8 --tris 2000000"
    PBRTGPU_WAVEFRONT2=$w2 python bench.py --workload $wl $extra --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wf2=$w2','$wl', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"
  done
done
