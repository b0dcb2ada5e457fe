run() { PBRTGPU_WAVEFRONT=$1 PBRTGPU_POOL=$2 python bench.py --workload $3 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wf=$1 pool=$2','$3', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'])"; }
run 0 0 bunny-c2
run 1 1048576 bunny-c2
run 1 4194304 bunny-c2
run 1 4194304 killeroo-c1
