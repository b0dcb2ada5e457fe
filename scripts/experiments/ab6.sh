# synthetic C5 (proper task grid): tail megakernel vs tail wavefront, 2M and 10M triangles
for w2 in 0 1; do
  PBRTGPU_WAVEFRONT2=$w2 python bench.py --workload synthetic-c5 --spp 32 --tris 2000000 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wf2=$w2 2M', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'], d['roofline'])"
done
PBRTGPU_WAVEFRONT2=0 python bench.py --workload synthetic-c5 --spp 32 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wf2=0 10M', round(d['value'],1), {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['gpu_launches'], d['roofline'])"
