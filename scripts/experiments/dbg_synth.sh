# iteration/queue trace of the phase-1 wavefront on the synthetic scene
PBRTGPU_WAVEFRONT2=0 PBRTGPU_TRACE_ITERS=1 python bench.py --workload synthetic-c5 --spp 8 --tris 2000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/dbg_synth.json 2> gpurun_out/dbg_synth.err
grep -c pbrtgpu gpurun_out/dbg_synth.err
awk 'NR%40==1' gpurun_out/dbg_synth.err | head -60
tail -1 gpurun_out/dbg_synth.json | cut -c1-600
