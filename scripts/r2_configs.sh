# BASELINE.json configs[2..4] at their real settings.  usage: r2_configs.sh c3|c4|c5
#   c3  killeroo-simple 1920x1080, 4096 spp, 1 GPU          (gpurun --gpus 1)
#   c4  anim-killeroos-moving (nlevels 3) 1920x1080, 1024 spp, 4 GPUs   (gpurun --gpus 4)
#   c5  synthetic 10 M triangles 4096x4096, 512 spp, maxdepth 8, 8 GPUs  (gpurun --gpus 8)
# One timed step each (--economy: the counting pass and the per-launch timing pass are the warm-up).
set -x
mkdir -p gpurun_out
case "$1" in
  c3) timeout 1500 python bench.py --workload killeroo-c3 --steps 1 --warmup 1 --economy 2>gpurun_out/c3_err.txt | tail -1 > gpurun_out/bench_c3.json
      cut -c1-1800 gpurun_out/bench_c3.json ;;
  c4) timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
          bench.py --gpus 4 --workload anim-c4 --steps 1 --warmup 1 --economy 2>gpurun_out/c4_err.txt | tail -1 > gpurun_out/bench_c4_n4.json
      cut -c1-1800 gpurun_out/bench_c4_n4.json ;;
  c5) timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
          bench.py --gpus 8 --workload synthetic-c5 --steps 1 --warmup 1 --economy 2>gpurun_out/c5_err.txt | tail -1 > gpurun_out/bench_c5_n8.json
      cut -c1-1800 gpurun_out/bench_c5_n8.json ;;
esac
tail -5 gpurun_out/c*_err.txt
# c3lean: configs[2] with the GPU minutes that were left at the end of round 2 (no per-launch timing pass, no e2e region)
if [ "$1" = c3lean ]; then
  timeout 170 python bench.py --workload killeroo-c3 --steps 1 --warmup 1 --economy --lean --no-cpu-baseline 2>gpurun_out/c3_err.txt | tail -1 > gpurun_out/bench_c3.json
  cut -c1-1500 gpurun_out/bench_c3.json; tail -3 gpurun_out/c3_err.txt
fi
