# Two ranks on two GPUs through the library's own NCCL communicator (pbrtgpu_render_sharded): C2 at 64 spp, and the CLI.
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --spp 64 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/n2_err.txt | tail -1 > gpurun_out/bench_n2_dev.json
cut -c1-1200 gpurun_out/bench_n2_dev.json; tail -5 gpurun_out/n2_err.txt
timeout 120 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4
