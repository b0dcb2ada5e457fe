set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-200 gpurun_out/bench_ours.json
python bench.py --workload killeroo-dl --steps 3 --warmup 3 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_dl.json
python bench.py --workload killeroo-c1 --steps 3 --warmup 3 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c1.json
