# the bench lines of the round (after the lanes default changed to 2) -> gpurun_out/
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-300 gpurun_out/bench_ours.json
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_reference.json
python bench.py --workload killeroo-c1 --steps 3 --warmup 3 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c1.json
python bench.py --workload killeroo-dl --steps 3 --warmup 3 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_dl.json
python bench.py --workload killeroo-dl --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_dl_reference.json
python bench.py --workload killeroo-c3 --spp 256 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c3_dev.json
python bench.py --workload anim-c4 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c4_dev.json
python bench.py --workload synthetic-c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c5_dev.json
for f in c1 dl dl_reference c3_dev c4_dev c5_dev; do cut -c1-200 gpurun_out/bench_$f.json; done
tail -3 gpurun_out/bench_err.txt
