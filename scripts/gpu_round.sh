# One GPU visit: parity tests, the default bench (both arms), the ncu launch list of the same
# command, and a full ncu capture of the path kernel on a reduced image.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cat gpurun_out/bench_ours.json | cut -c1-600
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_reference.json
cat gpurun_out/bench_reference.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
bash scripts/prof1.sh
