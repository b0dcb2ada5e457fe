# One GPU visit: parity tests, the default bench (both arms), the other configs at reduced spp, the ncu launch
# list of the bench command and full ncu captures of the hot kernels.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_ours.json
cut -c1-700 gpurun_out/bench_ours.json
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_reference.json
cut -c1-400 gpurun_out/bench_reference.json
# the other BASELINE configs (development runs at reduced spp are flagged INVALID in their own lines)
python bench.py --workload killeroo-c1 --steps 3 --warmup 3 --no-extras 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c1.json
python bench.py --workload killeroo-c3 --spp 256 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c3_dev.json
python bench.py --workload anim-c4 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c4_dev.json
python bench.py --workload synthetic-c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_c5_dev.json
for f in c1 c3_dev c4_dev c5_dev; do cut -c1-300 gpurun_out/bench_$f.json; done
# launch list of the bench command (serialised, cold-cache per-launch times: the kernels' SHARES are what counts)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_ gpurun_out/launches.csv
# full captures: one mid-render launch of each hot kernel (bench workload at 128 spp)
for k in k_trace k_logic k_eval k_path k_tables; do
  skip=200; [ $k = k_path ] && skip=2; [ $k = k_tables ] && skip=2
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/full_$k \
      python bench.py --spp 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/full_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/full_${k}_src.csv 2>/dev/null
  [ $k = k_trace ] || rm -f gpurun_out/full_$k.ncu-rep   # gpurun_out/ comes back capped at 64 MiB: keep the CSV exports
done
rm -f gpurun_out/prof_*.ncu-rep
ls -la gpurun_out | tail -30
