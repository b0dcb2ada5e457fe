mkdir -p gpurun_out
cat > /tmp/prof_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import pbrt_v2_b200 as pb
n, res, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hs = pb.HostScene.synthetic(n, res, res, spp, maxdepth=8, ncores=res*res//32)
gs = pb.GpuScene(hs)
for i in range(2):
    px, st = gs.render()
    print(st.as_dict())
PY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -c 4000 --csv --log-file gpurun_out/wf_launches.csv python /tmp/prof_run.py $1 $2 $3 > gpurun_out/wf.log 2>&1
tail -2 gpurun_out/wf.log
