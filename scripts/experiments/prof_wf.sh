mkdir -p gpurun_out
cat > /tmp/prof_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import pbrt_v2_b200 as pb
from pbrt_v2_b200 import scenes
base, res, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
p = scenes.variant(base, scenes.STAGED + '/prof-%s.pbrt' % base, res, res, spp)
hs = pb.HostScene(p, ncores=res*res//32)
gs = pb.GpuScene(hs)
for i in range(2):
    px, st = gs.render()
    print(st.as_dict())
PY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 1200 --csv --log-file gpurun_out/wf_launches.csv python /tmp/prof_run.py $1 $2 $3 > gpurun_out/wf.log 2>&1
tail -2 gpurun_out/wf.log
