# usage: prof_k.sh <kernel-regex> <skip> <scene> <res> <spp> <out>
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
px, st = gs.render()
print(st.as_dict())
PY
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$6 python /tmp/prof_run.py $3 $4 $5 > gpurun_out/prof_k.log 2>&1
tail -2 gpurun_out/prof_k.log
