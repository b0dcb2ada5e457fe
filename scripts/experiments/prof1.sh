set -e
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
ncu --set full --clock-control none --import-source on -k regex:k_path -s 2 -c 2 -o gpurun_out/prof_path2_bunny python /tmp/prof_run.py bunny 256 64 > gpurun_out/prof1.log 2>&1
tail -5 gpurun_out/prof1.log
