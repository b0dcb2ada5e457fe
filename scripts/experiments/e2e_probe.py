# where does the end-to-end step spend its time beyond the render itself?
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import pbrt_v2_b200 as pb
from pbrt_v2_b200 import scenes
import bench
base, res, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
p = scenes.variant(base, scenes.STAGED + '/probe-%s.pbrt' % base, res, res, spp)
hs = pb.HostScene(p, ncores=bench.per_pixel_ncores(res, res))
rd = hs.render_desc()
_, _, xc, yc = hs.film_extent()
dev = torch.device("cuda", 0)
film = torch.zeros((yc, xc, 4), dtype=torch.float32, device=dev)
host_film = torch.zeros((yc, xc, 4), dtype=torch.float32).pin_memory()
stream = torch.cuda.current_stream(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T(); g2 = pb.GpuScene(hs)
    t1 = T(); g2.render_device(rd, film.data_ptr(), stream.cuda_stream)
    t2 = T(); host_film.copy_(film, non_blocking=True); stream.synchronize()
    t3 = T(); g2.close()
    t4 = T()
    print("iter %d: create %.1f ms, render %.1f ms, d2h %.1f ms, close %.1f ms" % (it, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
gs = pb.GpuScene(hs)
for it in range(3):
    t1 = T(); gs.render_device(rd, film.data_ptr(), stream.cuda_stream); t2 = T()
    print("resident render %.1f ms" % (1e3*(t2-t1)))
