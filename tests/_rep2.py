import sys, numpy as np, ctypes as C, subprocess, os
sys.path.insert(0,'..')
import pbrt_v2_b200 as pb
import helpers, oracle_api as oa
a = C.CDLL("/tmp/doh_a.so")
hs = helpers.variant_scene("dof")[0]
rd = hs.render_desc(); px = hs.new_pixels()
a.doh_render.argtypes=[C.c_void_p, C.POINTER(pb.RenderDesc), C.c_void_p]
a.doh_render(hs.desc_ptr, C.byref(rd), px.ctypes.data)
b = C.CDLL("/tmp/doh_b.so")
b.doh_render.argtypes=[C.c_void_p, C.POINTER(pb.RenderDesc), C.c_void_p]
hs2 = helpers.variant_scene("path-glass-mirror")[0]
rd2 = hs2.render_desc(); px2 = hs2.new_pixels()
print("CALL", flush=True)
b.doh_render(hs2.desc_ptr, C.byref(rd2), px2.ctypes.data); print("b rendered", flush=True)
