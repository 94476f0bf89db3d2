"""development aid: pipeline timeline of CTA 0 of the engine-2 triangulation kernel (SM clocks)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam3_comments_ghr_b200 import matcher, synth
P = 4096
ctx = matcher.Context(0, stream=torch.cuda.current_stream().cuda_stream)
tc = synth.fill_geometry(synth.make_triangulation_case(9, n_pairs=P, n_feat=2000))
ks = ctx.upload_kfset(tc.kfs)
m = matcher.ORBmatcher(0.6, False, ctx)
dev = "cuda"
kf1, kf2 = torch.from_numpy(tc.kf1).to(dev), torch.from_numpy(tc.kf2).to(dev)
ep, f12 = torch.from_numpy(tc.ep).to(dev), torch.from_numpy(tc.f12).to(dev)
out = torch.empty((P, 2000), dtype=torch.int32, device=dev); nm = torch.empty(P, dtype=torch.int32, device=dev)
tl = torch.zeros((64, 8), dtype=torch.int64, device=dev)
L = matcher.load_library()
L.orbgpu_debug_triangulation_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for it in range(3):
    m.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nm.data_ptr())
L.orbgpu_debug_triangulation_timeline(ctx.handle, ctypes.c_void_p(tl.data_ptr()))
m.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nm.data_ptr())
torch.cuda.synchronize()
t = tl.cpu().numpy()
t0 = t[0, 0]
names = ["empty_ok", "full_ok", "join_done", "joined_ok", "cmp_done", "post_ready", "compared_ok", "post_done"]
print("pair " + " ".join(n.rjust(11) for n in names))
for i in range(28):
    if t[i, 0] == 0: break
    print(f"{i:4d} " + " ".join(f"{int(x - t0):11d}" for x in t[i]))
d = t[:27]
print("mean full_ok-empty_ok", (d[4:, 1] - d[4:, 0]).mean(), " join", (d[:, 2] - d[:, 1]).mean(), " joined->cmp_done", (d[:, 4] - d[:, 3]).mean(),
      " post", (d[:, 7] - d[:, 6]).mean(), " period", (d[-1, 7] - d[3, 7]) / 23.0)
