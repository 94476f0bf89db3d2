"""development aid: pipeline timeline of CTA 0 of triangulation_stream_kernel (SM clocks relative to the producer's first stamp).
PAIRS=<n> selects the batch size (default 4096 -> ~28 pairs per CTA), MODE=dense|compact the output form."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam3_comments_ghr_b200 import matcher, synth
P = int(os.environ.get("PAIRS", 4096))
MODE = os.environ.get("MODE", "dense")
dev = torch.device("cuda", 0)
ctx = matcher.Context(0, stream=torch.cuda.current_stream().cuda_stream)
tc = synth.fill_geometry(synth.make_triangulation_case(9, n_pairs=P, n_feat=2000))
ks = ctx.upload_kfset(tc.kfs)
m = matcher.ORBmatcher(0.6, False, ctx)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
kf1, kf2, ep, f12 = t(tc.kf1), t(tc.kf2), t(tc.ep), t(tc.f12)
out = torch.empty((P, 2000), dtype=torch.int32, device=dev); nm = torch.empty(P, dtype=torch.int32, device=dev)
flags = torch.zeros(8, dtype=torch.int32, device=dev)
state = torch.tensor([1, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
g = matcher.tri_gather_struct(0, [out.data_ptr()], [nm.data_ptr()], [flags.data_ptr()], state.data_ptr(), state.data_ptr() + 16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def step():
    if MODE == "compact":
        m.SearchForTriangulation_gather_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), g, 0)
    else:
        m.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nm.data_ptr())
n_my = (P + 147) // 148
tl = torch.zeros((max(n_my, 1) + 1, 16), dtype=torch.int64, device=dev)
L = matcher.load_library()
L.orbgpu_debug_triangulation_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for it in range(3):
    step()
L.orbgpu_debug_triangulation_timeline(ctx.handle, ctypes.c_void_p(tl.data_ptr()))
flush.fill_(1)
step()
torch.cuda.synchronize()
L.orbgpu_debug_triangulation_timeline(ctx.handle, ctypes.c_void_p(0))
tt = tl.cpu().numpy()[:n_my]
t0 = tt[0, 0]
names = ["empty_ok", "full_ok", "join_done", "joined_ok", "cmp_done", "post_ready", "compared_ok", "post_done", "output", "placed", "gated"]
print(f"# {MODE} output, {P} pairs, CTA 0 handles {n_my}; SM clocks after the producer's first stamp (L2 flushed before the launch)")
print("pair " + " ".join(n.rjust(11) for n in names))
for i in range(min(n_my, 28)):
    print(f"{i:4d} " + " ".join(f"{int(x - t0) if x else 0:11d}" for x in tt[i][:len(names)]) + "  listed %d ovf %d long %d" % tuple(int(x) for x in tt[i][12:15]))
d = tt
k = min(4, n_my - 1)
print("mean cycles: load (full_ok - empty_ok) %.0f | join %.0f | compare %.0f | post (post_done - compared_ok) %.0f" %
      ((d[k:, 1] - d[k:, 0]).mean(), (d[:, 2] - d[:, 1]).mean(), (d[:, 4] - d[:, 3]).mean(), (d[:, 7] - d[:, 6]).mean()))
if n_my > 8:
    print("steady-state period per pair: %.0f cycles" % ((d[-1, 7] - d[4, 7]) / (n_my - 5)))
print("last post_done: %d cycles after the first stamp" % int(d[:, 7].max() - t0))
