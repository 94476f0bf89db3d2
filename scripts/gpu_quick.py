"""quick device timings (development aid): knn2 engines and the C4 batch, CUDA-event timed"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from orb_slam3_comments_ghr_b200 import matcher, synth

ctx = matcher.Context(0, stream=torch.cuda.current_stream().cuda_stream)
m = matcher.ORBmatcher(0.8, True, ctx)
nq, nd = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
db = torch.randint(0, 256, (nd, 32), dtype=torch.uint8, device="cuda", generator=g)
bi = torch.empty(nq, dtype=torch.int32, device="cuda"); bd = torch.empty_like(bi); sd = torch.empty_like(bi); mt = torch.empty_like(bi)
d = ctx.database_from_device(db.data_ptr(), nd, keepalive=db)
engines = [int(e) for e in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "2"])]
res = {}
for eng in engines:
    ctx.set_knn_engine(eng)
    for it in range(2):
        m.SearchByNN_dev(d, nq, q.data_ptr(), bi.data_ptr(), bd.data_ptr(), sd.data_ptr(), mt.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(3):
        m.SearchByNN_dev(d, nq, q.data_ptr(), bi.data_ptr(), bd.data_ptr(), sd.data_ptr(), mt.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[eng] = (bi.clone(), bd.clone(), sd.clone())
    print(f"knn2 engine {eng}: {nq}x{nd} {ms:.3f} ms  {nq*nd/ms/1e9:.3f} T cmp/s", flush=True)
if len(res) > 1:
    ks = list(res)
    for k in ks[1:]:
        print("engines agree:", k, all(torch.equal(a, b) for a, b in zip(res[ks[0]], res[k])))

# C4
P, nf = 1024, 2000
tc = synth.fill_geometry(synth.make_triangulation_case(9, n_pairs=P, n_feat=nf))
ks = ctx.upload_kfset(tc.kfs)
mm = matcher.ORBmatcher(0.6, False, ctx)
kf1 = torch.from_numpy(tc.kf1).cuda(); kf2 = torch.from_numpy(tc.kf2).cuda(); ep = torch.from_numpy(tc.ep).cuda(); f12 = torch.from_numpy(tc.f12).cuda()
out = torch.empty((P, nf), dtype=torch.int32, device="cuda"); nm = torch.empty(P, dtype=torch.int32, device="cuda")
for it in range(3):
    mm.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nm.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    mm.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nm.data_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"C4 {P} pairs x {nf}: {ms:.3f} ms  {P/ms*1e3:.0f} pairs/s  nmatches mean {nm.float().mean().item():.1f}")
