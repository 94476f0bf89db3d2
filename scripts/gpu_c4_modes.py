"""C4 step time by output mode / launch mode on one GPU (development aid): dense rows vs compact pairs, direct launch vs graph"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200.sharding import TriangulationGather

P = int(os.environ.get("PAIRS", 4096))
dev = torch.device("cuda", 0)
case = synth.fill_geometry(synth.make_triangulation_case(20261018, n_pairs=P, n_feat=2000))
ctx = matcher.Context(0, stream=torch.cuda.current_stream().cuda_stream)
ks = ctx.upload_kfset(case.kfs)
m = matcher.ORBmatcher(0.6, False, ctx)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
kf1, kf2, ep, f12 = t(case.kf1), t(case.kf2), t(case.ep), t(case.f12)
out = torch.empty((P, 2000), dtype=torch.int32, device=dev)
nmt = torch.empty(P, dtype=torch.int32, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(step, n=300, do_flush=True):
    for _ in range(5): step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        if do_flush: flush.fill_(1)
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], sum(ts) / len(ts)

dense = lambda: m.SearchForTriangulation_dev(ks, P, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), out.data_ptr(), nmt.data_ptr())
print("dense rows, direct launch      median %.4f mean %.4f ms" % timeit(dense))
print("dense rows, no flush           median %.4f mean %.4f ms" % timeit(dense, do_flush=False))
for graph in (False, True):
    tg = TriangulationGather(matcher, case.kfs, P, 2000, 0, 1, dev, 0.6, False, use_graph=graph)
    tg.set_inputs(kf1, kf2, ep, f12)
    print("compact pairs, graph=%-5s      median %.4f mean %.4f ms" % ((graph,) + timeit(tg.step)))
    print("compact pairs, graph=%-5s nofl median %.4f mean %.4f ms" % ((graph,) + timeit(tg.step, do_flush=False)))
    del tg
