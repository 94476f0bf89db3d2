"""single-frame-pair latencies (C1-C3) through the host-pointer C-ABI vs the reference CPU code (oracle/_ref, -O3)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200._abi import HostVoc
from oracle.pyoracle import Oracle, Reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = matcher.Context(0)
ref = Reference(fast=True) if Reference.available(fast=True) else None
orc = Oracle()

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e6

rows = []
# C1
c = synth.make_init_case(11)
f1, f2 = ctx.upload_frame(c.f1), ctx.upload_frame(c.f2)
m = matcher.ORBmatcher(c.nnratio, True, ctx)
g = timeit(lambda: m.SearchForInitialization(f1, f2, c.prev_matched, c.window_size))
gu = timeit(lambda: m.SearchForInitialization(ctx.upload_frame(c.f1), ctx.upload_frame(c.f2), c.prev_matched, c.window_size))
r = timeit(lambda: ref.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, c.nnratio, 1)) if ref else None
rows.append(("C1 SearchForInitialization 2x1000", g, gu, r, ctx.last_comparisons))
# C2
for th in (1.0, 3.0):
    pc = synth.make_projection_case(21, th=th)
    fr = ctx.upload_frame(pc.frame)
    m = matcher.ORBmatcher(pc.nnratio, True, ctx)
    g = timeit(lambda: m.SearchByProjection(fr, pc.mps, th, False, 50.0, pc.kp_prior_obs, pc.kp_mp))
    gu = timeit(lambda: m.SearchByProjection(ctx.upload_frame(pc.frame), pc.mps, th, False, 50.0, pc.kp_prior_obs, pc.kp_mp))
    r = timeit(lambda: ref.search_by_projection_local(pc.frame, pc.mps, th, 0, 50.0, pc.nnratio, pc.kp_prior_obs, pc.kp_mp)) if ref else None
    rows.append((f"C2 SearchByProjection 2000 kp x 5000 MP th={th}", g, gu, r, ctx.last_comparisons))
# C3
voc = HostVoc.load(os.path.join(ROOT, "tests", "golden", "voc_k10_L4.npz"))
dv = ctx.upload_vocabulary(voc)
hv = ref.voc_from_flat(voc) if ref else None
bc = synth.make_bow_case(31, voc, 2000)
for levelsup in (2, 4):
    dk, df = ctx.upload_frame(bc.kf), ctx.upload_frame(bc.f)
    g = timeit(lambda: dk.transform(dv, levelsup, True))
    r = timeit(lambda: hv.transform(bc.kf.desc, levelsup), n=10, warm=2) if ref else None
    rows.append((f"C3 transform 2000 features levelsup={levelsup}", g, None, r, ctx.last_comparisons))
    df.transform(dv, levelsup, True)
    m = matcher.ORBmatcher(0.7, True, ctx)
    g = timeit(lambda: m.SearchByBoW(dk, df, bc.kf_mp_valid))
    cmpc = ctx.last_comparisons
    w, nid, wt = orc.voc_transform(voc, bc.kf.desc, levelsup); kf = bc.kf.with_featvec(*orc.featvec(nid, wt))
    w, nid, wt = orc.voc_transform(voc, bc.f.desc, levelsup); f = bc.f.with_featvec(*orc.featvec(nid, wt))
    r = timeit(lambda: ref.search_by_bow_kf_f(kf, f, bc.kf_mp_valid, 0.7, 1), n=10, warm=2) if ref else None
    rows.append((f"C3 SearchByBoW KF-F 2000x2000 levelsup={levelsup}", g, None, r, cmpc))
print(f"{'config':58s} {'gpu us':>9s} {'gpu+upload us':>14s} {'ref cpu us':>11s} {'comparisons':>12s}")
for name, g, gu, r, c_ in rows:
    print(f"{name:58s} {g:9.1f} {gu if gu is None else round(gu,1)!s:>14s} {r if r is None else round(r,1)!s:>11s} {c_:12d}")
