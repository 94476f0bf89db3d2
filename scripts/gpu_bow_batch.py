"""relocalisation loop: one frame against K candidate key frames, batched call vs K single calls (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200._abi import HostVoc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = matcher.Context(0)
voc = HostVoc.load(os.path.join(ROOT, "tests", "golden", "voc_k10_L4.npz"))
dv = ctx.upload_vocabulary(voc)
def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e6
for levelsup in (2, 3):
    bc = synth.make_bow_case(31, voc, 2000)
    df = ctx.upload_frame(bc.f); df.transform(dv, levelsup, True)
    m = matcher.ORBmatcher(0.7, True, ctx)
    for K in (1, 8, 32):
        dks, valids = [], []
        for k in range(K):
            ck = synth.make_bow_case(3100 + k, voc, 2000)
            d = ctx.upload_frame(ck.kf); d.transform(dv, levelsup, True)
            dks.append(d); valids.append(ck.kf_mp_valid)
        tb = timeit(lambda: m.SearchByBoWBatch(dks, df, valids))
        ts = timeit(lambda: [m.SearchByBoW(d, df, v) for d, v in zip(dks, valids)], n=10, warm=2)
        print(f"levelsup {levelsup} K={K:2d}: batched call {tb:8.1f} us = {tb / K:6.1f} us per candidate | {K} single calls {ts:8.1f} us = {ts / K:6.1f} us per candidate")
