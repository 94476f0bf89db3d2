"""development aid: every single-frame entry point once (after one warm-up pass), for an ncu capture of each kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200._abi import HostVoc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = matcher.Context(0)
voc = HostVoc.load(os.path.join(ROOT, "tests", "golden", "voc_k10_L4.npz"))
dv = ctx.upload_vocabulary(voc)
c = synth.make_init_case(11)
pc = synth.make_projection_case(21, th=3.0)
bc = synth.make_bow_case(31, voc, 2000)
fr, pts, kl = synth.make_projected_case(71, th=7.0)
db, qw, qv = synth.make_bowdb_case(121, n_kf=3000)
offs, desc = synth.make_distinctive_case(141, n_mp=3000)
bdb = ctx.upload_bow_database(db)
for it in range(2):
    f1, f2 = ctx.upload_frame(c.f1), ctx.upload_frame(c.f2)
    matcher.ORBmatcher(c.nnratio, True, ctx).SearchForInitialization(f1, f2, c.prev_matched, c.window_size)
    f = ctx.upload_frame(pc.frame)
    matcher.ORBmatcher(pc.nnratio, True, ctx).SearchByProjection(f, pc.mps, 3.0, False, 50.0, pc.kp_prior_obs, pc.kp_mp)
    dk, df = ctx.upload_frame(bc.kf), ctx.upload_frame(bc.f)
    dk.transform(dv, 2, True); df.transform(dv, 2, True)
    matcher.ORBmatcher(0.7, True, ctx).SearchByBoW(dk, df, bc.kf_mp_valid)
    matcher.ORBmatcher(0.9, True, ctx).SearchByBoW(dk, df, bc.kf_mp_valid, bc.f_mp_valid)
    dk.transform(dv, 4, True); df.transform(dv, 4, True)  # one root bucket: the big-node fixed point
    matcher.ORBmatcher(0.7, True, ctx).SearchByBoW(dk, df, bc.kf_mp_valid)
    matcher.ORBmatcher(0.9, True, ctx).SearchProjected(ctx.upload_frame(fr), pts, 100.0, True, kl)
    bdb.score(qw, qv)
    ctx.compute_distinctive_descriptors(offs, desc)
print("done")
