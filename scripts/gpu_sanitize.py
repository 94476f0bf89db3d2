"""development aid: every kernel once at small sizes, meant to run under `compute-sanitizer --tool memcheck` (or racecheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200._abi import HostVoc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = matcher.Context(0)
voc = HostVoc.load(os.path.join(ROOT, "tests", "golden", "voc_k10_L4.npz"))
dv = ctx.upload_vocabulary(voc)
c = synth.make_init_case(11, n=400)
f1, f2 = ctx.upload_frame(c.f1), ctx.upload_frame(c.f2)
print("init", matcher.ORBmatcher(c.nnratio, True, ctx).SearchForInitialization(f1, f2, c.prev_matched, c.window_size)[0])
pc = synth.make_projection_case(21, n_kp=500, n_mp=900, th=3.0)
print("proj", matcher.ORBmatcher(pc.nnratio, True, ctx).SearchByProjection(ctx.upload_frame(pc.frame), pc.mps, 3.0, False, 50.0, pc.kp_prior_obs, pc.kp_mp)[0])
bc = synth.make_bow_case(31, voc, 500)
dk, df = ctx.upload_frame(bc.kf), ctx.upload_frame(bc.f)
dk.transform(dv, 2, True); df.transform(dv, 2, True)
print("bow", matcher.ORBmatcher(0.7, True, ctx).SearchByBoW(dk, df, bc.kf_mp_valid)[0], matcher.ORBmatcher(0.9, True, ctx).SearchByBoW(dk, df, bc.kf_mp_valid, bc.f_mp_valid)[0])
fr, pts, kl = synth.make_projected_case(71, n_kp=500, n_pts=700, th=7.0, stereo=True, level_mode="fwd")
inv = (1.0 / fr.level_sigma2).astype(np.float32)
print("projected", matcher.ORBmatcher(0.9, True, ctx).SearchProjected(ctx.upload_frame(fr), pts, 100.0, True, kl, stereo_gate=True)[0],
      matcher.ORBmatcher(0.9, False, ctx).SearchProjected(ctx.upload_frame(fr), pts, 50.0, False, None, chi2_gate=True, inv_level_sigma2=inv)[0])
for eng in (1, 2):
    tc = synth.fill_geometry(synth.make_triangulation_case(41, n_pairs=300, n_feat=600))
    ks = ctx.upload_kfset(tc.kfs)
    ctx.set_triangulation_engine(eng)
    for ori in (False, True):
        nm, m = matcher.ORBmatcher(0.6, ori, ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    offs, pairs = matcher.ORBmatcher(0.6, False, ctx).SearchForTriangulationPairs(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    print("tri engine", eng, int(nm.sum()), pairs.shape)
ctx.set_triangulation_engine(0)
# dense ties: survivors everywhere, > 32 candidates per node, overflow lists
tc = synth.fill_geometry(synth.make_triangulation_case(901, n_pairs=3, n_feat=900, n_nodes=8))
rng = np.random.default_rng(1)
base = synth.random_descriptors(rng, 2)
pick = rng.integers(0, 2, size=tc.kfs.desc.shape[:2])
tc.kfs.desc[:] = base[pick] ^ synth.flip_mask(rng, pick.size, np.full(pick.size, 6)).reshape(*pick.shape, 32)
print("tri dense", int(matcher.ORBmatcher(0.6, True, ctx).SearchForTriangulation(ctx.upload_kfset(tc.kfs), tc.kf1, tc.kf2, tc.ep, tc.f12)[0].sum()))
kc = synth.make_knn_case(51, 300, 5000)
for eng in (1, 2, 3, 4):
    ctx.set_knn_engine(eng)
    r = matcher.ORBmatcher(kc.nnratio, True, ctx).SearchByNN(ctx.upload_database(kc.db), kc.q, kc.th_low)
    print("knn2 engine", eng, int((r[3] >= 0).sum()))
ctx.set_knn_engine(0)
db, qw, qv = synth.make_bowdb_case(121, n_kf=300)
print("bowdb", ctx.upload_bow_database(db).score(qw, qv)[0].max())
offs, desc = synth.make_distinctive_case(141, n_mp=60, max_obs=20)
print("distinctive", ctx.compute_distinctive_descriptors(offs, desc)[0][:5])
l, r, nr, mb, mbf = synth.make_stereo_case(161, n=400)
print("stereo", int((ctx.stereo_coarse_match(l, r, nr, mb, mbf)[0] >= 0).sum()))
print("done")
