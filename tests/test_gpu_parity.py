"""GPU parity tests (run with -m gpu on a B200): every search goes through the C-ABI library
(liborbmatch_b200.so) and must be BIT-EXACT against the CPU oracle on the same seeded inputs and
against the golden vectors generated from the reference itself."""
import os

import numpy as np
import pytest

from helpers import GOLD, attach_featvec, digest, golden_cases, golden_outputs, golden_voc
from orb_slam3_comments_ghr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from orb_slam3_comments_ghr_b200 import matcher
    return matcher.Context(0)


@pytest.fixture(scope="module")
def M():
    from orb_slam3_comments_ghr_b200 import matcher
    return matcher


def test_descriptor_distance(ctx, oracle):
    z = np.load(f"{GOLD}/descriptor_distance.npz")
    assert np.array_equal(ctx.descriptor_distance(z["a"], z["b"]), z["dist"])
    rng = np.random.default_rng(0)
    a, b = synth.random_descriptors(rng, 100000), synth.random_descriptors(rng, 100000)
    assert np.array_equal(ctx.descriptor_distance(a, b), np.unpackbits(a ^ b, axis=1).sum(axis=1))


def test_three_maxima(ctx):
    z = np.load(f"{GOLD}/three_maxima.npz")
    for h, ind in zip(z["histo"], z["ind"]):
        assert np.array_equal(ctx.compute_three_maxima(h), ind)


@pytest.mark.parametrize("seed", [101, 102])
def test_grid_and_area(ctx, oracle, seed):
    rng = np.random.default_rng(seed)
    f = synth.make_frame(rng, 2000)
    f.kp_xy[:40, 0] = np.arange(40, dtype=np.float32) * 5.0 + 5.0
    f.kp_xy[40:44] = np.array([[639.75, 479.75], [0, 0], [636.0, 476.0], [635.0, 475.0]], dtype=np.float32)
    d = ctx.upload_frame(f)
    cs, ci = d.grid()
    ocs, oci = oracle.grid(f)
    assert np.array_equal(cs, ocs) and np.array_equal(ci[:cs[-1]], oci[:ocs[-1]])
    nq = 500
    x, y = rng.uniform(-50, 700, nq).astype(np.float32), rng.uniform(-50, 530, nq).astype(np.float32)
    r = rng.choice(np.array([1.0, 2.5, 7.5, 40.0, 100.0, 1000.0], dtype=np.float32), nq)
    lv = rng.integers(0, 8, nq)
    kind = rng.integers(0, 5, nq)
    mn = np.select([kind == 0, kind == 1, kind == 2, kind == 3, kind == 4], [-1, lv - 1, lv, 0, lv]).astype(np.int32)
    mx = np.select([kind == 0, kind == 1, kind == 2, kind == 3, kind == 4], [-1, lv, lv, lv, -1]).astype(np.int32)
    off, idx = d.features_in_area(x, y, r, mn, mx)
    for q in range(nq):
        exp = oracle.features_in_area(f, x[q], y[q], r[q], mn[q], mx[q], grid=(ocs, oci))
        assert np.array_equal(idx[off[q]:off[q + 1]], exp), q


@pytest.mark.parametrize("name", ["init_s11", "init_s12_n5000"])
def test_init_golden(ctx, M, name):
    g = golden_outputs()
    c = golden_cases()[name]()
    m = M.ORBmatcher(c.nnratio, bool(c.check_ori), ctx)
    n, m12, prev = m.SearchForInitialization(ctx.upload_frame(c.f1), ctx.upload_frame(c.f2), c.prev_matched, c.window_size)
    assert n == int(g[name + "/nmatches"]) and np.array_equal(m12, g[name + "/matches12"]) and np.array_equal(prev, g[name + "/prev"])


@pytest.mark.parametrize("seed,n,ratio,ori", [(201, 1000, 0.9, 1), (202, 1000, 0.6, 0), (203, 300, 0.9, 1), (204, 2500, 0.95, 1), (205, 1, 0.9, 1),
                                              (206, 1800, 0.9, 1)])  # 1000: lists + inverse index in shared memory; 1800: inverse index in global memory; 2500: lists in global memory too
def test_init_oracle(ctx, M, oracle, seed, n, ratio, ori):
    c = synth.make_init_case(seed, n=n)
    m = M.ORBmatcher(ratio, bool(ori), ctx)
    got = m.SearchForInitialization(ctx.upload_frame(c.f1), ctx.upload_frame(c.f2), c.prev_matched, c.window_size)
    oracle.reset_comparisons()
    exp = oracle.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, ratio, ori)
    assert got[0] == exp[0] and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("seed,n,n_proto,ratio", [(211, 800, 6, 1.0), (212, 1000, 40, 0.95), (213, 400, 2, 1.0)])
def test_init_contention(ctx, M, oracle, seed, n, n_proto, ratio):
    """the accept fixed point under contention: level-0 descriptors of both frames are noisy copies of a few prototypes, so many
    key points accept the same partners with shrinking distances (vMatchedDistance chains, :790/:822) and displace each other
    (:813-817); ratio 1.0 keeps ties alive"""
    c = synth.make_init_case(seed, n=n)
    rng = np.random.default_rng(seed)
    proto = synth.random_descriptors(rng, n_proto)
    for fr in (c.f1, c.f2):
        l0 = np.flatnonzero(fr.octave == 0)
        fr.desc[l0] = proto[rng.integers(0, n_proto, l0.size)] ^ synth.flip_mask(rng, l0.size, rng.choice(np.array([4, 5, 6, 9]), size=l0.size))
    for ori in (0, 1):
        m = M.ORBmatcher(ratio, bool(ori), ctx)
        got = m.SearchForInitialization(ctx.upload_frame(c.f1), ctx.upload_frame(c.f2), c.prev_matched.copy(), c.window_size)
        exp = oracle.search_for_initialization(c.f1, c.f2, c.prev_matched.copy(), c.window_size, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    assert exp[0] > 5


@pytest.mark.parametrize("name", ["proj_s21_th1", "proj_s22_th3", "proj_s23_far"])
def test_projection_golden(ctx, M, name):
    g = golden_outputs()
    c = golden_cases()[name]()
    m = M.ORBmatcher(c.nnratio, True, ctx)
    n, k = m.SearchByProjection(ctx.upload_frame(c.frame), c.mps, c.th, bool(c.far_points), c.th_far, c.kp_prior_obs, c.kp_mp)
    assert n == int(g[name + "/nmatches"]) and np.array_equal(k, g[name + "/kp_mp"])


@pytest.mark.parametrize("seed,th,far", [(301, 1.0, 0), (302, 3.0, 0), (303, 15.0, 1), (304, 2.0, 1)])
def test_projection_oracle(ctx, M, oracle, seed, th, far):
    c = synth.make_projection_case(seed, th=th, far_points=far)
    m = M.ORBmatcher(c.nnratio, True, ctx)
    got = m.SearchByProjection(ctx.upload_frame(c.frame), c.mps, th, bool(far), 40.0, c.kp_prior_obs, c.kp_mp)
    oracle.reset_comparisons()
    exp = oracle.search_by_projection_local(c.frame, c.mps, th, far, 40.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("levelsup", [2, 4])
def test_transform_and_bow_golden(ctx, M, levelsup):
    g = golden_outputs()
    voc = golden_voc()
    dv = ctx.upload_vocabulary(voc)
    bc = synth.make_bow_case(31, voc, 2000)
    dev = {}
    for side, fr in (("kf", bc.kf), ("f", bc.f)):
        d = ctx.upload_frame(fr)
        w, nid, wt = d.transform(dv, levelsup, True)
        p = f"bow_s31/l{levelsup}/{side}/"
        assert np.array_equal(w, g[p + "word_id"]) and np.array_equal(nid, g[p + "node_id"]) and np.array_equal(wt, g[p + "weight"])
        bw, bv = d.bowvector()
        assert np.array_equal(bw, g[p + "bow_words"]) and np.array_equal(bv, g[p + "bow_values"])
        fn, fo, ff = d.featvec()
        assert np.array_equal(fn, g[p + "fv_node_ids"]) and np.array_equal(fo, g[p + "fv_offsets"]) and np.array_equal(ff, g[p + "fv_features"])
        dev[side] = d
    n, m = M.ORBmatcher(0.7, True, ctx).SearchByBoW(dev["kf"], dev["f"], bc.kf_mp_valid)
    assert n == int(g[f"bow_s31/l{levelsup}/kf_f/nmatches"]) and np.array_equal(m, g[f"bow_s31/l{levelsup}/kf_f/match"])
    n, m = M.ORBmatcher(0.9, True, ctx).SearchByBoW(dev["kf"], dev["f"], bc.kf_mp_valid, bc.f_mp_valid)
    assert n == int(g[f"bow_s31/l{levelsup}/kf_kf/nmatches"]) and np.array_equal(m, g[f"bow_s31/l{levelsup}/kf_kf/match"])


@pytest.mark.parametrize("levelsup", [0, 1, 2, 3, 4, 6])
def test_transform_random_vocabulary(ctx, oracle, levelsup):
    voc = synth.random_vocabulary(7, k=6, L=4, ragged=True)
    dv = ctx.upload_vocabulary(voc)
    rng = np.random.default_rng(levelsup)
    desc = synth.descriptors_near_words(rng, voc, 700)
    fr = synth.make_frame(rng, 700)
    fr.desc[:] = desc
    d = ctx.upload_frame(fr)
    w, nid, wt = d.transform(dv, levelsup, True)
    oracle.reset_comparisons()
    ow, onid, owt = oracle.voc_transform(voc, desc, levelsup)
    assert ctx.last_comparisons == oracle.comparisons()
    assert np.array_equal(w, ow) and np.array_equal(nid, onid) and np.array_equal(wt, owt)
    bw, bv = d.bowvector()
    obw, obv = oracle.bowvector(ow, owt)
    assert np.array_equal(bw, obw) and np.array_equal(bv, obv)
    fn, fo, ff = d.featvec()
    ofn, ofo, off_ = oracle.featvec(onid, owt)
    assert np.array_equal(fn, ofn) and np.array_equal(fo, ofo) and np.array_equal(ff, off_)


@pytest.mark.parametrize("seed,levelsup,ratio", [(401, 2, 0.7), (402, 3, 0.75), (403, 4, 0.9)])
def test_bow_host_featvec(ctx, M, oracle, seed, levelsup, ratio):
    """FeatureVectors supplied by the host (flattened std::map), as the drop-in adapter does"""
    voc = golden_voc()
    bc = synth.make_bow_case(seed, voc, 1200)
    kf = attach_featvec(oracle, voc, bc.kf, levelsup)
    f = attach_featvec(oracle, voc, bc.f, levelsup)
    dkf, df = ctx.upload_frame(kf), ctx.upload_frame(f)
    for ori in (0, 1):
        m = M.ORBmatcher(ratio, bool(ori), ctx)
        got = m.SearchByBoW(dkf, df, bc.kf_mp_valid)
        oracle.reset_comparisons()
        exp = oracle.search_by_bow_kf_f(kf, f, bc.kf_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
        assert ctx.last_comparisons == oracle.comparisons()
        got = m.SearchByBoW(dkf, df, bc.kf_mp_valid, bc.f_mp_valid)
        exp = oracle.search_by_bow_kf_kf(kf, f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])


@pytest.mark.parametrize("seed,layout,ratio,n1,n2", [(601, "root", 0.95, 1200, 1500), (602, "mixed", 0.9, 1200, 1500), (603, "root", 0.7, 1200, 1500),
                                                   (604, "root", 0.95, 40, 300), (605, "mixed", 1.0, 3000, 4000), (606, "root", 0.95, 31, 400)])
def test_bow_conflicts(ctx, M, oracle, seed, layout, ratio, n1, n2):
    """big node pairs go through the lock-time fixed point (bow_big_*_kernel): heavy contention for near-duplicate partners,
    candidate lists that run out (full rescans), a crowded group with more contenders than members; (606) just under the
    size that selects the path"""
    bc = synth.make_bow_conflict_case(seed, n1=n1, n2=n2, layout=layout)
    dkf, df = ctx.upload_frame(bc.kf), ctx.upload_frame(bc.f)
    for ori in (0, 1):
        m = M.ORBmatcher(ratio, bool(ori), ctx)
        got = m.SearchByBoW(dkf, df, bc.kf_mp_valid)
        oracle.reset_comparisons()
        exp = oracle.search_by_bow_kf_f(bc.kf, bc.f, bc.kf_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
        assert ctx.last_comparisons == oracle.comparisons()
        got = m.SearchByBoW(dkf, df, bc.kf_mp_valid, bc.f_mp_valid)
        oracle.reset_comparisons()
        exp = oracle.search_by_bow_kf_kf(bc.kf, bc.f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
        assert ctx.last_comparisons == oracle.comparisons()
    # all-invalid keyframe: nothing to match, nothing compared
    got = M.ORBmatcher(ratio, True, ctx).SearchByBoW(dkf, df, np.zeros_like(bc.kf_mp_valid))
    assert got[0] == 0 and (got[1] == -1).all() and ctx.last_comparisons == 0


# search core of the self-projecting overloads (row a6)
PROJECTED_MODES = [  # name, level_mode, ordered, stereo, stereo_gate, chi2_gate, max_dist, th, ori
    ("cur_last", "pm1", 1, False, 0, 0, 100.0, 7.0, 1),
    ("cur_last_fwd_stereo", "fwd", 1, True, 1, 0, 100.0, 15.0, 1),
    ("cur_last_bwd", "bwd", 1, False, 0, 0, 100.0, 15.0, 0),
    ("reloc", "pm1", 1, False, 0, 0, 64.0, 10.0, 1),
    ("sim3", "pred", 1, False, 0, 0, 50.0 * 0.9, 8.0, 0),
    ("fuse_mono", "pred", 0, False, 0, 1, 50.0, 3.0, 0),
    ("fuse_stereo", "pred", 0, True, 0, 1, 50.0, 3.0, 0),
    ("by_sim3", "pred", 0, False, 0, 0, 100.0, 7.5, 0),
]


@pytest.mark.parametrize("mode", PROJECTED_MODES, ids=[m[0] for m in PROJECTED_MODES])
@pytest.mark.parametrize("seed", [71, 72])
def test_search_projected_oracle(ctx, M, oracle, mode, seed):
    name, level_mode, ordered, stereo, sgate, chi2, max_dist, th, ori = mode
    frame, pts, kp_locked = synth.make_projected_case(seed, th=th, stereo=stereo, level_mode=level_mode,
                                                      lock_frac=1.0 if name in ("reloc", "sim3") else 0.85)
    inv = (1.0 / frame.level_sigma2).astype(np.float32)
    df = ctx.upload_frame(frame)
    got = M.ORBmatcher(0.9, bool(ori), ctx).SearchProjected(df, pts, max_dist, bool(ordered), kp_locked if ordered else None,
                                                           stereo_gate=bool(sgate), chi2_gate=bool(chi2), inv_level_sigma2=inv)
    oracle.reset_comparisons()
    exp = oracle.search_projected(frame, pts, max_dist, ordered, kp_locked if ordered else None, sgate, chi2, ori, inv)
    assert exp[0] > 50, "the case must produce matches"
    assert got[0] == exp[0]
    for a, b, what in zip(got[1:], exp[1:], ("best_idx", "best_dist", "kp_owner")):
        assert np.array_equal(a, b), what
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("name", ["curlast_pm1", "curlast_fwd_stereo", "curlast_bwd", "reloc", "sim3", "fuse_stereo", "fuse_sim3"])
def test_search_projected_golden(ctx, M, name):
    """GPU against the committed outputs of the reference's own self-projecting overloads (tests/golden/projected_outputs.npz)"""
    from helpers import projected_cases, projected_golden
    g = projected_golden()
    mk, kind, _, kw = projected_cases()[name]
    frame, pts, kl = mk()
    kw = dict(kw)
    inv = (np.float32(1.0) / frame.level_sigma2).astype(np.float32) if kw.get("chi2_gate") else None
    ordered = kw["ordered"]
    n, bi, bd, own = M.ORBmatcher(0.9, bool(kw.get("check_ori", 0)), ctx).SearchProjected(
        ctx.upload_frame(frame), pts, kw["max_dist"], bool(ordered), kl if ordered else None, stereo_gate=bool(kw.get("stereo_gate", 0)),
        chi2_gate=bool(kw.get("chi2_gate", 0)), inv_level_sigma2=inv)
    assert n == int(g[name + "/nmatches"])
    if name + "/kp_owner" in g:
        assert np.array_equal(own, g[name + "/kp_owner"])
    else:
        assert np.array_equal(bi, g[name + "/best_idx"])


def test_search_projected_edge_cases(ctx, M, oracle):
    frame, pts, kp_locked = synth.make_projected_case(5, n_kp=300, n_pts=0)
    df = ctx.upload_frame(frame)
    got = M.ORBmatcher(0.9, True, ctx).SearchProjected(df, pts, 100.0, True, kp_locked)
    assert got[0] == 0 and got[1].size == 0 and np.all(got[3] == -1)
    # every point on the same keypoint: a chain of takes and skips
    frame, pts, kp_locked = synth.make_projected_case(6, n_kp=64, n_pts=400, th=40.0, planted_frac=1.0, lock_frac=0.5)
    df = ctx.upload_frame(frame)
    got = M.ORBmatcher(0.9, True, ctx).SearchProjected(df, pts, 100.0, True, kp_locked)
    exp = oracle.search_projected(frame, pts, 100.0, 1, kp_locked, 0, 0, 1, None)
    assert got[0] == exp[0] and all(np.array_equal(a, b) for a, b in zip(got[1:], exp[1:]))


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("check_ori", [0, 1])
def test_triangulation_golden(ctx, M, check_ori, engine):
    g = golden_outputs()
    tc = synth.fill_geometry(synth.make_triangulation_case(41, n_pairs=8, n_feat=2000))
    ks = ctx.upload_kfset(tc.kfs)
    ctx.set_triangulation_engine(engine)
    nm, m = M.ORBmatcher(0.6, bool(check_ori), ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, g["tri_s41/ep"], g["tri_s41/f12"])
    ctx.set_triangulation_engine(0)
    assert np.array_equal(nm, g[f"tri_s41/ori{check_ori}/nmatches"])
    assert np.array_equal(m, g[f"tri_s41/ori{check_ori}/matches"].astype(np.int32))


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("seed,n_pairs,n_feat,coarse,ori", [(501, 6, 1500, 0, 0), (502, 6, 1500, 0, 1), (503, 6, 1500, 1, 1), (504, 96, 2000, 0, 0), (505, 3, 64, 0, 1),
                                                            (506, 700, 500, 0, 0), (507, 450, 300, 0, 1), (508, 1, 2000, 0, 0),
                                                            (509, 40, 777, 0, 1), (510, 9, 1001, 0, 0), (511, 300, 2, 1, 0)])
def test_triangulation_oracle(ctx, M, oracle, seed, n_pairs, n_feat, coarse, ori, engine):
    tc = synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=n_pairs, n_feat=n_feat))
    ks = ctx.upload_kfset(tc.kfs)
    ctx.set_triangulation_engine(engine)
    nm, m = M.ORBmatcher(0.6, bool(ori), ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, False, bool(coarse))
    ctx.set_triangulation_engine(0)
    oracle.reset_comparisons()
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, coarse, ori, n_threads=os.cpu_count() or 1)
    assert np.array_equal(nm, enm) and np.array_equal(m, em)
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("n_distinct,coarse,ori", [(1, 1, 0), (3, 1, 1), (5, 0, 0), (2, 0, 1)])
def test_triangulation_dense_ties(ctx, M, oracle, engine, n_distinct, coarse, ori):
    """adversarial: only a handful of distinct descriptors, so almost every candidate of a node survives
    dist <= TH_LOW with equal distances -- exercises the last-wins tie-break (:1180), the survivor list
    beyond one entry per thread and its overflow path."""
    tc = synth.fill_geometry(synth.make_triangulation_case(900 + n_distinct, n_pairs=5, n_feat=1200, n_nodes=12))
    rng = np.random.default_rng(n_distinct)
    base = synth.random_descriptors(rng, n_distinct)
    pick = rng.integers(0, n_distinct, size=tc.kfs.desc.shape[:2])
    tc.kfs.desc[:] = base[pick] ^ synth.flip_mask(rng, pick.size, np.full(pick.size, 6)).reshape(*pick.shape, 32)
    ks = ctx.upload_kfset(tc.kfs)
    ctx.set_triangulation_engine(engine)
    nm, m = M.ORBmatcher(0.6, bool(ori), ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, False, bool(coarse))
    ctx.set_triangulation_engine(0)
    oracle.reset_comparisons()
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, coarse, ori, n_threads=os.cpu_count() or 1)
    assert int(enm.sum()) > 0
    assert np.array_equal(nm, enm) and np.array_equal(m, em)
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("engine", [1, 2, 3, 4])
def test_knn2_golden(ctx, M, engine):
    g = golden_outputs()
    kc = synth.make_knn_case(51, 512, 20000)
    ctx.set_knn_engine(engine)
    got = M.ORBmatcher(kc.nnratio, True, ctx).SearchByNN(ctx.upload_database(kc.db), kc.q, kc.th_low)
    ctx.set_knn_engine(0)
    for a, name in zip(got, ("best_idx", "best_dist", "second_dist", "match")):
        assert np.array_equal(a, g["knn_s51/" + name]), name


@pytest.mark.parametrize("engine", [1, 2, 3, 4])
@pytest.mark.parametrize("nq,nd", [(1, 1), (3, 2), (1000, 257), (4097, 70001), (300, 3), (128, 1024), (129, 1025), (1000, 4097)])
def test_knn2_oracle_ragged(ctx, M, oracle, engine, nq, nd):
    kc = synth.make_knn_case(nq * 7 + nd, nq, nd)
    ctx.set_knn_engine(engine)
    got = M.ORBmatcher(0.8, True, ctx).SearchByNN(ctx.upload_database(kc.db), kc.q, 50)
    ctx.set_knn_engine(0)
    exp = oracle.knn2_ratio(kc.q, kc.db, 50, 0.8, n_threads=os.cpu_count() or 1)
    for a, e, name in zip(got, exp, ("best_idx", "best_dist", "second_dist", "match")):
        assert np.array_equal(a, e), name
    assert ctx.last_comparisons == nq * nd


@pytest.mark.parametrize("engine", [3, 4])
def test_knn2_tc_ties_and_duplicates(ctx, M, oracle, engine):
    """tensor engine: exact duplicates and equal distances across stages / splits must resolve to the FIRST index"""
    rng = np.random.default_rng(5)
    nd, nq = 40000, 640
    db = synth.random_descriptors(rng, nd)
    q = synth.random_descriptors(rng, nq)
    for i in range(nq):  # plant the same near-copy of the query at several rows (same distance -> first index wins)
        rows = np.sort(rng.choice(nd, size=4, replace=False))
        near = q[i] ^ synth.flip_mask(rng, 1, 5)[0]
        db[rows] = near
    ctx.set_knn_engine(engine)
    got = M.ORBmatcher(0.8, True, ctx).SearchByNN(ctx.upload_database(db), q, 50)
    ctx.set_knn_engine(0)
    exp = oracle.knn2_ratio(q, db, 50, 0.8, n_threads=os.cpu_count() or 1)
    for a, e, name in zip(got, exp, ("best_idx", "best_dist", "second_dist", "match")):
        assert np.array_equal(a, e), name


@pytest.mark.parametrize("engine", [0, 1, 3])
@pytest.mark.parametrize("nd", [1500, 131072, 700001])
def test_knn2_update_and_search_overlapped(ctx, M, oracle, engine, nd):
    """orbgpu_knn2_ratio_update: new descriptors uploaded in chunks on the database's own stream, every chunk searched as it arrives --
    same results as update + search, as the oracle on the NEW descriptors, for sizes of less than one split, exactly one split and
    several chunks with a ragged tail; then a smaller database in the same object, then the plain call on the resident rows"""
    rng = np.random.default_rng(nd + engine)
    nq = 6656 if nd > 200000 else 700  # the large case has enough query tiles for whole 131 072-row splits: the chunked path proper
    old = synth.random_descriptors(rng, nd)
    new = synth.random_descriptors(rng, nd)
    q = new[rng.integers(0, nd, nq)] ^ synth.flip_mask(rng, nq, 6)
    q[::7] = synth.random_descriptors(rng, len(q[::7]))
    ctx.set_knn_engine(engine)
    m = M.ORBmatcher(0.8, True, ctx)
    d = ctx.upload_database(old)
    m.SearchByNN(d, q, 50)  # the expansion of the OLD rows is cached now
    got = m.SearchByNN(d, q, 50, database=new)
    exp = oracle.knn2_ratio(q, new, 50, 0.8, n_threads=os.cpu_count() or 1)
    for a, e, name in zip(got, exp, ("best_idx", "best_dist", "second_dist", "match")):
        assert np.array_equal(a, e), name
    again = m.SearchByNN(d, q, 50)  # resident rows = the new ones
    assert all(np.array_equal(a, b) for a, b in zip(got, again))
    n2 = nd - nd // 3  # fewer rows in the same object: the tail of the old expansion must not be seen
    got2 = m.SearchByNN(d, q, 50, database=old[:n2])
    exp2 = oracle.knn2_ratio(q, old[:n2], 50, 0.8, n_threads=os.cpu_count() or 1)
    for a, e, name in zip(got2, exp2, ("best_idx", "best_dist", "second_dist", "match")):
        assert np.array_equal(a, e), name
    d.update(new)  # the plain update after a chunked one
    got3 = m.SearchByNN(d, q, 50)
    assert all(np.array_equal(a, b) for a, b in zip(got, got3))
    ctx.set_knn_engine(0)


def test_knn2_properties_large(ctx, M):
    """size-independent properties at a size the oracle cannot finish: planted queries must find their
    source row (or an earlier exact duplicate), best <= second, idempotence."""
    rng = np.random.default_rng(77)
    nd, nq = 1 << 20, 1 << 14
    db = synth.random_descriptors(rng, nd)
    src = rng.integers(0, nd, nq)
    q = db[src] ^ synth.flip_mask(rng, nq, 5)
    m = M.ORBmatcher(0.8, True, ctx)
    d = ctx.upload_database(db)
    bi, bd, sd, mt = m.SearchByNN(d, q, 50)
    true_d = np.unpackbits(q ^ db[src], axis=1).sum(axis=1)
    assert (bd <= true_d).all() and (bd <= sd).all()
    same = bi == src
    assert same.mean() > 0.999
    got_d = np.unpackbits(q ^ db[bi], axis=1).sum(axis=1)
    assert np.array_equal(got_d, bd)
    again = m.SearchByNN(d, q, 50)
    assert all(np.array_equal(a, b) for a, b in zip((bi, bd, sd, mt), again))


def test_triangulation_full_size_properties(ctx, M, oracle):
    """BASELINE.json config C4 at full size (4096 key-frame pairs x 2000 features): the two engines agree on every match row,
    count and comparison counter; a 64-pair sample equals the oracle; every reported match satisfies the reference's
    acceptance conditions (both features without a map point, same vocabulary node, distance <= TH_LOW); idempotence."""
    P = 4096
    tc = synth.fill_geometry(synth.make_triangulation_case(20261018, n_pairs=P, n_feat=2000))
    ks = ctx.upload_kfset(tc.kfs)
    mm = M.ORBmatcher(0.6, False, ctx)
    res = {}
    for eng in (1, 2):
        ctx.set_triangulation_engine(eng)
        nm, m = mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
        res[eng] = (nm.copy(), m.copy(), ctx.last_comparisons)
    ctx.set_triangulation_engine(0)
    nm2, m2 = mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    assert np.array_equal(res[1][0], res[2][0]) and np.array_equal(res[1][1], res[2][1]) and res[1][2] == res[2][2]
    assert np.array_equal(nm2, res[2][0]) and np.array_equal(m2, res[2][1])
    nm, m, _ = res[2]
    assert np.array_equal(nm, (m >= 0).sum(axis=1)) and nm.sum() > 100 * P
    sel = np.arange(0, P, 64)
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1[sel], tc.kf2[sel], tc.ep[sel], tc.f12[sel], 0, 0, 0,
                                                    n_threads=os.cpu_count() or 1)
    assert np.array_equal(nm[sel], enm) and np.array_equal(m[sel], em)
    p, i1 = np.nonzero(m >= 0)
    i2 = m[p, i1]
    k1, k2 = tc.kf1[p], tc.kf2[p]
    assert not tc.kfs.has_mp[k1, i1].any() and not tc.kfs.has_mp[k2, i2].any()
    assert np.array_equal(tc.kfs.node_id[k1, i1], tc.kfs.node_id[k2, i2])
    d = np.unpackbits(tc.kfs.desc[k1, i1] ^ tc.kfs.desc[k2, i2], axis=1).sum(axis=1)
    assert d.max() <= 50


@pytest.mark.parametrize("seed,n_kf", [(121, 3000), (122, 257), (123, 1)])
def test_bow_score_l1(ctx, oracle, seed, n_kf):
    """8(f) rank 2: KeyFrameDatabase candidate scoring -- bit-exact doubles (the sum runs in word order per key frame)"""
    db, qw, qv = synth.make_bowdb_case(seed, n_kf=n_kf)
    ddb = ctx.upload_bow_database(db)
    gc, gs = ddb.score(qw, qv)
    ec, es = oracle.bow_score_l1(db, qw, qv)
    assert np.array_equal(gc, ec) and np.array_equal(gs.view(np.uint64), es.view(np.uint64))
    gc, gs = ddb.score(qw[:0], qv[:0])
    assert not gc.any() and not gs.any()


@pytest.mark.parametrize("seed,n_mp,max_obs", [(141, 3000, 40), (142, 50, 300), (143, 1, 1)])
def test_compute_distinctive_descriptors(ctx, oracle, seed, n_mp, max_obs):
    """8(f) rank 4: batched MapPoint::ComputeDistinctiveDescriptors"""
    offs, desc = synth.make_distinctive_case(seed, n_mp=n_mp, max_obs=max_obs)
    gi, gm = ctx.compute_distinctive_descriptors(offs, desc)
    oracle.reset_comparisons()
    ei, em = oracle.compute_distinctive_descriptors(offs, desc)
    assert np.array_equal(gi, ei) and np.array_equal(gm, em)
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("seed,n", [(161, 2000), (162, 333), (163, 1)])
def test_stereo_coarse_match(ctx, oracle, seed, n):
    """8(f) rank 3: coarse stage of Frame::ComputeStereoMatches"""
    left, right, n_rows, mb, mbf = synth.make_stereo_case(seed, n=n)
    gi, gd = ctx.stereo_coarse_match(left, right, n_rows, mb, mbf)
    oracle.reset_comparisons()
    ei, ed = oracle.stereo_coarse_match(left, right, n_rows, mb, mbf)
    assert np.array_equal(gi, ei) and np.array_equal(gd, ed)
    assert ctx.last_comparisons == oracle.comparisons()
    if n >= 333:
        assert (gi >= 0).sum() > n // 10


@pytest.mark.parametrize("seed,n_pairs,n_feat,ori", [(171, 37, 1500, 0), (172, 300, 600, 1), (173, 1, 64, 0)])
def test_triangulation_pairs_form(ctx, M, seed, n_pairs, n_feat, ori):
    """the vMatchedPairs form equals the dense rows: (idx1, idx2) ascending in idx1, per pair (ORBmatcher.cc:1317-1325)"""
    tc = synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=n_pairs, n_feat=n_feat))
    ks = ctx.upload_kfset(tc.kfs)
    mm = M.ORBmatcher(0.6, bool(ori), ctx)
    nm, m = mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    offs, pairs = mm.SearchForTriangulationPairs(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    assert np.array_equal(np.diff(offs), nm) and offs[0] == 0 and offs[-1] == pairs.shape[0]
    p, i1 = np.nonzero(m >= 0)
    assert np.array_equal(pairs[:, 0], i1) and np.array_equal(pairs[:, 1], m[p, i1])
    if pairs.shape[0] > 0:
      with pytest.raises(M.OrbGpuError):  # capacity too small is reported, not truncated silently
        mm.SearchForTriangulationPairs(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, out=(np.empty(n_pairs + 1, np.int32), np.empty((max(pairs.shape[0] - 1, 0), 2), np.int32)))


@pytest.mark.parametrize("seed,only_stereo,coarse,ori", [(521, 0, 0, 0), (522, 1, 0, 1), (523, 1, 1, 0), (524, 0, 0, 1)])
def test_triangulation_stereo(ctx, M, oracle, seed, only_stereo, coarse, ori):
    """key frames with mvuRight: served by the per-pair kernel (engine 1; auto must route there), engine 2 refuses"""
    from helpers import add_stereo
    tc = add_stereo(synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=40, n_feat=1200)), seed)
    ks = ctx.upload_kfset(tc.kfs)
    mm = M.ORBmatcher(0.6, bool(ori), ctx)
    oracle.reset_comparisons()
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, only_stereo, coarse, ori, n_threads=os.cpu_count() or 1)
    for eng in (0, 1):
        ctx.set_triangulation_engine(eng)
        nm, m = mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, bool(only_stereo), bool(coarse))
        assert np.array_equal(nm, enm) and np.array_equal(m, em)
        assert ctx.last_comparisons == oracle.comparisons()
    ctx.set_triangulation_engine(2)
    with pytest.raises(M.OrbGpuError):
        mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, bool(only_stereo), bool(coarse))
    ctx.set_triangulation_engine(0)
    # bOnlyStereo on a monocular set matches nothing (every feature fails :1136 / :1170)
    tm = synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=5, n_feat=600))
    nm, m = mm.SearchForTriangulation(ctx.upload_kfset(tm.kfs), tm.kf1, tm.kf2, tm.ep, tm.f12, True, False)
    assert not nm.any() and (m == -1).all()


@pytest.mark.parametrize("seed,th", [(321, 1.0), (322, 3.0)])
def test_projection_stereo(ctx, M, oracle, seed, th):
    from helpers import stereo_projection_case
    c = stereo_projection_case(seed, th)
    got = M.ORBmatcher(c.nnratio, True, ctx).SearchByProjection(ctx.upload_frame(c.frame), c.mps, th, False, 50.0, c.kp_prior_obs, c.kp_mp)
    oracle.reset_comparisons()
    exp = oracle.search_by_projection_local(c.frame, c.mps, th, 0, 50.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
    assert ctx.last_comparisons == oracle.comparisons()


@pytest.mark.parametrize("seed,n,all_level0", [(221, 7500, False), (222, 10000, False), (223, 9000, True), (224, 11000, True)])
def test_init_large_monocular(ctx, M, oracle, seed, n, all_level0):
    """monocular initialisation extracts 5 x nFeatures key points (Tracking.cc:667): 7.5 k - 10 k per frame.  With every key point
    on level 0 the per-partner state (9000) and also the per-key-point state (11000) no longer fit shared memory and live in
    global memory -- same fixed point, same results."""
    c = synth.make_init_case(seed, n=n, window_size=60 if all_level0 else 100)
    if all_level0:
        c.f1.octave[:] = 0
        c.f2.octave[:] = 0
    got = M.ORBmatcher(0.9, True, ctx).SearchForInitialization(ctx.upload_frame(c.f1), ctx.upload_frame(c.f2), c.prev_matched, c.window_size)
    oracle.reset_comparisons()
    exp = oracle.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, 0.9, 1)
    assert got[0] == exp[0] and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    assert ctx.last_comparisons == oracle.comparisons() and exp[0] > 100


def test_large_frames_shared_memory_opt_in(ctx, M, oracle):
    """frames with more than 12 k keypoints: the lock tables of the ordered passes exceed the 48 KB default shared memory"""
    frame, pts, kl = synth.make_projected_case(181, n_kp=14000, n_pts=4000, th=7.0)
    got = M.ORBmatcher(0.9, True, ctx).SearchProjected(ctx.upload_frame(frame), pts, 100.0, True, kl)
    exp = oracle.search_projected(frame, pts, 100.0, 1, kl, check_ori=1)
    assert got[0] == exp[0] and all(np.array_equal(a, b) for a, b in zip(got[1:], exp[1:]))
    c = synth.make_projection_case(182, n_kp=13000, n_mp=6000, th=3.0)
    g2 = M.ORBmatcher(c.nnratio, True, ctx).SearchByProjection(ctx.upload_frame(c.frame), c.mps, 3.0, False, 50.0, c.kp_prior_obs, c.kp_mp)
    e2 = oracle.search_by_projection_local(c.frame, c.mps, 3.0, 0, 50.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert g2[0] == e2[0] and np.array_equal(g2[1], e2[1])
    ci = synth.make_init_case(183, n=6000)
    f1, f2 = ctx.upload_frame(ci.f1), ctx.upload_frame(ci.f2)
    n, m12, prev = M.ORBmatcher(ci.nnratio, True, ctx).SearchForInitialization(f1, f2, ci.prev_matched, ci.window_size)
    en, em, ep = oracle.search_for_initialization(ci.f1, ci.f2, ci.prev_matched, ci.window_size, ci.nnratio, 1)
    assert n == en and np.array_equal(m12, em) and np.array_equal(prev, ep)


def test_triangulation_multi_target_output(ctx, M):
    """the fused all-gather entry point on one GPU: two target buffers (stand-ins for two ranks' result buffers), rows preset to -1,
    this 'rank' owning pairs [lo, hi) of a larger batch -- both targets must hold exactly the dense rows of the plain call"""
    import torch
    tc = synth.fill_geometry(synth.make_triangulation_case(191, n_pairs=300, n_feat=900))
    ks = ctx.upload_kfset(tc.kfs)
    mm = M.ORBmatcher(0.6, False, ctx)
    nm, m = mm.SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
    dev = torch.device("cuda", 0)
    P_total, NF, lo, hi = 300, 900, 100, 260
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kf1, kf2, ep, f12 = t(tc.kf1[lo:hi]), t(tc.kf2[lo:hi]), t(tc.ep[lo:hi]), t(tc.f12[lo:hi])
    for preset in (True, False, 2):
        rows = [torch.full((P_total, NF), -1 if preset is True else 7, dtype=torch.int32, device=dev) for _ in range(2)]
        cnts = [torch.full((P_total,), -5, dtype=torch.int32, device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        mm.SearchForTriangulation_peers_dev(ks, hi - lo, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(),
                                            [r.data_ptr() for r in rows], [c.data_ptr() for c in cnts], lo, preset)
        ctx.synchronize()
        for r, c in zip(rows, cnts):
            assert np.array_equal(r[lo:hi].cpu().numpy(), m[lo:hi]) and np.array_equal(c[lo:hi].cpu().numpy(), nm[lo:hi])
            other = -1 if preset is True else 7
            assert (r[:lo] == other).all() and (r[hi:] == other).all()  # other ranks' rows untouched
            assert (c[:lo] == -5).all() and (c[hi:] == -5).all()


@pytest.mark.parametrize("seed,n_pairs,n_feat,ori,dense", [(541, 300, 900, 0, 0), (542, 64, 2000, 1, 0), (543, 10, 1200, 0, 1), (544, 10, 1200, 1, 1),
                                                          (545, 2, 64, 0, 0)])
def test_triangulation_compact_gather(ctx, M, oracle, seed, n_pairs, n_feat, ori, dense):
    """the fused search + all-gather in the vMatchedPairs form on ONE GPU: two 'ranks' (one call each, each waiting only for its
    own epoch) own the two halves of the batch and store into both ranks' buffers; afterwards both buffers hold, for every pair,
    exactly the compact form of the oracle's dense rows (ascending idx1), both flag arrays carry both epochs, and a second step
    (epoch 2) lands the same way."""
    import torch
    from orb_slam3_comments_ghr_b200.sharding import compact_pairs_from_rows, pairs_from_compact
    tc = synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=n_pairs, n_feat=n_feat, n_nodes=12 if dense else 100))
    if dense:  # a handful of distinct descriptors: nodes with more than 32 surviving candidates -> the sBest / overflow path
        rng = np.random.default_rng(seed)
        base = synth.random_descriptors(rng, 3)
        pick = rng.integers(0, 3, size=tc.kfs.desc.shape[:2])
        tc.kfs.desc[:] = base[pick] ^ synth.flip_mask(rng, pick.size, np.full(pick.size, 6)).reshape(*pick.shape, 32)
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, 0, ori, n_threads=os.cpu_count() or 1)
    ecnt, eent = compact_pairs_from_rows(em)
    assert np.array_equal(ecnt, enm)
    ks = ctx.upload_kfset(tc.kfs)
    mm = M.ORBmatcher(0.6, bool(ori), ctx)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    half = n_pairs // 2
    bounds = [(0, half), (half, n_pairs)]
    pairs = [torch.full((n_pairs, n_feat), 0x5A5A5A5A, dtype=torch.int32, device=dev) for _ in range(2)]
    cnts = [torch.full((n_pairs,), -5, dtype=torch.int32, device=dev) for _ in range(2)]
    flags = [torch.zeros(8, dtype=torch.int32, device=dev) for _ in range(2)]
    state = [torch.tensor([1, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev) for _ in range(2)]
    ins = [[t(a[lo:hi]) for a in (tc.kf1, tc.kf2, tc.ep, tc.f12)] for lo, hi in bounds]
    for step in (1, 2):
        for r, (lo, hi) in enumerate(bounds):
            g = M.tri_gather_struct(r, [x.data_ptr() for x in pairs], [x.data_ptr() for x in cnts], [x.data_ptr() for x in flags],
                                    state[r].data_ptr(), state[r].data_ptr() + 16, wait_mask=1 << r)
            mm.SearchForTriangulation_gather_dev(ks, hi - lo, ins[r][0].data_ptr(), ins[r][1].data_ptr(), ins[r][2].data_ptr(),
                                                 ins[r][3].data_ptr(), g, lo)
        ctx.synchronize()
        for r in range(2):
            assert flags[r][:2].tolist() == [step, step] and state[r].tolist()[:5] == [step + 1, 0, 0, 0, 0]
            c = cnts[r].cpu().numpy()
            e = pairs[r].cpu().numpy().view(np.uint32)
            assert np.array_equal(c, ecnt)
            for p in range(n_pairs):
                assert np.array_equal(e[p, :c[p]], eent[p, :c[p]]), (step, r, p)
                pad = e[p, c[p]:(c[p] + 3) // 4 * 4]
                assert np.all(pad == 0xFFFFFFFF) and np.all(e[p, (c[p] + 3) // 4 * 4:] == 0x5A5A5A5A)
                assert np.array_equal(pairs_from_compact(c, e, p)[:, 0], np.flatnonzero(em[p] >= 0))
        # the gathered result on the host in the vMatchedPairs form of orbgpu_search_for_triangulation_batch_pairs
        offs, pr = mm.TriangulationGatherDownload(n_pairs, n_feat, cnts[1].data_ptr(), pairs[1].data_ptr())
        assert np.array_equal(offs, np.concatenate([[0], np.cumsum(ecnt)]).astype(np.int32))
        for p in range(n_pairs):
            i1 = np.flatnonzero(em[p] >= 0)
            assert np.array_equal(pr[offs[p]:offs[p + 1]], np.stack([i1, em[p, i1]], axis=1)), p
        for x in pairs:
            x.fill_(0x5A5A5A5A)
    assert int(enm.sum()) > 0


@pytest.mark.parametrize("levelsup", [2, 3])
def test_kfset_transform_then_triangulation(ctx, M, oracle, levelsup):
    """KeyFrame::ComputeBoW for a whole key-frame set on the device: FeatureVector nodes from the vocabulary descent feed the
    batched SearchForTriangulation without a host round trip; equals oracle transform + oracle search."""
    voc = golden_voc()
    rng = np.random.default_rng(201)
    tc = synth.fill_geometry(synth.make_triangulation_case(201, n_pairs=24, n_feat=800))
    nk, nf = tc.kfs.desc.shape[:2]
    # descriptors near vocabulary words so that the key frames share nodes; planted pairs keep (nearly) equal descriptors
    base = synth.descriptors_near_words(rng, voc, nk * nf).reshape(nk, nf, 32)
    tc.kfs.desc[:] = base
    tc.kfs.desc[tc.kf2] = np.where(rng.random((tc.kf2.shape[0], nf, 1)) < 0.5, tc.kfs.desc[tc.kf1] ^ synth.flip_mask(rng, tc.kf1.shape[0] * nf, np.full(tc.kf1.shape[0] * nf, 5)).reshape(-1, nf, 32), tc.kfs.desc[tc.kf2])
    # oracle: per-feature descent -> node ids (stopped words dropped)
    w, nid, wt = oracle.voc_transform(voc, tc.kfs.desc.reshape(-1, 32), levelsup)
    node = np.where(wt > 0, nid, np.uint32(0xFFFFFFFF)).astype(np.uint32).reshape(nk, nf)
    host_nodes = tc.kfs.node_id
    tc.kfs.node_id = None
    ks = ctx.upload_kfset(tc.kfs)
    dv = ctx.upload_vocabulary(voc)
    oracle.reset_comparisons()
    oracle.voc_transform(voc, tc.kfs.desc.reshape(-1, 32), levelsup)
    n_cmp = ks.transform(dv, levelsup)
    assert n_cmp == oracle.comparisons()
    nm, m = M.ORBmatcher(0.6, False, ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12, False, True)
    tc.kfs.node_id = node
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, 1, 0, n_threads=os.cpu_count() or 1)
    assert np.array_equal(nm, enm) and np.array_equal(m, em) and enm.sum() > 100
    tc.kfs.node_id = host_nodes


def test_knn2_full_size_engines_agree(ctx, M):
    """BASELINE.json config C5 at full size (262144 queries x 4194304 database rows, 1.1e12 comparisons): the tensor-core engine
    (tcgen05, +-1 fp8 contraction) and the LOP3+POPC engine return identical best index / best / second / match vectors, and the
    size-independent properties hold (planted queries find their source or an equal-distance earlier row, distances recompute)."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(4242)
    nq, nd = 262144, 4194304
    db = torch.randint(0, 256, (nd, 32), dtype=torch.uint8, device=dev, generator=g)
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
    n_pl = nq // 10
    who = torch.randperm(nq, device=dev, generator=g)[:n_pl]
    src = torch.randint(0, nd, (n_pl,), device=dev, generator=g)
    mask = torch.randint(0, 256, (n_pl, 32), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(3):
        mask &= torch.randint(0, 256, (n_pl, 32), dtype=torch.uint8, device=dev, generator=g)
    q[who] = db[src] ^ mask
    mm = M.ORBmatcher(0.8, True, ctx)
    ddb = ctx.database_from_device(db.data_ptr(), nd, keepalive=db)
    res = {}
    for eng in (4, 1):
        ctx.set_knn_engine(eng)
        out = torch.empty((4, nq), dtype=torch.int32, device=dev)
        mm.SearchByNN_dev(ddb, nq, q.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), 50)
        ctx.synchronize()
        res[eng] = out
    ctx.set_knn_engine(0)
    assert torch.equal(res[4], res[1])
    bi, bd, sd, mt = [x.long() for x in res[4]]
    assert bool((bd <= sd).all())
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
    got_d = lut[(q ^ db[bi]).long()].sum(dim=1)
    assert torch.equal(got_d.long(), bd)
    true_d = lut[(q[who] ^ db[src]).long()].sum(dim=1).long()
    assert bool((bd[who] <= true_d).all())
    assert float((bi[who] == src).float().mean()) > 0.999
    acc = (bd <= 50) & (bd.float() < 0.8 * sd.float())
    assert torch.equal(mt >= 0, acc) and torch.equal(mt[acc], bi[acc])
    assert int(acc.sum()) >= int(0.95 * n_pl)


@pytest.mark.parametrize("seed,n_mp", [(711, 5000), (712, 20000), (713, 1)])
def test_is_in_frustum(ctx, oracle, seed, n_mp):
    """row f1: Frame::isInFrustum (Frame.cc:676-782) on the device == the C restatement (pinned on the reference's own text): every
    member the function writes, bit for bit, incl. the predicted level (log taken in double on the device, logf on the host)"""
    from orb_slam3_comments_ghr_b200._abi import frustum_struct
    c = synth.make_frustum_case(seed, n_mp=n_mp)
    f = c.frame
    fr = frustum_struct(c.Rcw, c.tcw, c.Ow, c.K, c.mbf, (f.min_x, f.min_y, f.max_x, f.max_y), c.viewing_cos_limit, c.log_scale_factor, 8)
    g = ctx.is_in_frustum(fr, c.world_pos, c.normal, c.min_distance, c.max_distance)
    o = oracle.is_in_frustum(fr, c.world_pos, c.normal, c.min_distance, c.max_distance)
    for k in ("in_view", "proj_xy", "proj_xr", "depth", "scale_level", "view_cos"):
        assert np.array_equal(g[k], o[k]), k
    if n_mp > 100:
        assert 0.15 < o["in_view"].mean() < 0.8


@pytest.mark.parametrize("seed,th,far", [(721, 1.0, 0), (722, 3.0, 0), (723, 3.0, 1)])
def test_search_local_points_fused(ctx, M, oracle, seed, th, far):
    """Tracking::SearchLocalPoints on the device (isInFrustum + SearchByProjection, one call, projections never leave HBM) == the two
    host-visible steps of the oracle; also == the two GPU calls made separately"""
    from orb_slam3_comments_ghr_b200._abi import HostLocalPoints, HostMapPoints, frustum_struct
    c = synth.make_frustum_case(seed)
    f = c.frame
    fr = frustum_struct(c.Rcw, c.tcw, c.Ow, c.K, c.mbf, (f.min_x, f.min_y, f.max_x, f.max_y), c.viewing_cos_limit, c.log_scale_factor, 8)
    o = oracle.is_in_frustum(fr, c.world_pos, c.normal, c.min_distance, c.max_distance)
    inv = (o["in_view"] > 0) & (c.skip == 0)  # the loop of SearchLocalPoints does not test skipped points: mbTrackInView stays false
    mps = HostMapPoints(c.desc, o["proj_xy"], o["scale_level"], o["view_cos"], o["depth"], inv.astype(np.uint8), c.bad, c.n_obs, proj_xr=o["proj_xr"])
    en, ekp = oracle.search_by_projection_local(f, mps, th, far, 6.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    d = ctx.upload_frame(f)
    m = M.ORBmatcher(c.nnratio, True, ctx)
    pts = HostLocalPoints(c.desc, c.world_pos, c.normal, c.min_distance, c.max_distance, c.bad, c.n_obs, skip=c.skip)
    n, kp, giv = m.SearchLocalPoints(d, fr, pts, th, bool(far), 6.0, c.kp_prior_obs, c.kp_mp)
    assert n == en and np.array_equal(kp, ekp) and np.array_equal(giv, inv.astype(np.uint8))
    n2, kp2 = m.SearchByProjection(d, mps, th, bool(far), 6.0, c.kp_prior_obs, c.kp_mp)
    assert n2 == n and np.array_equal(kp2, kp) and en > 200


@pytest.mark.parametrize("levelsup,K", [(2, 12), (4, 3), (3, 1)])
def test_search_by_bow_batch(ctx, M, oracle, levelsup, K):
    """one frame against K candidate key frames in one call (the relocalisation loop, Tracking.cc:4469-4495) == K separate calls ==
    the oracle, incl. the comparison counter summed over the candidates"""
    voc = golden_voc()
    f_host = None
    kfs, valids, exp = [], [], []
    total_cmp = 0
    for k in range(K):
        c = synth.make_bow_case(2000 + 10 * levelsup + k, voc, 1500 if k % 2 else 2000)
        if f_host is None:
            f_host = attach_featvec(oracle, voc, c.f, levelsup)
        kf = attach_featvec(oracle, voc, c.kf, levelsup)
        if k > 0:  # candidates that really look like the frame: planted copies of its features
            rng = np.random.default_rng(k)
            src = rng.permutation(f_host.n)[: kf.n // 2]
            kf.desc[: src.size] = synth.planted_copies(rng, f_host.desc[src])
            kf = attach_featvec(oracle, voc, kf, levelsup)
        kfs.append(kf)
        valids.append(c.kf_mp_valid[: kf.n])
        oracle.reset_comparisons()
        exp.append(oracle.search_by_bow_kf_f(kf, f_host, valids[-1], 0.7, 1))
        total_cmp += oracle.comparisons()
    df = ctx.upload_frame(f_host)
    dks = [ctx.upload_frame(kf) for kf in kfs]
    m = M.ORBmatcher(0.7, True, ctx)
    nm, out = m.SearchByBoWBatch(dks, df, valids)
    assert ctx.last_comparisons == total_cmp
    for k in range(K):
        assert nm[k] == exp[k][0] and np.array_equal(out[k], exp[k][1]), k
        n1, m1 = m.SearchByBoW(dks[k], df, valids[k])
        assert n1 == nm[k] and np.array_equal(m1, out[k])
    assert int(nm.sum()) > 0


@pytest.mark.parametrize("levelsup,K", [(2, 6), (4, 5), (6, 2)])
def test_search_by_bow_batch_kf_kf(ctx, M, oracle, levelsup, K):
    """the current key frame against the K key frames of a candidate window in one call (LoopClosing.cc:909-925) == K separate
    SearchByBoW(KF, KF) calls == the oracle; window key frames of different sizes, levelsup 6 = one root node (the big-node path)"""
    voc = golden_voc()
    kf1 = None
    kf2s, valids, exp = [], [], []
    total_cmp = 0
    for k in range(K):
        c = synth.make_bow_case(2600 + 10 * levelsup + k, voc, 1200 if k % 2 else 1800)
        if kf1 is None:
            kf1 = attach_featvec(oracle, voc, c.kf, levelsup)
            v1 = c.kf_mp_valid[: kf1.n]
        f = c.f
        if k > 0:  # window key frames that really look like the current one
            rng = np.random.default_rng(100 + k)
            src = rng.permutation(kf1.n)[: f.n // 2]
            f.desc[: src.size] = synth.planted_copies(rng, kf1.desc[src])
        f = attach_featvec(oracle, voc, f, levelsup)
        kf2s.append(f)
        valids.append(c.f_mp_valid[: f.n])
        oracle.reset_comparisons()
        exp.append(oracle.search_by_bow_kf_kf(kf1, f, v1, valids[-1], 0.8, 1))
        total_cmp += oracle.comparisons()
    d1 = ctx.upload_frame(kf1)
    d2s = [ctx.upload_frame(f) for f in kf2s]
    m = M.ORBmatcher(0.8, True, ctx)
    nm, out = m.SearchByBoWBatchKF(d1, v1, d2s, valids)
    assert ctx.last_comparisons == total_cmp
    for k in range(K):
        assert nm[k] == exp[k][0] and np.array_equal(out[k], exp[k][1]), k
        n1, m1 = m.SearchByBoW(d1, d2s[k], v1, valids[k])
        assert n1 == nm[k] and np.array_equal(m1, out[k])
    assert int(nm.sum()) > 0
    nm0, out0 = m.SearchByBoWBatchKF(d1, v1, [], [])
    assert nm0.size == 0 and out0.shape[0] == 0


def test_triangulation_shared_keyframes(ctx, M, oracle):
    """C4 shared-key-frame variant (LocalMapping::CreateNewMapPoints: every key frame against its 8 best neighbours): both engines and
    the compact form against the oracle"""
    tc = synth.fill_geometry(synth.make_triangulation_case_shared(551, n_kf=24, n_neighbours=8, n_feat=1600))
    ks = ctx.upload_kfset(tc.kfs)
    enm, em = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, 0, 0, n_threads=os.cpu_count() or 1)
    for engine in (1, 2):
        ctx.set_triangulation_engine(engine)
        nm, m = M.ORBmatcher(0.6, False, ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
        ctx.set_triangulation_engine(0)
        assert np.array_equal(nm, enm) and np.array_equal(m, em), engine
    assert enm.shape[0] == 192 and enm[0] > 50 and enm[7] > 20  # nearer neighbours share more landmarks


def test_candidate_pool_regrow(M, oracle):
    """the candidate lists of the projection searches live in a pool sized for ~64 candidates per point; wide windows need more: the
    kernels count what they would have written, the host sees the count exceed the capacity and repeats the call with a pool of that
    size (twice the launches on the first call of a fresh context, none extra on the second).  Results equal the oracle's either way."""
    from orb_slam3_comments_ghr_b200 import matcher
    from orb_slam3_comments_ghr_b200._abi import HostLocalPoints, frustum_struct
    ctx = matcher.Context(0)  # a fresh context: no pool hint yet
    frame, pts, kl = synth.make_projected_case(191, n_kp=6000, n_pts=1500, th=60.0)
    d = ctx.upload_frame(frame)
    m = M.ORBmatcher(0.9, True, ctx)
    exp = oracle.search_projected(frame, pts, 100.0, 1, kl, check_ori=1)
    l0 = ctx.launch_count
    got = m.SearchProjected(d, pts, 100.0, True, kl)
    l1 = ctx.launch_count
    got2 = m.SearchProjected(d, pts, 100.0, True, kl)
    l2 = ctx.launch_count
    for g in (got, got2):
        assert g[0] == exp[0] and all(np.array_equal(a, b) for a, b in zip(g[1:], exp[1:]))
    assert l1 - l0 == 4 and l2 - l1 == 2, (l1 - l0, l2 - l1)
    # SearchByProjection(Frame, MapPoints): in/out F.mvpMapPoints must survive the repeated call
    ctx2 = matcher.Context(0)
    c = synth.make_projection_case(192, n_kp=8000, n_mp=3000, th=30.0)
    d2 = ctx2.upload_frame(c.frame)
    l0 = ctx2.launch_count
    g = M.ORBmatcher(c.nnratio, True, ctx2).SearchByProjection(d2, c.mps, 30.0, False, 50.0, c.kp_prior_obs, c.kp_mp)
    e = oracle.search_by_projection_local(c.frame, c.mps, 30.0, 0, 50.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert g[0] == e[0] and np.array_equal(g[1], e[1]) and ctx2.launch_count - l0 == 4


def test_round2_entry_points_edge_cases(ctx, M, oracle):
    """empty and degenerate inputs of the entry points added in round 2: no map points, an empty frame, every point skipped, an empty
    candidate list, an empty descriptor list"""
    import ctypes as C
    from orb_slam3_comments_ghr_b200._abi import HostLocalPoints, frustum_struct, u8p, u32p, f64p, i32p
    c = synth.make_frustum_case(731, n_kp=300, n_mp=200)
    f = c.frame
    fr = frustum_struct(c.Rcw, c.tcw, c.Ow, c.K, c.mbf, (f.min_x, f.min_y, f.max_x, f.max_y), c.viewing_cos_limit, c.log_scale_factor, 8)
    d = ctx.upload_frame(f)
    m = M.ORBmatcher(0.8, True, ctx)
    # no map points
    e = np.zeros((0,), dtype=np.float32)
    empty = HostLocalPoints(np.zeros((0, 32), np.uint8), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), e, e,
                            np.zeros(0, np.uint8), np.zeros(0, np.int32))
    n, kp, iv = m.SearchLocalPoints(d, fr, empty, 3.0, False, 50.0, c.kp_prior_obs, c.kp_mp)
    assert n == 0 and np.array_equal(kp, c.kp_mp) and iv.size == 0
    # every point skipped: nothing in view, nothing matched, F.mvpMapPoints untouched
    pts = HostLocalPoints(c.desc, c.world_pos, c.normal, c.min_distance, c.max_distance, c.bad, c.n_obs, skip=np.ones(c.desc.shape[0], np.uint8))
    n, kp, iv = m.SearchLocalPoints(d, fr, pts, 3.0, False, 50.0, c.kp_prior_obs, c.kp_mp)
    assert n == 0 and np.array_equal(kp, c.kp_mp) and not iv.any()
    g = ctx.is_in_frustum(fr, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), e, e)
    assert g["in_view"].size == 0
    # an empty candidate list for the batched SearchByBoW
    nm, out = m.SearchByBoWBatch([], d, [])
    assert nm.size == 0 and out.shape[0] == 0
    # transform of an empty descriptor list
    voc = ctx.upload_vocabulary(golden_voc())
    L = M.load_library()
    L.orbgpu_transform_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, u8p, C.c_int32, C.POINTER(C.c_int32), u32p, f64p,
                                               C.POINTER(C.c_int32), u32p, i32p, u32p]
    nw, nn = C.c_int32(7), C.c_int32(7)
    off = np.full(1, 5, dtype=np.int32)
    rc = L.orbgpu_transform_descriptors(ctx.handle, voc.handle, 0, C.cast(None, u8p), 4, C.byref(nw), C.cast(None, u32p), C.cast(None, f64p),
                                        C.byref(nn), C.cast(None, u32p), off.ctypes.data_as(i32p), C.cast(None, u32p))
    assert rc == 0 and nw.value == 0 and nn.value == 0 and off[0] == 0
