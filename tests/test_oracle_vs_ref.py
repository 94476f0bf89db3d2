"""CPU, only where oracle/_ref exists: the C restatement equals the reference's own
ORBmatcher.cc / DBoW2 (compiled unmodified) on fresh seeds and on the edge cases."""
import numpy as np
import pytest

from helpers import add_stereo, attach_featvec, golden_voc, stereo_projection_case
from orb_slam3_comments_ghr_b200 import synth
from orb_slam3_comments_ghr_b200._abi import HostFrame, HostMapPoints

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_grid_and_area(oracle, reference, seed):
    rng = np.random.default_rng(seed)
    f = synth.make_frame(rng, 1500)
    # exact cell-boundary and out-of-image keypoints
    f.kp_xy[:40, 0] = np.arange(40, dtype=np.float32) * 5.0 + 5.0  # x*0.1 = k+0.5 -> round-half-away ties
    f.kp_xy[40:44] = np.array([[639.75, 479.75], [0, 0], [636.0, 476.0], [635.0, 475.0]], dtype=np.float32)
    cs_o, ci_o = oracle.grid(f)
    cs_r, ci_r = reference.grid(f)
    assert np.array_equal(cs_o, cs_r) and np.array_equal(ci_o, ci_r)
    for _ in range(200):
        x, y = rng.uniform(-50, 700), rng.uniform(-50, 530)
        r = float(rng.choice([1.0, 2.5, 7.5, 40.0, 100.0, 1000.0]))
        lv = int(rng.integers(0, 8))
        mn, mx = [(-1, -1), (lv - 1, lv), (lv, lv), (0, lv), (lv, -1)][int(rng.integers(0, 5))]
        a = oracle.features_in_area(f, x, y, r, mn, mx, grid=(cs_o, ci_o))
        b = reference.features_in_area(f, x, y, r, mn, mx)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed,n,ratio,ori", [(201, 1000, 0.9, 1), (202, 1000, 0.6, 0), (203, 300, 0.9, 1), (204, 2500, 0.95, 1)])
def test_init(oracle, reference, seed, n, ratio, ori):
    c = synth.make_init_case(seed, n=n)
    a = oracle.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, ratio, ori)
    b = reference.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, ratio, ori)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("seed,th,far", [(301, 1.0, 0), (302, 3.0, 0), (303, 15.0, 1), (304, 2.0, 1)])
def test_projection(oracle, reference, seed, th, far):
    c = synth.make_projection_case(seed, th=th, far_points=far)
    a = oracle.search_by_projection_local(c.frame, c.mps, th, far, 40.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    b = reference.search_by_projection_local(c.frame, c.mps, th, far, 40.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("levelsup", [0, 1, 2, 3, 4, 6])
def test_transform_random_vocabulary(oracle, reference, levelsup):
    voc = synth.random_vocabulary(7, k=6, L=4, ragged=True)
    hv = reference.voc_from_flat(voc)
    rng = np.random.default_rng(levelsup)
    d = synth.descriptors_near_words(rng, voc, 700)
    r = hv.transform(d, levelsup)
    w, nid, wt = oracle.voc_transform(voc, d, levelsup)
    assert np.array_equal(w, r["word_id"]) and np.array_equal(nid, r["node_id"]) and np.array_equal(wt, r["weight"])
    bw, bv = oracle.bowvector(w, wt)
    assert np.array_equal(bw, r["bow_words"]) and np.array_equal(bv, r["bow_values"])
    fn, fo, ff = oracle.featvec(nid, wt)
    assert np.array_equal(fn, r["fv_node_ids"]) and np.array_equal(fo, r["fv_offsets"]) and np.array_equal(ff, r["fv_features"])
    assert (wt == 0).any(), "the random vocabulary must exercise stopped words"


@pytest.mark.parametrize("seed,levelsup,ratio", [(401, 2, 0.7), (402, 3, 0.75), (403, 4, 0.9)])
def test_bow(oracle, reference, seed, levelsup, ratio):
    voc = golden_voc()
    bc = synth.make_bow_case(seed, voc, 1200)
    kf = attach_featvec(oracle, voc, bc.kf, levelsup)
    f = attach_featvec(oracle, voc, bc.f, levelsup)
    for ori in (0, 1):
        a = oracle.search_by_bow_kf_f(kf, f, bc.kf_mp_valid, ratio, ori)
        b = reference.search_by_bow_kf_f(kf, f, bc.kf_mp_valid, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
        a = oracle.search_by_bow_kf_kf(kf, f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        b = reference.search_by_bow_kf_kf(kf, f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("seed,coarse,ori", [(501, 0, 0), (502, 0, 1), (503, 1, 1)])
def test_triangulation(oracle, reference, seed, coarse, ori):
    tc = synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=6, n_feat=1500))
    for p in range(6):
        ep, f12 = reference.triangulation_geometry(tc.T1w[p], tc.T2w[p], tc.K, tc.K)
        assert np.array_equal(ep, tc.ep[p]) and np.array_equal(f12, tc.f12[p])
    a = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, coarse, ori, n_threads=2)
    b = reference.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.T1w, tc.T2w, tc.K, 0, coarse, ori, 0.6, n_threads=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_edge_cases(oracle, reference):
    rng = np.random.default_rng(9)
    # empty frames / no candidates / single candidate (INT_MAX second best in init)
    f1 = synth.make_frame(rng, 3)
    f2 = synth.make_frame(rng, 1)
    f1.octave[:] = 0
    f2.octave[:] = 0
    f2.kp_xy[0] = f1.kp_xy[0]
    f2.desc[0] = f1.desc[0]
    a = oracle.search_for_initialization(f1, f2, f1.kp_xy.copy(), 100, 0.9, 1)
    b = reference.search_for_initialization(f1, f2, f1.kp_xy.copy(), 100, 0.9, 1)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] >= 1


# ---- row a6: the search core of the self-projecting overloads against the reference's OWN functions.  The harness
# (oracle/ref_adapter.cc) gives them identity poses and a unit pinhole so that their prologue reproduces the given
# projections exactly; everything from GetFeaturesInArea on is the reference's code.
def _frustum(pts, frame, is_in_image):
    """the image-bounds gate belongs to the caller's prologue: Frame callers use u < min || u > max (:2013-2016,
    :2226-2229), KeyFrame callers IsInImage: min <= u < max (KeyFrame.cc:910-913)"""
    u, v = pts.uv[:, 0], pts.uv[:, 1]
    if is_in_image:
        ok = (u >= frame.min_x) & (u < frame.max_x) & (v >= frame.min_y) & (v < frame.max_y)
    else:
        ok = ~((u < frame.min_x) | (u > frame.max_x) | (v < frame.min_y) | (v > frame.max_y))
    pts.active = (pts.active.astype(bool) & ok).astype(np.uint8)
    return pts


@pytest.mark.parametrize("seed", [81, 82])
@pytest.mark.parametrize("mode,level_mode,stereo", [(0, "pm1", False), (1, "fwd", True), (2, "bwd", False), (0, "pm1", True)])
@pytest.mark.parametrize("ori", [0, 1])
def test_projected_cur_last(oracle, reference, seed, mode, level_mode, stereo, ori):
    th, mbf = (15.0 if stereo else 7.0), 40.0
    frame, pts, kl = synth.make_projected_case(seed, n_kp=1200, n_pts=1500, th=th, stereo=stereo, level_mode=level_mode, lock_frac=0.85)
    pts.ur = (pts.uv[:, 0] - np.float32(mbf)).astype(np.float32)  # what the reference derives from mbf * invzc (:2056)
    exp_n, _ = reference.projected_cur_last(frame, pts, th, mode, mbf, kl, ori)
    pts = _frustum(pts, frame, False)
    exp_n, exp_own = reference.projected_cur_last(frame, pts, th, mode, mbf, kl, ori)
    got = oracle.search_projected(frame, pts, 100.0, 1, kl, stereo_gate=1, check_ori=ori)
    assert exp_n > 100
    assert got[0] == exp_n and np.array_equal(got[3], exp_own)


@pytest.mark.parametrize("seed", [83, 84])
@pytest.mark.parametrize("ori", [0, 1])
def test_projected_reloc(oracle, reference, seed, ori):
    frame, pts, kl = synth.make_projected_case(seed, n_kp=1200, n_pts=1500, th=10.0, level_mode="pm1", lock_frac=1.0)
    pts = _frustum(pts, frame, False)
    exp_n, exp_own = reference.projected_reloc(frame, pts, 10.0, 64, kl, ori)
    got = oracle.search_projected(frame, pts, 64.0, 1, kl, check_ori=ori)
    assert exp_n > 100
    assert got[0] == exp_n and np.array_equal(got[3], exp_own)


@pytest.mark.parametrize("seed,ratio", [(85, 1.0), (86, 0.9)])
def test_projected_sim3(oracle, reference, seed, ratio):
    frame, pts, kl = synth.make_projected_case(seed, n_kp=1200, n_pts=1500, th=8.0, level_mode="pred", lock_frac=1.0)
    pts = _frustum(pts, frame, True)
    exp_n, exp_own = reference.projected_sim3(frame, pts, 8, ratio, kl)
    got = oracle.search_projected(frame, pts, float(np.float32(50) * np.float32(ratio)), 1, kl)
    assert exp_n > 100
    assert got[0] == exp_n and np.array_equal(got[3], exp_own)


@pytest.mark.parametrize("seed,stereo", [(87, False), (88, True)])
def test_projected_fuse(oracle, reference, seed, stereo):
    bf = 40.0
    frame, pts, _ = synth.make_projected_case(seed, n_kp=1200, n_pts=1500, th=3.0, stereo=stereo, level_mode="pred")
    pts.ur = (pts.uv[:, 0] - np.float32(bf)).astype(np.float32)  # ur = uv(0) - bf * invz (:1398)
    inv = (np.float32(1.0) / frame.level_sigma2).astype(np.float32)
    pts = _frustum(pts, frame, True)
    exp_n, exp_bi = reference.projected_fuse(frame, pts, 3.0, bf)
    got = oracle.search_projected(frame, pts, 50.0, 0, None, chi2_gate=1, inv_level_sigma2=inv)
    assert exp_n > 100
    assert got[0] == exp_n and np.array_equal(got[1], exp_bi)
    exp_n, exp_bi = reference.projected_fuse_sim3(frame, pts, 3.0)
    got = oracle.search_projected(frame, pts, 50.0, 0, None)
    assert got[0] == exp_n and np.array_equal(got[1], exp_bi)


@pytest.mark.parametrize("seed", [111, 112])
def test_bow_score_l1(oracle, reference, seed):
    """8(f) rank 2: the restated L1Scoring::score / common-word count against the reference's ScoringObject.cpp"""
    db, qw, qv = synth.make_bowdb_case(seed, n_kf=600)
    ec, es = reference.bow_score_l1(db, qw, qv)
    gc, gs = oracle.bow_score_l1(db, qw, qv)
    assert np.array_equal(gc, ec) and np.array_equal(gs.view(np.uint64), es.view(np.uint64))
    assert es.max() > 0.3 and (ec == 0).any()
    ec, es = reference.bow_score_l1(db, qw[:0], qv[:0])  # empty query
    gc, gs = oracle.bow_score_l1(db, qw[:0], qv[:0])
    assert np.array_equal(gc, ec) and np.array_equal(gs.view(np.uint64), es.view(np.uint64))


@pytest.mark.parametrize("seed,only_stereo,coarse,ori", [(511, 0, 0, 0), (512, 1, 0, 1), (513, 1, 1, 0), (514, 0, 0, 1)])
def test_triangulation_stereo(oracle, reference, seed, only_stereo, coarse, ori):
    """mvuRight >= 0 features: bOnlyStereo filter (:1134-1138, :1168-1172) and the epipole gate only for mono-mono pairs (:1191)"""
    tc = add_stereo(synth.fill_geometry(synth.make_triangulation_case(seed, n_pairs=6, n_feat=1500)), seed)
    a = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, only_stereo, coarse, ori, n_threads=2)
    b = reference.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.T1w, tc.T2w, tc.K, only_stereo, coarse, ori, 0.6, n_threads=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].sum() > 50


@pytest.mark.parametrize("seed,th", [(311, 1.0), (312, 3.0)])
def test_projection_stereo(oracle, reference, seed, th):
    c = stereo_projection_case(seed, th)
    a = oracle.search_by_projection_local(c.frame, c.mps, th, 0, 40.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    b = reference.search_by_projection_local(c.frame, c.mps, th, 0, 40.0, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 100


def _clip_long_lists(offs, desc, cap):
    """the reference keeps float Distances[N][N] on the stack (MapPoint.cc:492): lists are clipped to what 8 MB holds"""
    keep, new_offs = [], [0]
    for p in range(offs.shape[0] - 1):
        s, e = int(offs[p]), min(int(offs[p + 1]), int(offs[p]) + cap)
        keep.append(np.arange(s, e))
        new_offs.append(new_offs[-1] + e - s)
    return np.array(new_offs, dtype=np.int32), np.ascontiguousarray(desc[np.concatenate(keep)])


@pytest.mark.parametrize("seed", [131, 132, 133])
def test_compute_distinctive_descriptors(oracle, reference, seed):
    """8(f) rank 4: the restatement against the reference's own MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc compiled
    unmodified with its real MapPoint.h, oracle/_ref/libref_mappoint.so) -- same kept descriptor, same index among equal
    medians (first), nothing selected for an empty list"""
    if not reference.mappoint_available():
        pytest.skip("oracle/_ref/libref_mappoint.so not built")
    offs, desc = _clip_long_lists(*synth.make_distinctive_case(seed, n_mp=600, max_obs=40), cap=900)
    ei, ed = reference.compute_distinctive_descriptors(offs, desc)
    gi, gm = oracle.compute_distinctive_descriptors(offs, desc)
    assert np.array_equal(gi, ei)
    sel = gi >= 0
    assert sel.sum() > 500 and (~sel).any()
    assert np.array_equal(desc[offs[:-1][sel] + gi[sel]], ed[sel])
    assert (np.diff(offs) > 800).any()  # the long list went through both


def test_compute_distinctive_descriptors_bad_keyframes(oracle, reference):
    """observations of bad key frames are skipped (:471): the reference on the full lists == the restatement on the filtered ones"""
    if not reference.mappoint_available():
        pytest.skip("oracle/_ref/libref_mappoint.so not built")
    offs, desc = _clip_long_lists(*synth.make_distinctive_case(134, n_mp=300, max_obs=25), cap=200)
    rng = np.random.default_rng(7)
    bad = (rng.random(desc.shape[0]) < 0.3).astype(np.uint8)
    bad[offs[5]:offs[6]] = 1  # a point whose observations are all bad keeps its (empty) descriptor
    ei, ed = reference.compute_distinctive_descriptors(offs, desc, kf_bad=bad)
    good = np.flatnonzero(bad == 0)
    f_offs = np.searchsorted(good, offs).astype(np.int32)
    gi, _ = oracle.compute_distinctive_descriptors(f_offs, np.ascontiguousarray(desc[good]))
    for p in range(offs.shape[0] - 1):
        if gi[p] < 0:
            assert ei[p] == -1
        else:
            assert good[f_offs[p] + gi[p]] == offs[p] + ei[p]
    assert ei[5] == -1


@pytest.mark.parametrize("seed,layout,ratio", [(601, "root", 0.95), (602, "mixed", 0.9), (603, "root", 0.7)])
def test_bow_conflicts(oracle, reference, seed, layout, ratio):
    """the "partner already matched" rule under heavy contention (groups of near-duplicate partners): restatement == reference"""
    bc = synth.make_bow_conflict_case(seed, layout=layout)
    for ori in (0, 1):
        exp = reference.search_by_bow_kf_f(bc.kf, bc.f, bc.kf_mp_valid, ratio, ori)
        got = oracle.search_by_bow_kf_f(bc.kf, bc.f, bc.kf_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
        exp = reference.search_by_bow_kf_kf(bc.kf, bc.f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        got = oracle.search_by_bow_kf_kf(bc.kf, bc.f, bc.kf_mp_valid, bc.f_mp_valid, ratio, ori)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1])
    assert exp[0] > 20


# ---------------------------------------------------------------------------------------------------------------------------------
# The helpers whose bodies are the reference's OWN TEXT, cut out of Frame.cc / KeyFrame.cc / MapPoint.cc / Pinhole.cpp by line range at
# build time (oracle/extract_ref.py): the C restatement must agree with them bit for bit.
@pytest.mark.parametrize("seed", [111, 112])
def test_keyframe_area_and_is_in_image(oracle, reference, seed):
    """KeyFrame::GetFeaturesInArea / IsInImage with the key frame's int bounds and the grid copied from its Frame (KeyFrame.cc:66-82)"""
    rng = np.random.default_rng(seed)
    f = synth.make_frame(rng, 1500)
    f.min_x, f.min_y, f.max_x, f.max_y = 0.0, 0.0, 640.0, 480.0
    cs, ci = oracle.grid(f)
    for _ in range(300):
        x, y = float(rng.uniform(-40, 690)), float(rng.uniform(-40, 520))
        if rng.random() < 0.2:
            x, y = float(rng.choice([0.0, 640.0, 639.75])), float(rng.choice([0.0, 480.0, 479.75]))
        r = float(rng.choice([1.0, 3.0, 7.5, 60.0]))
        a = oracle.features_in_area(f, x, y, r, -1, -1, grid=(cs, ci))
        b, inimg = reference.keyframe_features_in_area(f, x, y, r)
        assert np.array_equal(a, b)
        assert inimg == (x >= 0 and x < 640 and y >= 0 and y < 480)  # what the adapter's prologue evaluates (ORBmatcher.hpp)


def test_predict_scale_and_distance_invariance(oracle, reference):
    rng = np.random.default_rng(5)
    n = 20000
    lsf = float(np.log(np.float32(1.2)))
    mx = rng.uniform(0.5, 60, n).astype(np.float32)
    mn = (mx / 3.5).astype(np.float32)
    cd = (mx / (np.float32(1.2) ** rng.integers(-2, 10, n)) * rng.choice(np.array([1.0, 1.0, 1.0000001, 0.9999999, 1.01], dtype=np.float32), n)).astype(np.float32)
    lk, lf, a, b = reference.predict_scale(mx, mn, cd, lsf, 8)
    o = oracle.predict_scale(mx, cd, lsf, 8)
    assert np.array_equal(lk, o) and np.array_equal(lf, o)
    assert np.array_equal(a, np.float32(0.8) * mn) and np.array_equal(b, np.float32(1.2) * mx)
    assert set(np.unique(o)) == set(range(8))


def test_pinhole_project_and_epipolar_constrain(oracle, reference):
    rng = np.random.default_rng(6)
    K = np.array([synth.FX, synth.FY, synth.CX, synth.CY], dtype=np.float32)
    xyz = np.stack([rng.uniform(-5, 5, 5000), rng.uniform(-4, 4, 5000), rng.uniform(0.2, 15, 5000)], axis=1).astype(np.float32)
    assert np.array_equal(oracle.pinhole_project(K, xyz), reference.pinhole_project(K, xyz))
    tc = synth.fill_geometry(synth.make_triangulation_case(17, n_pairs=6, n_feat=400))
    for p in range(6):
        R1, t1 = tc.T1w[p][:9].reshape(3, 3).astype(np.float64), tc.T1w[p][9:12].astype(np.float64)
        R2, t2 = tc.T2w[p][:9].reshape(3, 3).astype(np.float64), tc.T2w[p][9:12].astype(np.float64)
        R12, t12 = (R1 @ R2.T).astype(np.float32), (t1 - R1 @ R2.T @ t2).astype(np.float32)  # any pair of floats: both sides consume them
        k1, k2 = tc.kf1[p], tc.kf2[p]
        a, b = tc.kfs.kp_xy[k1], tc.kfs.kp_xy[k2][rng.permutation(400)]
        b[:200] = tc.kfs.kp_xy[k2][:200]  # planted correspondences next to random pairs
        unc = tc.kfs.level_sigma2[tc.kfs.octave[k2]]
        ok, f12 = reference.epipolar_constrain(K, K, R12, t12, a, b, unc)
        assert np.array_equal(ok, oracle.epipolar_constrain(f12, a, b, unc))
        assert 0 < ok.sum() < ok.size


@pytest.mark.parametrize("seed", [701, 702, 703])
def test_is_in_frustum(oracle, reference, seed):
    """Frame::isInFrustum (Frame.cc:676-782): every member the function writes, for points that fail each gate in turn"""
    from orb_slam3_comments_ghr_b200._abi import frustum_struct
    c = synth.make_frustum_case(seed)
    f = c.frame
    ret, r = reference.is_in_frustum(f, c.Tcw34, c.K, c.mbf, c.viewing_cos_limit, c.world_pos, c.normal, c.min_distance, c.max_distance)
    fr = frustum_struct(c.Rcw, c.tcw, c.Ow, c.K, c.mbf, (f.min_x, f.min_y, f.max_x, f.max_y), c.viewing_cos_limit, c.log_scale_factor, 8)
    o = oracle.is_in_frustum(fr, c.world_pos, c.normal, c.min_distance, c.max_distance)
    assert np.array_equal(ret, r["in_view"])
    for k in ("in_view", "proj_xy"):
        assert np.array_equal(o[k], r[k]), k
    v = r["in_view"] > 0  # the other members are only written for accepted points
    for k in ("proj_xr", "depth", "scale_level", "view_cos"):
        assert np.array_equal(o[k][v], r[k][v]), k
    assert 0.15 < v.mean() < 0.8 and (r["proj_xy"][~v, 0] >= 0).any() and (r["proj_xy"][~v, 0] < 0).any()
    assert len(np.unique(r["scale_level"][v])) >= 6


@pytest.mark.parametrize("seed,n", [(801, 2000), (802, 500), (803, 1)])
def test_stereo_coarse_match_reference(oracle, reference, seed, n):
    """row f3: the coarse stage of Frame::ComputeStereoMatches, the reference's own text (Frame.cc:1117-1247).  The reference indexes
    vRowIndices with rows outside the image (undefined behaviour), so the case is shifted 16 rows down inside a taller image."""
    left, right, n_rows, mb, mbf = synth.make_stereo_case(seed, n=n)
    for fr in (left, right):
        fr.kp_xy[:, 1] += 16.0
    n_rows += 32
    a = oracle.stereo_coarse_match(left, right, n_rows, mb, mbf)
    b = reference.stereo_coarse_match(left, right, n_rows, mb, mbf)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    if n >= 500:
        assert (a[0] >= 0).mean() > 0.3
