"""CPU: the plain-C oracle reproduces the golden vectors generated from the reference itself
(oracle/_ref, scripts/make_golden.py) -- this is what pins the oracle (no GPU needed)."""
import numpy as np
import pytest

from helpers import GOLD, attach_featvec, digest, golden_cases, golden_outputs, golden_voc, projected_cases, projected_golden
from orb_slam3_comments_ghr_b200 import synth


def test_descriptor_distance_kats(oracle):
    z = np.load(f"{GOLD}/descriptor_distance.npz")
    got = np.array([oracle.descriptor_distance(z["a"][i], z["b"][i]) for i in range(z["a"].shape[0])])
    assert np.array_equal(got, z["dist"])
    assert got[0] == 256 and got[1] == 0 and (got[2:34] == 1).all()
    # against numpy popcount
    ref = np.unpackbits(z["a"] ^ z["b"], axis=1).sum(axis=1)
    assert np.array_equal(got, ref)


def test_three_maxima(oracle):
    z = np.load(f"{GOLD}/three_maxima.npz")
    for h, ind in zip(z["histo"], z["ind"]):
        assert np.array_equal(oracle.compute_three_maxima(h), ind)


@pytest.mark.parametrize("name", ["init_s11", "init_s12_n5000"])
def test_search_for_initialization(oracle, name):
    g = golden_outputs()
    c = golden_cases()[name]()
    assert np.array_equal(digest(c.f1.desc, c.f1.kp_xy, c.f2.desc, c.f2.kp_xy, c.f2.octave, c.f2.angle), g[name + "/in"]), "generator drift"
    n, m, prev = oracle.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, c.nnratio, c.check_ori)
    assert n == int(g[name + "/nmatches"])
    assert np.array_equal(m, g[name + "/matches12"])
    assert np.array_equal(prev, g[name + "/prev"])
    assert n == int((m >= 0).sum())


@pytest.mark.parametrize("name", ["proj_s21_th1", "proj_s22_th3", "proj_s23_far"])
def test_search_by_projection(oracle, name):
    g = golden_outputs()
    c = golden_cases()[name]()
    assert np.array_equal(digest(c.frame.desc, c.frame.kp_xy, c.mps.desc, c.mps.proj_xy, c.mps.n_obs), g[name + "/in"]), "generator drift"
    n, k = oracle.search_by_projection_local(c.frame, c.mps, c.th, c.far_points, c.th_far, c.nnratio, c.kp_prior_obs, c.kp_mp)
    assert n == int(g[name + "/nmatches"])
    assert np.array_equal(k, g[name + "/kp_mp"])


@pytest.mark.parametrize("levelsup", [2, 4])
def test_transform_and_bow(oracle, levelsup):
    g = golden_outputs()
    voc = golden_voc()
    bc = synth.make_bow_case(31, voc, 2000)
    assert np.array_equal(digest(bc.kf.desc, bc.f.desc, bc.kf.angle, bc.f.angle, bc.kf_mp_valid), g["bow_s31/in"]), "generator drift"
    frames = {}
    for side, fr in (("kf", bc.kf), ("f", bc.f)):
        w, nid, wt = oracle.voc_transform(voc, fr.desc, levelsup)
        p = f"bow_s31/l{levelsup}/{side}/"
        assert np.array_equal(w, g[p + "word_id"]) and np.array_equal(nid, g[p + "node_id"]) and np.array_equal(wt, g[p + "weight"])
        bw, bv = oracle.bowvector(w, wt)
        assert np.array_equal(bw, g[p + "bow_words"])
        assert np.array_equal(bv, g[p + "bow_values"]), "BowVector must be bit-exact in double"
        fn, fo, ff = oracle.featvec(nid, wt)
        assert np.array_equal(fn, g[p + "fv_node_ids"]) and np.array_equal(fo, g[p + "fv_offsets"]) and np.array_equal(ff, g[p + "fv_features"])
        frames[side] = fr.with_featvec(fn, fo, ff)
    n, m = oracle.search_by_bow_kf_f(frames["kf"], frames["f"], bc.kf_mp_valid, 0.7, 1)
    assert n == int(g[f"bow_s31/l{levelsup}/kf_f/nmatches"]) and np.array_equal(m, g[f"bow_s31/l{levelsup}/kf_f/match"])
    n, m = oracle.search_by_bow_kf_kf(frames["kf"], frames["f"], bc.kf_mp_valid, bc.f_mp_valid, 0.9, 1)
    assert n == int(g[f"bow_s31/l{levelsup}/kf_kf/nmatches"]) and np.array_equal(m, g[f"bow_s31/l{levelsup}/kf_kf/match"])


@pytest.mark.parametrize("check_ori", [0, 1])
def test_search_for_triangulation(oracle, check_ori):
    g = golden_outputs()
    tc = synth.fill_geometry(synth.make_triangulation_case(41, n_pairs=8, n_feat=2000))
    assert np.array_equal(digest(tc.kfs.desc, tc.kfs.kp_xy, tc.kfs.node_id, tc.kfs.has_mp, tc.T1w, tc.T2w), g["tri_s41/in"]), "generator drift"
    # the numpy restatement of the host pose algebra equals what the reference build computed
    assert np.array_equal(tc.ep, g["tri_s41/ep"]) and np.array_equal(tc.f12, g["tri_s41/f12"])
    nm, m = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12, 0, 0, check_ori, n_threads=4)
    assert np.array_equal(nm, g[f"tri_s41/ori{check_ori}/nmatches"])
    assert np.array_equal(m, g[f"tri_s41/ori{check_ori}/matches"].astype(np.int32))
    assert (nm > 100).all()


def test_knn2(oracle):
    g = golden_outputs()
    kc = synth.make_knn_case(51, 512, 20000)
    assert np.array_equal(digest(kc.q, kc.db), g["knn_s51/in"]), "generator drift"
    bi, bd, sd, mt = oracle.knn2_ratio(kc.q, kc.db, kc.th_low, kc.nnratio, 4)
    assert np.array_equal(bi, g["knn_s51/best_idx"]) and np.array_equal(bd, g["knn_s51/best_dist"])
    assert np.array_equal(sd, g["knn_s51/second_dist"]) and np.array_equal(mt, g["knn_s51/match"])
    assert oracle.comparisons() >= 0


@pytest.mark.parametrize("name", ["curlast_pm1", "curlast_fwd_stereo", "curlast_bwd", "reloc", "sim3", "fuse_stereo", "fuse_sim3"])
def test_search_projected(oracle, name):
    """row a6: the search core against the reference's own SearchByProjection(Cur,Last) / (Cur,KF) / (KF,Sim3) / Fuse x2 outputs"""
    g = projected_golden()
    mk, kind, _, kw = projected_cases()[name]
    frame, pts, kl = mk()
    assert np.array_equal(digest(frame.desc, frame.kp_xy, pts.desc, pts.uv, pts.radius, pts.active, kl), g[name + "/in"]), "generator drift"
    kw = dict(kw)
    if kw.get("chi2_gate"):
        kw["inv_level_sigma2"] = (np.float32(1.0) / frame.level_sigma2).astype(np.float32)
    ordered = kw.pop("ordered")
    n, bi, bd, own = oracle.search_projected(frame, pts, kw.pop("max_dist"), ordered, kl if ordered else None, **kw)
    assert n == int(g[name + "/nmatches"]) and n > 100
    if name + "/kp_owner" in g:
        assert np.array_equal(own, g[name + "/kp_owner"])
    else:
        assert np.array_equal(bi, g[name + "/best_idx"])


def test_compute_distinctive_descriptors_vs_numpy(oracle):
    """8(f) rank 4: besides the pin on the reference's MapPoint.cc (test_oracle_vs_ref.py), an independent numpy evaluation of
    MapPoint.cc:487-515 (all-pairs popcount, np.sort, index int(0.5*(N-1)), first minimum) that also runs where oracle/_ref is absent."""
    offs, desc = synth.make_distinctive_case(131, n_mp=400, max_obs=30)
    bi, bm = oracle.compute_distinctive_descriptors(offs, desc)
    bits = np.unpackbits(desc, axis=1).astype(np.int32)
    for p in range(offs.shape[0] - 1):
        s, e = offs[p], offs[p + 1]
        if e == s:
            assert bi[p] == -1
            continue
        if e - s > 200:
            continue
        b = bits[s:e]
        dist = (b[:, None, :] != b[None, :, :]).sum(axis=2)
        med = np.sort(dist, axis=1)[:, int(0.5 * (e - s - 1))]
        assert bi[p] == int(np.argmin(med)) and bm[p] == int(med.min())


@pytest.mark.parametrize("seed,n,min_acc", [(151, 600, 100), (152, 900, 150), (153, 37, 1), (154, 1, 0)])
def test_stereo_coarse_match_vs_numpy(oracle, seed, n, min_acc):
    """8(f) rank 3: Frame.cc does not compile here; the restatement of Frame.cc:1139-1216 is cross-checked with an independent
    numpy evaluation (row bands, octave and disparity filters, first minimum, acceptance threshold)."""
    left, right, n_rows, mb, mbf = synth.make_stereo_case(seed, n=n)
    bi, bd = oracle.stereo_coarse_match(left, right, n_rows, mb, mbf)
    r = np.float32(2.0) * left.scale_factors[right.octave]
    maxr = np.ceil(right.kp_xy[:, 1] + r).astype(int)
    minr = np.floor(right.kp_xy[:, 1] - r).astype(int)
    max_d = np.float32(mbf) / np.float32(mb)
    dist = np.unpackbits(left.desc[:, None, :] ^ right.desc[None, :, :], axis=2).sum(axis=2)
    n_acc = 0
    for i in range(left.n):
        row = int(left.kp_xy[i, 1])
        ok = (minr <= row) & (row <= maxr) & (np.abs(right.octave - left.octave[i]) <= 1)
        ok &= (right.kp_xy[:, 0] >= left.kp_xy[i, 0] - max_d) & (right.kp_xy[:, 0] <= left.kp_xy[i, 0])
        d = np.where(ok, dist[i], 1000)
        j = int(np.argmin(d))  # first minimum == ascending right index
        best = int(d[j]) if d[j] < 100 else 100
        assert bd[i] == best
        assert bi[i] == (j if best < 75 else -1)
        n_acc += bi[i] >= 0
    assert n_acc >= min_acc
