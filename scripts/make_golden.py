#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref: the reference's own
ORBmatcher.cc + DBoW2 compiled unmodified, see oracle/Makefile).

Run in the container where /root/reference is mounted:
    make -C oracle && python scripts/make_golden.py
The fixtures hold the reference's outputs on the seeded synthetic inputs of
orb_slam3_comments_ghr_b200.synth plus a checksum of those inputs (so generator drift is
detected), and the synthetic vocabulary built by the reference's TemplatedVocabulary::create.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Reference  # noqa: E402
from orb_slam3_comments_ghr_b200 import synth  # noqa: E402
from orb_slam3_comments_ghr_b200._abi import HostVoc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def golden_cases():
    """(name, builder) pairs shared with tests/test_golden.py"""
    return {
        "init_s11": lambda: synth.make_init_case(11),
        "init_s12_n5000": lambda: synth.make_init_case(12, n=5000),
        "proj_s21_th1": lambda: synth.make_projection_case(21, th=1.0),
        "proj_s22_th3": lambda: synth.make_projection_case(22, th=3.0),
        "proj_s23_far": lambda: synth.make_projection_case(23, th=5.0, far_points=1),
    }


def projected_cases():
    """row a6 (self-projecting overloads): name -> (builder of (frame, pts, kp_locked), reference call spec).  Shared with the
    tests.  The image-bounds gate of the reference's prologue is applied to `active` here (see tests/test_oracle_vs_ref.py)."""
    def build(seed, th, level_mode, stereo, lock_frac, in_image, off=None):
        frame, pts, kl = synth.make_projected_case(seed, n_kp=1500, n_pts=2000, th=th, stereo=stereo, level_mode=level_mode,
                                                   lock_frac=lock_frac)
        if off is not None:
            pts.ur = (pts.uv[:, 0] - np.float32(off)).astype(np.float32)
        u, v = pts.uv[:, 0], pts.uv[:, 1]
        if in_image:
            ok = (u >= frame.min_x) & (u < frame.max_x) & (v >= frame.min_y) & (v < frame.max_y)
        else:
            ok = ~((u < frame.min_x) | (u > frame.max_x) | (v < frame.min_y) | (v > frame.max_y))
        pts.active = (pts.active.astype(bool) & ok).astype(np.uint8)
        return frame, pts, kl
    return {
        # name: (builder, kind, reference args, search_projected kwargs)
        "curlast_pm1": (lambda: build(91, 7.0, "pm1", False, 0.85, False, 40.0), "cur_last", dict(th=7.0, mode=0, mbf=40.0, check_ori=1),
                        dict(max_dist=100.0, ordered=1, stereo_gate=1, check_ori=1)),
        "curlast_fwd_stereo": (lambda: build(92, 15.0, "fwd", True, 0.85, False, 40.0), "cur_last", dict(th=15.0, mode=1, mbf=40.0, check_ori=1),
                               dict(max_dist=100.0, ordered=1, stereo_gate=1, check_ori=1)),
        "curlast_bwd": (lambda: build(93, 15.0, "bwd", False, 0.85, False, 40.0), "cur_last", dict(th=15.0, mode=2, mbf=40.0, check_ori=0),
                        dict(max_dist=100.0, ordered=1, stereo_gate=1, check_ori=0)),
        "reloc": (lambda: build(94, 10.0, "pm1", False, 1.0, False), "reloc", dict(th=10.0, orb_dist=64, check_ori=1),
                  dict(max_dist=64.0, ordered=1, check_ori=1)),
        "sim3": (lambda: build(95, 8.0, "pred", False, 1.0, True), "sim3", dict(th=8, ratio_hamming=0.9),
                 dict(max_dist=float(np.float32(50) * np.float32(0.9)), ordered=1)),
        "fuse_stereo": (lambda: build(96, 3.0, "pred", True, 1.0, True, 40.0), "fuse", dict(th=3.0, bf=40.0),
                        dict(max_dist=50.0, ordered=0, chi2_gate=1)),
        "fuse_sim3": (lambda: build(97, 4.0, "pred", False, 1.0, True), "fuse_sim3", dict(th=4.0), dict(max_dist=50.0, ordered=0)),
    }


def make_projected_golden(ref):
    out = {}
    for name, (mk, kind, rargs, _) in projected_cases().items():
        frame, pts, kl = mk()
        out[name + "/in"] = np.frombuffer(bytes.fromhex(digest(frame.desc, frame.kp_xy, pts.desc, pts.uv, pts.radius, pts.active, kl)), dtype=np.uint8)
        if kind == "cur_last":
            n, own = ref.projected_cur_last(frame, pts, rargs["th"], rargs["mode"], rargs["mbf"], kl, rargs["check_ori"])
            out[name + "/kp_owner"] = own
        elif kind == "reloc":
            n, own = ref.projected_reloc(frame, pts, rargs["th"], rargs["orb_dist"], kl, rargs["check_ori"])
            out[name + "/kp_owner"] = own
        elif kind == "sim3":
            n, own = ref.projected_sim3(frame, pts, rargs["th"], rargs["ratio_hamming"], kl)
            out[name + "/kp_owner"] = own
        elif kind == "fuse":
            n, bi = ref.projected_fuse(frame, pts, rargs["th"], rargs["bf"])
            out[name + "/best_idx"] = bi
        else:
            n, bi = ref.projected_fuse_sim3(frame, pts, rargs["th"])
            out[name + "/best_idx"] = bi
        out[name + "/nmatches"] = np.int32(n)
        print("projected", name, "nmatches", n)
    np.savez_compressed(os.path.join(GOLD, "projected_outputs.npz"), **out)


def vocabulary_training_set(seed=5, n_images=200, n_per_image=500, n_centers=20000):
    rng = np.random.default_rng(seed)
    centers = synth.random_descriptors(rng, n_centers)
    return np.stack([centers[rng.integers(0, n_centers, n_per_image)] ^ synth.flip_mask(rng, n_per_image, rng.choice(np.array([3, 4]), size=n_per_image))
                     for _ in range(n_images)])


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = Reference()
    rng = np.random.default_rng(1234)

    # --- DescriptorDistance KATs (ORBmatcher.cc:2388-2408)
    a = synth.random_descriptors(rng, 256)
    b = synth.random_descriptors(rng, 256)
    a[0] = 0; b[0] = 255          # 256
    a[1] = b[1]                   # 0
    for i in range(2, 34):        # single-bit flips
        b[i] = a[i]
        b[i, (i - 2)] ^= np.uint8(1 << ((i - 2) % 8))
    d = np.array([ref.descriptor_distance(a[i], b[i]) for i in range(256)], dtype=np.int32)
    assert d[0] == 256 and d[1] == 0 and (d[2:34] == 1).all()
    np.savez_compressed(os.path.join(GOLD, "descriptor_distance.npz"), a=a, b=b, dist=d)

    # --- ComputeThreeMaxima (ORBmatcher.cc:2341-2383)
    hs = []
    for _ in range(64):
        h = rng.integers(0, 12, 30)
        if rng.random() < 0.5:
            h[rng.integers(0, 30, 3)] = rng.integers(0, 200, 3)
        if rng.random() < 0.3:
            h[rng.integers(0, 30)] = h.max()  # ties
        hs.append(h)
    hs.append(np.zeros(30, dtype=np.int64))
    hs = np.asarray(hs, dtype=np.int32)
    ind = np.stack([ref.compute_three_maxima(h) for h in hs])
    np.savez_compressed(os.path.join(GOLD, "three_maxima.npz"), histo=hs, ind=ind)

    # --- vocabulary built by the reference's create() (TemplatedVocabulary.h:560), k=10, L=4
    train = vocabulary_training_set()
    hv = ref.voc_create(train, 10, 4, seed=42)
    voc = hv.export()
    lev = voc.node_levels()
    leaves = np.diff(voc.child_offsets) == 0
    leaves[0] = False
    assert lev[leaves].min() == 4, "synthetic vocabulary must have no leaf above the nid level"
    voc.save(os.path.join(GOLD, "voc_k10_L4.npz"))
    print("vocabulary:", voc.n_nodes, "nodes", int(leaves.sum()), "words")

    # --- C1 / C2
    out = {}
    for name, mk in golden_cases().items():
        c = mk()
        if name.startswith("init"):
            n, m, prev = ref.search_for_initialization(c.f1, c.f2, c.prev_matched, c.window_size, c.nnratio, c.check_ori)
            out[name + "/in"] = np.frombuffer(bytes.fromhex(digest(c.f1.desc, c.f1.kp_xy, c.f2.desc, c.f2.kp_xy, c.f2.octave, c.f2.angle)), dtype=np.uint8)
            out[name + "/nmatches"] = np.int32(n)
            out[name + "/matches12"] = m
            out[name + "/prev"] = prev
            print(name, "nmatches", n)
        else:
            n, k = ref.search_by_projection_local(c.frame, c.mps, c.th, c.far_points, c.th_far, c.nnratio, c.kp_prior_obs, c.kp_mp)
            out[name + "/in"] = np.frombuffer(bytes.fromhex(digest(c.frame.desc, c.frame.kp_xy, c.mps.desc, c.mps.proj_xy, c.mps.n_obs)), dtype=np.uint8)
            out[name + "/nmatches"] = np.int32(n)
            out[name + "/kp_mp"] = k
            print(name, "nmatches", n)

    # --- C3: transform + SearchByBoW at levelsup 2 (bucketed) and 4 (root bucket, Frame.cc:1008)
    bc = synth.make_bow_case(31, voc, 2000)
    out["bow_s31/in"] = np.frombuffer(bytes.fromhex(digest(bc.kf.desc, bc.f.desc, bc.kf.angle, bc.f.angle, bc.kf_mp_valid)), dtype=np.uint8)
    for levelsup in (2, 4):
        tk = hv.transform(bc.kf.desc, levelsup)
        tf = hv.transform(bc.f.desc, levelsup)
        for side, t in (("kf", tk), ("f", tf)):
            for key, val in t.items():
                out[f"bow_s31/l{levelsup}/{side}/{key}"] = val
        kf = bc.kf.with_featvec(tk["fv_node_ids"], tk["fv_offsets"], tk["fv_features"])
        f = bc.f.with_featvec(tf["fv_node_ids"], tf["fv_offsets"], tf["fv_features"])
        n, m = ref.search_by_bow_kf_f(kf, f, bc.kf_mp_valid, 0.7, 1)
        out[f"bow_s31/l{levelsup}/kf_f/nmatches"] = np.int32(n)
        out[f"bow_s31/l{levelsup}/kf_f/match"] = m
        n2, m2 = ref.search_by_bow_kf_kf(kf, f, bc.kf_mp_valid, bc.f_mp_valid, 0.9, 1)
        out[f"bow_s31/l{levelsup}/kf_kf/nmatches"] = np.int32(n2)
        out[f"bow_s31/l{levelsup}/kf_kf/match"] = m2
        print("bow levelsup", levelsup, "kf_f", n, "kf_kf", n2)

    # --- C4: 8 pairs x 2000 features through the reference (poses in, its own geometry)
    tc = synth.make_triangulation_case(41, n_pairs=8, n_feat=2000)
    P = tc.kf1.shape[0]
    ep = np.zeros((P, 2), np.float32)
    f12 = np.zeros((P, 9), np.float32)
    for p in range(P):
        ep[p], f12[p] = ref.triangulation_geometry(tc.T1w[p], tc.T2w[p], tc.K, tc.K)
    out["tri_s41/in"] = np.frombuffer(bytes.fromhex(digest(tc.kfs.desc, tc.kfs.kp_xy, tc.kfs.node_id, tc.kfs.has_mp, tc.T1w, tc.T2w)), dtype=np.uint8)
    out["tri_s41/ep"] = ep
    out["tri_s41/f12"] = f12
    for co in (0, 1):
        nm, m = ref.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.T1w, tc.T2w, tc.K, 0, 0, co, 0.6, n_threads=4)
        out[f"tri_s41/ori{co}/nmatches"] = nm
        out[f"tri_s41/ori{co}/matches"] = m.astype(np.int16)
        print("tri checkOri", co, nm)

    # --- C5: 2-NN over the reference's DescriptorDistance
    kc = synth.make_knn_case(51, 512, 20000)
    bi, bd, sd, mt = ref.knn2_ratio(kc.q, kc.db, kc.th_low, kc.nnratio, 8)
    out["knn_s51/in"] = np.frombuffer(bytes.fromhex(digest(kc.q, kc.db)), dtype=np.uint8)
    out["knn_s51/best_idx"], out["knn_s51/best_dist"], out["knn_s51/second_dist"], out["knn_s51/match"] = bi, bd, sd, mt
    print("knn matches", int((mt >= 0).sum()))

    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)
    make_projected_golden(ref)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "projected":  # only the row-a6 fixtures
        make_projected_golden(Reference())
    else:
        main()
