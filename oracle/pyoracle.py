"""ORACLE python bindings (test infrastructure, NOT product code).

Loads oracle/liborb_oracle.so (plain-C restatement) and, when present,
oracle/_ref/libref_orbmatcher.so (the reference's own ORBmatcher.cc + DBoW2 compiled
unmodified).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from orb_slam3_comments_ghr_b200._abi import (BowDbHostStruct, HostBowDb, FrameHostStruct, HostFrame, HostKfSet, HostMapPoints, HostProjPoints, HostVoc,
                                              ProjPointsHostStruct, ProjSearchParamsStruct, proj_params,
                                              KfSetHostStruct, MapPointsHostStruct, VocHostStruct, as_f32, as_i32,
                                              as_u8, f32p, f64p, i32p, u8p, u32p)

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref: bool = True) -> None:
    """compile the C restatement (and oracle/_ref when /root/reference is mounted)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborb_oracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else C.cast(None, t)


class Oracle:
    """Plain-C restatement (orb_oracle.c)."""

    kind = "port"

    def __init__(self):
        path = os.path.join(_HERE, "liborb_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.oracle_comparisons.restype = C.c_int64
        L.oracle_descriptor_distance.argtypes = [u8p, u8p]
        L.oracle_grid_build.argtypes = [C.POINTER(FrameHostStruct), i32p, i32p]
        L.oracle_features_in_area.argtypes = [C.POINTER(FrameHostStruct), i32p, i32p, C.c_float, C.c_float, C.c_float,
                                              C.c_int, C.c_int, i32p]
        L.oracle_compute_three_maxima.argtypes = [i32p, C.c_int, i32p]
        L.oracle_search_for_initialization.argtypes = [C.POINTER(FrameHostStruct), C.POINTER(FrameHostStruct), f32p,
                                                       C.c_int, C.c_float, C.c_int, i32p]
        L.oracle_search_by_projection_local.argtypes = [C.POINTER(FrameHostStruct), C.POINTER(MapPointsHostStruct),
                                                        C.c_float, C.c_int, C.c_float, C.c_float, i32p, i32p]
        L.oracle_search_projected.argtypes = [C.POINTER(FrameHostStruct), C.POINTER(ProjPointsHostStruct),
                                              C.POINTER(ProjSearchParamsStruct), u8p, i32p, i32p, i32p]
        L.oracle_voc_transform.argtypes = [C.POINTER(VocHostStruct), C.c_int32, u8p, C.c_int, u32p, u32p, f64p]
        L.oracle_bowvector.argtypes = [C.c_int32, u32p, f64p, u32p, f64p]
        L.oracle_featvec.argtypes = [C.c_int32, u32p, f64p, u32p, i32p, u32p]
        L.oracle_search_by_bow_kf_f.argtypes = [C.POINTER(FrameHostStruct), C.POINTER(FrameHostStruct), u8p, C.c_float,
                                                C.c_int, i32p]
        L.oracle_search_by_bow_kf_kf.argtypes = [C.POINTER(FrameHostStruct), C.POINTER(FrameHostStruct), u8p, u8p,
                                                 C.c_float, C.c_int, i32p]
        L.oracle_search_for_triangulation.argtypes = [C.POINTER(KfSetHostStruct), C.c_int, C.c_int, f32p, f32p, C.c_int,
                                                      C.c_int, C.c_int, i32p]
        L.oracle_search_for_triangulation_batch.argtypes = [C.POINTER(KfSetHostStruct), C.c_int, i32p, i32p, f32p, f32p,
                                                            C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_int]
        L.oracle_knn2_ratio.argtypes = [C.c_int64, u8p, C.c_int64, u8p, C.c_int, C.c_float, i32p, i32p, i32p, i32p,
                                        C.c_int]

    # -- counters
    def comparisons(self) -> int:
        return int(self.lib.oracle_comparisons())

    def reset_comparisons(self):
        self.lib.oracle_comparisons_reset()

    # -- primitives
    def descriptor_distance(self, a, b) -> int:
        a, b = as_u8(a), as_u8(b)
        return int(self.lib.oracle_descriptor_distance(_p(a, u8p), _p(b, u8p)))

    def grid(self, f: HostFrame):
        cs = np.zeros(f.grid_cols * f.grid_rows + 1, dtype=np.int32)
        ci = np.full(max(f.n, 1), -1, dtype=np.int32)
        s = f.struct()
        self.lib.oracle_grid_build(C.byref(s), _p(cs, i32p), _p(ci, i32p))
        return cs, ci

    def features_in_area(self, f: HostFrame, x, y, r, min_level=-1, max_level=-1, grid=None):
        cs, ci = grid if grid is not None else self.grid(f)
        out = np.empty(max(f.n, 1), dtype=np.int32)
        s = f.struct()
        n = self.lib.oracle_features_in_area(C.byref(s), _p(cs, i32p), _p(ci, i32p), float(x), float(y), float(r),
                                             int(min_level), int(max_level), _p(out, i32p))
        return out[:n].copy()

    def compute_three_maxima(self, sizes):
        sizes = as_i32(sizes)
        ind = np.zeros(3, dtype=np.int32)
        self.lib.oracle_compute_three_maxima(_p(sizes, i32p), int(sizes.shape[0]), _p(ind, i32p))
        return ind

    # -- searches
    def search_for_initialization(self, f1: HostFrame, f2: HostFrame, prev_matched, window_size, nnratio, check_ori):
        prev = as_f32(prev_matched).copy()
        m = np.empty(f1.n, dtype=np.int32)
        s1, s2 = f1.struct(), f2.struct()
        n = self.lib.oracle_search_for_initialization(C.byref(s1), C.byref(s2), _p(prev, f32p), int(window_size),
                                                      float(nnratio), int(check_ori), _p(m, i32p))
        return int(n), m, prev

    def search_by_projection_local(self, f: HostFrame, mps: HostMapPoints, th, far_points, th_far, nnratio,
                                   kp_prior_obs, kp_mp):
        kp_mp = as_i32(kp_mp).copy()
        prior = as_i32(kp_prior_obs)
        sf, sm = f.struct(), mps.struct()
        n = self.lib.oracle_search_by_projection_local(C.byref(sf), C.byref(sm), float(th), int(far_points),
                                                       float(th_far), float(nnratio), _p(prior, i32p), _p(kp_mp, i32p))
        return int(n), kp_mp

    def stereo_coarse_match(self, left: HostFrame, right: HostFrame, n_rows, mb, mbf):
        nl = left.n
        bi = np.full(max(nl, 1), -1, dtype=np.int32)
        bd = np.full(max(nl, 1), 100, dtype=np.int32)
        self.lib.oracle_stereo_coarse_match.argtypes = [C.c_int32, u8p, f32p, i32p, C.c_int32, u8p, f32p, i32p, f32p, C.c_int32,
                                                        C.c_float, C.c_float, i32p, i32p]
        self.lib.oracle_stereo_coarse_match.restype = None
        self.lib.oracle_stereo_coarse_match(nl, _p(left.desc, u8p), _p(left.kp_xy, f32p), _p(left.octave, i32p), right.n,
                                            _p(right.desc, u8p), _p(right.kp_xy, f32p), _p(right.octave, i32p),
                                            _p(left.scale_factors, f32p), int(n_rows), float(mb), float(mbf), _p(bi, i32p), _p(bd, i32p))
        return bi[:nl], bd[:nl]

    # -- Frame::isInFrustum and its helpers
    def is_in_frustum(self, fr, world_pos, normal, min_distance, max_distance):
        """Frame.cc:676-782 -> dict of the MapPoint members the function writes"""
        wp, nm = as_f32(world_pos).reshape(-1, 3), as_f32(normal).reshape(-1, 3)
        mn, mx = as_f32(min_distance), as_f32(max_distance)
        n = wp.shape[0]
        out = _frustum_outputs(n)
        from orb_slam3_comments_ghr_b200._abi import FrustumHostStruct
        self.lib.oracle_is_in_frustum.argtypes = [C.POINTER(FrustumHostStruct), C.c_int32, f32p, f32p, f32p, f32p, u8p, f32p, f32p, f32p,
                                                  i32p, f32p]
        self.lib.oracle_is_in_frustum.restype = None
        self.lib.oracle_is_in_frustum(C.byref(fr), n, _p(wp, f32p), _p(nm, f32p), _p(mn, f32p), _p(mx, f32p), _p(out["in_view"], u8p),
                                      _p(out["proj_xy"], f32p), _p(out["proj_xr"], f32p), _p(out["depth"], f32p),
                                      _p(out["scale_level"], i32p), _p(out["view_cos"], f32p))
        return {k: v[:n] for k, v in out.items()}

    def predict_scale(self, max_distance, current_dist, log_scale_factor, n_levels):
        self.lib.oracle_predict_scale.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        return np.array([self.lib.oracle_predict_scale(float(a), float(b), float(log_scale_factor), int(n_levels))
                         for a, b in zip(max_distance, current_dist)], dtype=np.int32)

    def pinhole_project(self, K, xyz):
        K, xyz = as_f32(K), as_f32(xyz).reshape(-1, 3)
        uv = np.empty((xyz.shape[0], 2), dtype=np.float32)
        self.lib.oracle_pinhole_project.argtypes = [f32p, f32p, f32p]
        self.lib.oracle_pinhole_project.restype = None
        for i in range(xyz.shape[0]):
            self.lib.oracle_pinhole_project(_p(K, f32p), xyz[i].ctypes.data_as(f32p), uv[i].ctypes.data_as(f32p))
        return uv

    def epipolar_constrain(self, f12, kp1_xy, kp2_xy, unc):
        f12, a, b, u = as_f32(f12).reshape(9), as_f32(kp1_xy).reshape(-1, 2), as_f32(kp2_xy).reshape(-1, 2), as_f32(unc)
        self.lib.oracle_epipolar_constrain.argtypes = [f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        return np.array([self.lib.oracle_epipolar_constrain(_p(f12, f32p), float(a[i, 0]), float(a[i, 1]), float(b[i, 0]), float(b[i, 1]),
                                                            float(u[i])) for i in range(a.shape[0])], dtype=np.uint8)

    def compute_distinctive_descriptors(self, offsets, desc):
        off = as_i32(offsets)
        d = as_u8(desc).reshape(-1, 32)
        n = off.shape[0] - 1
        bi = np.full(max(n, 1), -1, dtype=np.int32)
        bm = np.full(max(n, 1), -1, dtype=np.int32)
        self.lib.oracle_compute_distinctive_descriptors.argtypes = [C.c_int32, i32p, u8p, i32p, i32p]
        self.lib.oracle_compute_distinctive_descriptors.restype = None
        self.lib.oracle_compute_distinctive_descriptors(n, _p(off, i32p), _p(d, u8p), _p(bi, i32p), _p(bm, i32p))
        return bi[:n], bm[:n]

    def bow_score_l1(self, db: HostBowDb, q_words, q_values):
        qw = np.ascontiguousarray(q_words, dtype=np.uint32)
        qv = np.ascontiguousarray(q_values, dtype=np.float64)
        common = np.zeros(max(db.n_kf, 1), dtype=np.int32)
        scores = np.zeros(max(db.n_kf, 1), dtype=np.float64)
        self.lib.oracle_bow_score_l1.argtypes = [C.POINTER(BowDbHostStruct), C.c_int32, u32p, f64p, i32p, f64p]
        self.lib.oracle_bow_score_l1.restype = None
        s = db.struct()
        self.lib.oracle_bow_score_l1(C.byref(s), qw.shape[0], _p(qw, u32p), _p(qv, f64p), _p(common, i32p), _p(scores, f64p))
        return common[:db.n_kf], scores[:db.n_kf]

    def search_projected(self, f: HostFrame, pts: HostProjPoints, max_dist, ordered, kp_locked=None, stereo_gate=False,
                         chi2_gate=False, check_ori=False, inv_level_sigma2=None):
        prm = proj_params(max_dist, ordered, stereo_gate, chi2_gate, check_ori, inv_level_sigma2)
        sf, sp = f.struct(), pts.struct()
        kl = as_u8(kp_locked) if kp_locked is not None else None
        bi = np.full(max(pts.n, 1), -1, dtype=np.int32)
        bd = np.full(max(pts.n, 1), 256, dtype=np.int32)
        own = np.full(max(f.n, 1), -1, dtype=np.int32)
        n = self.lib.oracle_search_projected(C.byref(sf), C.byref(sp), C.byref(prm),
                                             _p(kl, u8p) if kl is not None else C.cast(None, u8p), _p(bi, i32p), _p(bd, i32p),
                                             _p(own, i32p))
        return int(n), bi[:pts.n], bd[:pts.n], own[:f.n]

    def voc_transform(self, voc: HostVoc, desc, levelsup):
        desc = as_u8(desc).reshape(-1, 32)
        n = desc.shape[0]
        w = np.empty(n, dtype=np.uint32)
        nid = np.empty(n, dtype=np.uint32)
        wt = np.empty(n, dtype=np.float64)
        sv = voc.struct()
        self.lib.oracle_voc_transform(C.byref(sv), n, _p(desc, u8p), int(levelsup), _p(w, u32p), _p(nid, u32p),
                                      _p(wt, f64p))
        return w, nid, wt

    def bowvector(self, word_id, weight):
        n = word_id.shape[0]
        words = np.empty(max(n, 1), dtype=np.uint32)
        vals = np.empty(max(n, 1), dtype=np.float64)
        m = self.lib.oracle_bowvector(n, _p(word_id, u32p), _p(weight, f64p), _p(words, u32p), _p(vals, f64p))
        return words[:m].copy(), vals[:m].copy()

    def featvec(self, node_id, weight):
        n = node_id.shape[0]
        nodes = np.empty(max(n, 1), dtype=np.uint32)
        offs = np.empty(n + 1, dtype=np.int32)
        feats = np.empty(max(n, 1), dtype=np.uint32)
        m = self.lib.oracle_featvec(n, _p(node_id, u32p), _p(weight, f64p), _p(nodes, u32p), _p(offs, i32p),
                                    _p(feats, u32p))
        return nodes[:m].copy(), offs[:m + 1].copy(), feats[:offs[m]].copy()

    def search_by_bow_kf_f(self, kf: HostFrame, f: HostFrame, kf_mp_valid, nnratio, check_ori):
        valid = as_u8(kf_mp_valid)
        out = np.empty(f.n, dtype=np.int32)
        a, b = kf.struct(), f.struct()
        n = self.lib.oracle_search_by_bow_kf_f(C.byref(a), C.byref(b), _p(valid, u8p), float(nnratio), int(check_ori),
                                               _p(out, i32p))
        return int(n), out

    def search_by_bow_kf_kf(self, kf1: HostFrame, kf2: HostFrame, v1, v2, nnratio, check_ori):
        v1, v2 = as_u8(v1), as_u8(v2)
        out = np.empty(kf1.n, dtype=np.int32)
        a, b = kf1.struct(), kf2.struct()
        n = self.lib.oracle_search_by_bow_kf_kf(C.byref(a), C.byref(b), _p(v1, u8p), _p(v2, u8p), float(nnratio),
                                                int(check_ori), _p(out, i32p))
        return int(n), out

    def search_for_triangulation_batch(self, s: HostKfSet, kf1, kf2, ep, f12, only_stereo=0, coarse=0, check_ori=0,
                                       n_threads=1):
        kf1, kf2 = as_i32(kf1), as_i32(kf2)
        ep, f12 = as_f32(ep), as_f32(f12)
        P = kf1.shape[0]
        m = np.empty((P, s.n_feat), dtype=np.int32)
        nm = np.empty(P, dtype=np.int32)
        ss = s.struct()
        self.lib.oracle_search_for_triangulation_batch(C.byref(ss), P, _p(kf1, i32p), _p(kf2, i32p), _p(ep, f32p),
                                                       _p(f12, f32p), int(only_stereo), int(coarse), int(check_ori),
                                                       _p(m, i32p), _p(nm, i32p), int(n_threads))
        return nm, m

    def knn2_ratio(self, q, db, th_low=50, nnratio=0.8, n_threads=1):
        q, db = as_u8(q).reshape(-1, 32), as_u8(db).reshape(-1, 32)
        nq = q.shape[0]
        bi = np.empty(nq, dtype=np.int32)
        bd = np.empty(nq, dtype=np.int32)
        sd = np.empty(nq, dtype=np.int32)
        mt = np.empty(nq, dtype=np.int32)
        self.lib.oracle_knn2_ratio(nq, _p(q, u8p), db.shape[0], _p(db, u8p), int(th_low), float(nnratio), _p(bi, i32p),
                                   _p(bd, i32p), _p(sd, i32p), _p(mt, i32p), int(n_threads))
        return bi, bd, sd, mt


def _frustum_outputs(n):
    m = max(n, 1)
    return {"in_view": np.zeros(m, dtype=np.uint8), "proj_xy": np.zeros((m, 2), dtype=np.float32), "proj_xr": np.zeros(m, dtype=np.float32),
            "depth": np.zeros(m, dtype=np.float32), "scale_level": np.zeros(m, dtype=np.int32), "view_cos": np.zeros(m, dtype=np.float32)}


class Reference:
    """The reference's own ORBmatcher.cc / DBoW2 (oracle/_ref/libref_orbmatcher*.so)."""

    kind = "reference"

    @staticmethod
    def available(fast: bool = False) -> bool:
        name = "libref_orbmatcher_fast.so" if fast else "libref_orbmatcher.so"
        return os.path.exists(os.path.join(_HERE, "_ref", name))

    def __init__(self, fast: bool = False):
        name = "libref_orbmatcher_fast.so" if fast else "libref_orbmatcher.so"
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference is mounted")
        self.lib = L = C.CDLL(path)
        FH, MP, KS, VH = C.POINTER(FrameHostStruct), C.POINTER(MapPointsHostStruct), C.POINTER(KfSetHostStruct), \
            C.POINTER(VocHostStruct)
        L.ref_descriptor_distance.argtypes = [u8p, u8p]
        L.ref_compute_three_maxima.argtypes = [i32p, C.c_int, i32p]
        L.ref_features_in_area.argtypes = [FH, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, i32p]
        L.ref_grid.argtypes = [FH, i32p, i32p]
        L.ref_search_for_initialization.argtypes = [FH, FH, f32p, C.c_int, C.c_float, C.c_int, i32p]
        L.ref_search_by_projection_local.argtypes = [FH, MP, C.c_float, C.c_int, C.c_float, C.c_float, i32p, i32p]
        PP = C.POINTER(ProjPointsHostStruct)
        if not hasattr(L, "ref_projected_cur_last"):
            raise RuntimeError(f"{path} predates the row-a6 harnesses: rebuild with `make -C oracle ref`")
        L.ref_projected_cur_last.argtypes = [FH, PP, C.c_float, C.c_int, C.c_float, u8p, C.c_int, i32p]
        L.ref_projected_reloc.argtypes = [FH, PP, C.c_float, C.c_int, u8p, C.c_int, i32p]
        L.ref_projected_sim3.argtypes = [FH, PP, C.c_int, C.c_float, u8p, i32p]
        L.ref_projected_fuse.argtypes = [FH, PP, C.c_float, C.c_float, i32p]
        L.ref_projected_fuse_sim3.argtypes = [FH, PP, C.c_float, i32p]
        L.ref_search_by_bow_kf_f.argtypes = [FH, FH, u8p, C.c_float, C.c_int, i32p]
        L.ref_search_by_bow_kf_kf.argtypes = [FH, FH, u8p, u8p, C.c_float, C.c_int, i32p]
        L.ref_triangulation_geometry.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p]
        L.ref_search_for_triangulation.argtypes = [KS, C.c_int, C.c_int, f32p, f32p, f32p, f32p, C.c_int, C.c_int,
                                                   C.c_int, C.c_float, i32p]
        L.ref_search_for_triangulation_batch.argtypes = [KS, C.c_int, i32p, i32p, f32p, f32p, f32p, C.c_int, C.c_int,
                                                         C.c_int, C.c_float, i32p, i32p, C.c_int]
        L.ref_knn2_ratio.argtypes = [C.c_int64, u8p, C.c_int64, u8p, C.c_int, C.c_float, i32p, i32p, i32p, i32p, C.c_int]
        L.ref_voc_create.restype = C.c_void_p
        L.ref_voc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.ref_voc_from_flat.restype = C.c_void_p
        L.ref_voc_from_flat.argtypes = [VH]
        L.ref_voc_destroy.argtypes = [C.c_void_p]
        L.ref_voc_n_nodes.argtypes = [C.c_void_p]
        L.ref_voc_n_children.argtypes = [C.c_void_p]
        L.ref_voc_n_words.argtypes = [C.c_void_p]
        L.ref_voc_export.argtypes = [C.c_void_p, u8p, i32p, u32p, f64p, u32p]
        L.ref_voc_transform.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, u32p, u32p, f64p, u32p, f64p, i32p, u32p,
                                        i32p, u32p]

    def descriptor_distance(self, a, b) -> int:
        a, b = as_u8(a), as_u8(b)
        return int(self.lib.ref_descriptor_distance(_p(a, u8p), _p(b, u8p)))

    # ---- the helpers whose bodies are the reference's own text cut out by oracle/extract_ref.py
    def keyframe_features_in_area(self, f: HostFrame, x, y, r):
        """KeyFrame::GetFeaturesInArea (KeyFrame.cc:859-907) + IsInImage (:910-913) -> (indices, in_image)"""
        s = f.struct()
        out = np.empty(max(f.n, 1), dtype=np.int32)
        inimg = C.c_int32(0)
        self.lib.ref_keyframe_features_in_area.argtypes = [C.POINTER(FrameHostStruct), C.c_float, C.c_float, C.c_float, i32p, C.POINTER(C.c_int32)]
        n = self.lib.ref_keyframe_features_in_area(C.byref(s), float(x), float(y), float(r), _p(out, i32p), C.byref(inimg))
        return out[:n].copy(), bool(inimg.value)

    def predict_scale(self, max_distance, min_distance, current_dist, log_scale_factor, n_levels):
        """MapPoint::PredictScale x2 + Get{Min,Max}DistanceInvariance (MapPoint.cc:665-738) -> (level_kf, level_f, min_inv, max_inv)"""
        mx, mn, cd = as_f32(max_distance), as_f32(min_distance), as_f32(current_dist)
        n = mx.shape[0]
        lk, lf = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        a, b = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
        self.lib.ref_predict_scale.argtypes = [C.c_int, f32p, f32p, f32p, C.c_float, C.c_int, i32p, i32p, f32p, f32p]
        self.lib.ref_predict_scale.restype = None
        self.lib.ref_predict_scale(n, _p(mx, f32p), _p(mn, f32p), _p(cd, f32p), float(log_scale_factor), int(n_levels), _p(lk, i32p),
                                   _p(lf, i32p), _p(a, f32p), _p(b, f32p))
        return lk, lf, a, b

    def pinhole_project(self, K, xyz):
        K, xyz = as_f32(K), as_f32(xyz).reshape(-1, 3)
        uv = np.empty((xyz.shape[0], 2), dtype=np.float32)
        self.lib.ref_pinhole_project.argtypes = [f32p, C.c_int, f32p, f32p]
        self.lib.ref_pinhole_project.restype = None
        self.lib.ref_pinhole_project(_p(K, f32p), xyz.shape[0], _p(xyz, f32p), _p(uv, f32p))
        return uv

    def epipolar_constrain(self, K1, K2, R12, t12, kp1_xy, kp2_xy, unc):
        """Pinhole::epipolarConstrain (Pinhole.cpp:189-219) -> (ok[n], the F12 it forms)"""
        a, b, u = as_f32(kp1_xy).reshape(-1, 2), as_f32(kp2_xy).reshape(-1, 2), as_f32(unc)
        ok = np.zeros(a.shape[0], dtype=np.uint8)
        f12 = np.zeros(9, dtype=np.float32)
        self.lib.ref_epipolar_constrain.argtypes = [f32p, f32p, f32p, f32p, C.c_int, f32p, f32p, f32p, u8p, f32p]
        self.lib.ref_epipolar_constrain.restype = None
        self.lib.ref_epipolar_constrain(_p(as_f32(K1), f32p), _p(as_f32(K2), f32p), _p(as_f32(R12).reshape(9), f32p), _p(as_f32(t12), f32p),
                                        a.shape[0], _p(a, f32p), _p(b, f32p), _p(u, f32p), _p(ok, u8p), _p(f12, f32p))
        return ok, f12

    def is_in_frustum(self, f: HostFrame, Tcw34, K, mbf, viewing_cos_limit, world_pos, normal, min_distance, max_distance):
        """Frame::isInFrustum (Frame.cc:676-782) on a stand-in Frame with pose Tcw -> (ret, dict of the MapPoint members written)"""
        wp, nm = as_f32(world_pos).reshape(-1, 3), as_f32(normal).reshape(-1, 3)
        mn, mx = as_f32(min_distance), as_f32(max_distance)
        n = wp.shape[0]
        out = _frustum_outputs(n)
        ret = np.zeros(max(n, 1), dtype=np.uint8)
        s = f.struct()
        self.lib.ref_is_in_frustum.argtypes = [C.POINTER(FrameHostStruct), f32p, f32p, C.c_float, C.c_float, C.c_int, f32p, f32p, f32p, f32p,
                                               u8p, f32p, f32p, f32p, i32p, f32p, u8p]
        self.lib.ref_is_in_frustum.restype = None
        self.lib.ref_is_in_frustum(C.byref(s), _p(as_f32(Tcw34).reshape(12), f32p), _p(as_f32(K), f32p), float(mbf), float(viewing_cos_limit), n,
                                   _p(wp, f32p), _p(nm, f32p), _p(mn, f32p), _p(mx, f32p), _p(out["in_view"], u8p), _p(out["proj_xy"], f32p),
                                   _p(out["proj_xr"], f32p), _p(out["depth"], f32p), _p(out["scale_level"], i32p), _p(out["view_cos"], f32p),
                                   _p(ret, u8p))
        return ret[:n], {k: v[:n] for k, v in out.items()}

    def stereo_coarse_match(self, left: HostFrame, right: HostFrame, n_rows, mb, mbf):
        """coarse stage of Frame::ComputeStereoMatches (Frame.cc:1117-1247), the reference's own text"""
        nl = left.n
        bi = np.full(max(nl, 1), -1, dtype=np.int32)
        bd = np.full(max(nl, 1), 100, dtype=np.int32)
        self.lib.ref_stereo_coarse_match.argtypes = [C.c_int32, u8p, f32p, i32p, C.c_int32, u8p, f32p, i32p, f32p, C.c_int32, C.c_int32,
                                                     C.c_float, C.c_float, i32p, i32p]
        self.lib.ref_stereo_coarse_match.restype = None
        self.lib.ref_stereo_coarse_match(nl, _p(left.desc, u8p), _p(left.kp_xy, f32p), _p(left.octave, i32p), right.n, _p(right.desc, u8p),
                                         _p(right.kp_xy, f32p), _p(right.octave, i32p), _p(left.scale_factors, f32p),
                                         int(left.scale_factors.shape[0]), int(n_rows), float(mb), float(mbf), _p(bi, i32p), _p(bd, i32p))
        return bi[:nl], bd[:nl]

    def compute_three_maxima(self, sizes):
        sizes = as_i32(sizes)
        ind = np.zeros(3, dtype=np.int32)
        self.lib.ref_compute_three_maxima(_p(sizes, i32p), int(sizes.shape[0]), _p(ind, i32p))
        return ind

    def grid(self, f: HostFrame):
        cs = np.zeros(f.grid_cols * f.grid_rows + 1, dtype=np.int32)
        ci = np.full(max(f.n, 1), -1, dtype=np.int32)
        s = f.struct()
        self.lib.ref_grid(C.byref(s), _p(cs, i32p), _p(ci, i32p))
        return cs, ci

    def features_in_area(self, f: HostFrame, x, y, r, min_level=-1, max_level=-1):
        out = np.empty(max(f.n, 1), dtype=np.int32)
        s = f.struct()
        n = self.lib.ref_features_in_area(C.byref(s), float(x), float(y), float(r), int(min_level), int(max_level),
                                          _p(out, i32p))
        return out[:n].copy()

    def search_for_initialization(self, f1, f2, prev_matched, window_size, nnratio, check_ori):
        prev = as_f32(prev_matched).copy()
        m = np.empty(f1.n, dtype=np.int32)
        s1, s2 = f1.struct(), f2.struct()
        n = self.lib.ref_search_for_initialization(C.byref(s1), C.byref(s2), _p(prev, f32p), int(window_size),
                                                   float(nnratio), int(check_ori), _p(m, i32p))
        return int(n), m, prev

    def search_by_projection_local(self, f, mps, th, far_points, th_far, nnratio, kp_prior_obs, kp_mp):
        kp_mp = as_i32(kp_mp).copy()
        prior = as_i32(kp_prior_obs)
        sf, sm = f.struct(), mps.struct()
        n = self.lib.ref_search_by_projection_local(C.byref(sf), C.byref(sm), float(th), int(far_points), float(th_far),
                                                    float(nnratio), _p(prior, i32p), _p(kp_mp, i32p))
        return int(n), kp_mp

    def bow_score_l1(self, db: HostBowDb, q_words, q_values):
        qw = np.ascontiguousarray(q_words, dtype=np.uint32)
        qv = np.ascontiguousarray(q_values, dtype=np.float64)
        common = np.zeros(max(db.n_kf, 1), dtype=np.int32)
        scores = np.zeros(max(db.n_kf, 1), dtype=np.float64)
        self.lib.ref_bow_score_l1.argtypes = [C.POINTER(BowDbHostStruct), C.c_int32, u32p, f64p, i32p, f64p]
        self.lib.ref_bow_score_l1.restype = None
        s = db.struct()
        self.lib.ref_bow_score_l1(C.byref(s), qw.shape[0], _p(qw, u32p), _p(qv, f64p), _p(common, i32p), _p(scores, f64p))
        return common[:db.n_kf], scores[:db.n_kf]

    @staticmethod
    def mappoint_available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libref_mappoint.so"))

    def compute_distinctive_descriptors(self, offsets, desc, kf_bad=None):
        """the reference's own MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc with the real include/MapPoint.h,
        oracle/_ref/libref_mappoint.so): per map point the position of the kept descriptor (-1: none) and the descriptor."""
        path = os.path.join(_HERE, "_ref", "libref_mappoint.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference is mounted")
        L = C.CDLL(path)
        L.ref_compute_distinctive.argtypes = [C.c_int32, i32p, u8p, u8p, i32p, u8p]
        off, d = as_i32(offsets), as_u8(desc)
        n = off.shape[0] - 1
        bad = as_u8(kf_bad) if kf_bad is not None else None
        bi = np.full(max(n, 1), -1, dtype=np.int32)
        out = np.zeros((max(n, 1), 32), dtype=np.uint8)
        L.ref_compute_distinctive(n, _p(off, i32p), _p(d, u8p), _p(bad, u8p), _p(bi, i32p), _p(out, u8p))
        return bi[:n], out[:n]

    # the reference's self-projecting overloads on degenerate geometry (see ref_adapter.cc): everything from the window on
    def projected_cur_last(self, cur: HostFrame, pts: HostProjPoints, th, mode, mbf, kp_locked, check_ori):
        own = np.full(max(cur.n, 1), -1, dtype=np.int32)
        a, b, kl = cur.struct(), pts.struct(), as_u8(kp_locked)
        n = self.lib.ref_projected_cur_last(C.byref(a), C.byref(b), float(th), int(mode), float(mbf), _p(kl, u8p), int(check_ori),
                                            _p(own, i32p))
        return int(n), own[:cur.n]

    def projected_reloc(self, cur: HostFrame, pts: HostProjPoints, th, orb_dist, kp_locked, check_ori):
        own = np.full(max(cur.n, 1), -1, dtype=np.int32)
        a, b, kl = cur.struct(), pts.struct(), as_u8(kp_locked)
        n = self.lib.ref_projected_reloc(C.byref(a), C.byref(b), float(th), int(orb_dist), _p(kl, u8p), int(check_ori), _p(own, i32p))
        return int(n), own[:cur.n]

    def projected_sim3(self, kf: HostFrame, pts: HostProjPoints, th, ratio_hamming, kp_locked):
        own = np.full(max(kf.n, 1), -1, dtype=np.int32)
        a, b, kl = kf.struct(), pts.struct(), as_u8(kp_locked)
        n = self.lib.ref_projected_sim3(C.byref(a), C.byref(b), int(th), float(ratio_hamming), _p(kl, u8p), _p(own, i32p))
        return int(n), own[:kf.n]

    def projected_fuse(self, kf: HostFrame, pts: HostProjPoints, th, bf):
        bi = np.full(max(pts.n, 1), -1, dtype=np.int32)
        a, b = kf.struct(), pts.struct()
        n = self.lib.ref_projected_fuse(C.byref(a), C.byref(b), float(th), float(bf), _p(bi, i32p))
        return int(n), bi[:pts.n]

    def projected_fuse_sim3(self, kf: HostFrame, pts: HostProjPoints, th):
        bi = np.full(max(pts.n, 1), -1, dtype=np.int32)
        a, b = kf.struct(), pts.struct()
        n = self.lib.ref_projected_fuse_sim3(C.byref(a), C.byref(b), float(th), _p(bi, i32p))
        return int(n), bi[:pts.n]

    def search_by_bow_kf_f(self, kf, f, kf_mp_valid, nnratio, check_ori):
        valid = as_u8(kf_mp_valid)
        out = np.empty(f.n, dtype=np.int32)
        a, b = kf.struct(), f.struct()
        n = self.lib.ref_search_by_bow_kf_f(C.byref(a), C.byref(b), _p(valid, u8p), float(nnratio), int(check_ori),
                                            _p(out, i32p))
        return int(n), out

    def search_by_bow_kf_kf(self, kf1, kf2, v1, v2, nnratio, check_ori):
        v1, v2 = as_u8(v1), as_u8(v2)
        out = np.empty(kf1.n, dtype=np.int32)
        a, b = kf1.struct(), kf2.struct()
        n = self.lib.ref_search_by_bow_kf_kf(C.byref(a), C.byref(b), _p(v1, u8p), _p(v2, u8p), float(nnratio),
                                             int(check_ori), _p(out, i32p))
        return int(n), out

    def triangulation_geometry(self, T1w, T2w, K1, K2):
        """ep[2], f12[9] from poses ([R row-major | t], 12 floats) the way ORBmatcher.cc:1053-1071 +
        Pinhole.cpp:194-197 compute them in this build."""
        T1w, T2w, K1, K2 = as_f32(T1w), as_f32(T2w), as_f32(K1), as_f32(K2)
        ep = np.empty(2, dtype=np.float32)
        f12 = np.empty(9, dtype=np.float32)
        self.lib.ref_triangulation_geometry(_p(T1w, f32p), _p(T2w, f32p), _p(K1, f32p), _p(K2, f32p), _p(ep, f32p),
                                            _p(f12, f32p))
        return ep, f12

    def search_for_triangulation_batch(self, s: HostKfSet, kf1, kf2, T1w, T2w, K, only_stereo=0, coarse=0, check_ori=0,
                                       nnratio=0.6, n_threads=1):
        kf1, kf2 = as_i32(kf1), as_i32(kf2)
        T1w, T2w, K = as_f32(T1w), as_f32(T2w), as_f32(K)
        P = kf1.shape[0]
        m = np.empty((P, s.n_feat), dtype=np.int32)
        nm = np.empty(P, dtype=np.int32)
        ss = s.struct()
        self.lib.ref_search_for_triangulation_batch(C.byref(ss), P, _p(kf1, i32p), _p(kf2, i32p), _p(T1w, f32p),
                                                    _p(T2w, f32p), _p(K, f32p), int(only_stereo), int(coarse),
                                                    int(check_ori), float(nnratio), _p(m, i32p), _p(nm, i32p),
                                                    int(n_threads))
        return nm, m

    def knn2_ratio(self, q, db, th_low=50, nnratio=0.8, n_threads=1):
        q, db = as_u8(q).reshape(-1, 32), as_u8(db).reshape(-1, 32)
        nq = q.shape[0]
        bi = np.empty(nq, dtype=np.int32)
        bd = np.empty(nq, dtype=np.int32)
        sd = np.empty(nq, dtype=np.int32)
        mt = np.empty(nq, dtype=np.int32)
        self.lib.ref_knn2_ratio(nq, _p(q, u8p), db.shape[0], _p(db, u8p), int(th_low), float(nnratio), _p(bi, i32p),
                                _p(bd, i32p), _p(sd, i32p), _p(mt, i32p), int(n_threads))
        return bi, bd, sd, mt

    # -- vocabulary
    def voc_create(self, training_desc, k, L, seed) -> "RefVocHandle":
        """training_desc [n_images, n_per_image, 32] -> TemplatedVocabulary::create (TemplatedVocabulary.h:560)"""
        d = as_u8(training_desc)
        h = self.lib.ref_voc_create(int(k), int(L), int(d.shape[0]), int(d.shape[1]), _p(d, u8p), int(seed))
        return RefVocHandle(self, h, k, L)

    def voc_from_flat(self, voc: HostVoc) -> "RefVocHandle":
        s = voc.struct()
        return RefVocHandle(self, self.lib.ref_voc_from_flat(C.byref(s)), voc.k, voc.L)


class RefVocHandle:
    def __init__(self, ref: Reference, h, k, L):
        self.ref, self.h, self.k, self.L = ref, h, k, L

    def __del__(self):
        try:
            self.ref.lib.ref_voc_destroy(self.h)
        except Exception:
            pass

    def export(self) -> HostVoc:
        L = self.ref.lib
        n = L.ref_voc_n_nodes(self.h)
        nc = L.ref_voc_n_children(self.h)
        desc = np.zeros((n, 32), dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.int32)
        ch = np.zeros(max(nc, 1), dtype=np.uint32)
        w = np.zeros(n, dtype=np.float64)
        wid = np.zeros(n, dtype=np.uint32)
        L.ref_voc_export(self.h, _p(desc, u8p), _p(off, i32p), _p(ch, u32p), _p(w, f64p), _p(wid, u32p))
        return HostVoc(self.k, self.L, desc, off, ch[:nc], w, wid)

    def transform(self, desc, levelsup):
        desc = as_u8(desc).reshape(-1, 32)
        n = desc.shape[0]
        w = np.empty(n, dtype=np.uint32)
        nid = np.empty(n, dtype=np.uint32)
        wt = np.empty(n, dtype=np.float64)
        bw = np.empty(max(n, 1), dtype=np.uint32)
        bv = np.empty(max(n, 1), dtype=np.float64)
        nn = C.c_int32(0)
        fn = np.empty(max(n, 1), dtype=np.uint32)
        fo = np.empty(n + 1, dtype=np.int32)
        ff = np.empty(max(n, 1), dtype=np.uint32)
        nw = self.ref.lib.ref_voc_transform(self.h, n, _p(desc, u8p), int(levelsup), _p(w, u32p), _p(nid, u32p),
                                            _p(wt, f64p), _p(bw, u32p), _p(bv, f64p), C.byref(nn), _p(fn, u32p),
                                            _p(fo, i32p), _p(ff, u32p))
        m = nn.value
        return dict(word_id=w, node_id=nid, weight=wt, bow_words=bw[:nw].copy(), bow_values=bv[:nw].copy(),
                    fv_node_ids=fn[:m].copy(), fv_offsets=fo[:m + 1].copy(), fv_features=ff[:fo[m]].copy())
