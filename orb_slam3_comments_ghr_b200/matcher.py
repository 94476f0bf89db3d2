"""Host-side mirror of the reference's matcher interface over the CUDA C-ABI library.

`ORBmatcher(nnratio, checkOri)` keeps the reference's constructor and method names
(include/ORBmatcher.h:37-84); methods take the flat host views of `_abi` (the members the
reference reads) or device handles, and return the reference's outputs in index form
(MapPoint* -> index, NULL -> -1).  Everything runs on the GPU through liborbmatch_b200.so:
if the library or a B200 is missing this module raises -- there is NO CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from ._abi import (FrustumHostStruct, HostLocalPoints, LocalPointsHostStruct, BowDbHostStruct, HostBowDb, FrameHostStruct, HostFrame, HostKfSet, HostMapPoints, HostProjPoints, HostVoc, KfSetHostStruct, MapPointsHostStruct,
                   ProjPointsHostStruct, ProjSearchParamsStruct, proj_params, VocHostStruct, as_f32, as_i32, as_u8, f32p, f64p, i32p, i64p, u8p, u32p)

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liborbmatch_b200.so")
_lib = None

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # ORBmatcher.cc:34-36
TC_DEFAULT = True  # engine 0 (auto) resolves to the tcgen05 engine for nq >= 128 and nd >= 1024


class OrbGpuError(RuntimeError):
    pass


class TriGatherStruct(C.Structure):
    """orbgpu_tri_gather (include/orbmatch_b200.h)"""
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("pairs", C.c_void_p * 8), ("counts", C.c_void_p * 8),
                ("flags", C.c_void_p * 8), ("epoch_done", C.c_void_p), ("status", C.c_void_p), ("wait_mask", C.c_uint32)]


def tri_gather_struct(rank: int, pairs, counts, flags, epoch_done: int, status: int = 0, wait_mask: Optional[int] = None) -> TriGatherStruct:
    n = len(pairs)
    g = TriGatherStruct()
    g.n_ranks, g.rank = n, int(rank)
    for r in range(n):
        g.pairs[r], g.counts[r], g.flags[r] = int(pairs[r]), int(counts[r]), int(flags[r])
    g.epoch_done = int(epoch_done)
    g.status = int(status) if status else None
    g.wait_mask = ((1 << n) - 1) if wait_mask is None else int(wait_mask)
    return g


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """Loads the CUDA extension; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise OrbGpuError(f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or make -C orb_slam3_comments_ghr_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    FH, MP, KS, VH = C.POINTER(FrameHostStruct), C.POINTER(MapPointsHostStruct), C.POINTER(KfSetHostStruct), C.POINTER(VocHostStruct)
    L.orbgpu_last_error.restype = C.c_char_p
    L.orbgpu_version.restype = C.c_char_p
    L.orbgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.orbgpu_create_on_stream.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.orbgpu_destroy.argtypes = [vp]
    L.orbgpu_destroy.restype = None
    L.orbgpu_synchronize.argtypes = [vp]
    L.orbgpu_stream.argtypes = [vp]
    L.orbgpu_stream.restype = vp
    L.orbgpu_launch_count.argtypes = [vp]
    L.orbgpu_launch_count.restype = C.c_int64
    L.orbgpu_last_comparisons.argtypes = [vp]
    L.orbgpu_last_comparisons.restype = C.c_int64
    L.orbgpu_descriptor_distance.argtypes = [vp, C.c_int64, u8p, u8p, i32p]
    L.orbgpu_frame_upload.argtypes = [vp, FH, C.POINTER(vp)]
    L.orbgpu_frame_destroy.argtypes = [vp]
    L.orbgpu_frame_destroy.restype = None
    L.orbgpu_frame_n.argtypes = [vp]
    L.orbgpu_frame_grid_download.argtypes = [vp, vp, i32p, i32p]
    L.orbgpu_features_in_area.argtypes = [vp, vp, C.c_int32, f32p, f32p, f32p, i32p, i32p, i32p, i32p, C.c_int64, i64p]
    L.orbgpu_search_for_initialization.argtypes = [vp, vp, vp, f32p, C.c_int32, C.c_float, C.c_int32, i32p, i32p]
    L.orbgpu_search_by_projection_local.argtypes = [vp, vp, MP, C.c_float, C.c_int32, C.c_float, C.c_float, i32p, i32p, i32p]
    L.orbgpu_voc_upload.argtypes = [vp, VH, C.POINTER(vp)]
    L.orbgpu_voc_destroy.argtypes = [vp]
    L.orbgpu_voc_destroy.restype = None
    L.orbgpu_transform.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, u32p, u32p, f64p]
    L.orbgpu_bowvector_download.argtypes = [vp, vp, i32p, u32p, f64p]
    L.orbgpu_featvec_download.argtypes = [vp, vp, i32p, u32p, i32p, u32p]
    L.orbgpu_search_by_bow_kf_f.argtypes = [vp, vp, vp, u8p, C.c_float, C.c_int32, i32p, i32p]
    L.orbgpu_search_by_bow_kf_kf.argtypes = [vp, vp, vp, u8p, u8p, C.c_float, C.c_int32, i32p, i32p]
    L.orbgpu_kfset_upload.argtypes = [vp, KS, C.POINTER(vp)]
    L.orbgpu_kfset_destroy.argtypes = [vp]
    L.orbgpu_kfset_destroy.restype = None
    L.orbgpu_search_for_triangulation_batch.argtypes = [vp, vp, C.c_int32, i32p, i32p, f32p, f32p, C.c_int32, C.c_int32, C.c_int32,
                                                        i32p, i32p]
    L.orbgpu_search_for_triangulation_batch_dev.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]
    L.orbgpu_db_upload.argtypes = [vp, C.c_int64, u8p, C.POINTER(vp)]
    L.orbgpu_db_from_dev.argtypes = [vp, C.c_int64, vp, C.POINTER(vp)]
    L.orbgpu_db_destroy.argtypes = [vp]
    L.orbgpu_db_destroy.restype = None
    L.orbgpu_knn2_ratio.argtypes = [vp, vp, C.c_int64, u8p, C.c_int32, C.c_float, i32p, i32p, i32p, i32p]
    L.orbgpu_knn2_ratio_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_int32, C.c_float, vp, vp, vp, vp]
    L.orbgpu_knn2_set_engine.argtypes = [vp, C.c_int32]
    L.orbgpu_search_projected.argtypes = [vp, vp, C.POINTER(ProjPointsHostStruct), C.POINTER(ProjSearchParamsStruct), u8p, i32p, i32p,
                                          i32p, i32p]
    L.orbgpu_triangulation_set_engine.argtypes = [vp, C.c_int32]
    L.orbgpu_compute_three_maxima.argtypes = [vp, i32p, C.c_int32, i32p]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise OrbGpuError(f"orbgpu error {rc}: {load_library().orbgpu_last_error().decode()}")


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else C.cast(None, t)


class Context:
    """One CUDA stream + workspace (orbgpu_ctx).  Use one per calling thread."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        L = load_library()
        self._h = C.c_void_p()
        if stream is None:
            _check(L.orbgpu_create(int(device), C.byref(self._h)))
        else:
            _check(L.orbgpu_create_on_stream(int(device), C.c_void_p(stream), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load_library().orbgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        _check(load_library().orbgpu_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(load_library().orbgpu_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(load_library().orbgpu_launch_count(self._h))

    @property
    def last_comparisons(self) -> int:
        return int(load_library().orbgpu_last_comparisons(self._h))

    def fetch_comparisons(self) -> int:
        out = C.c_int64(0)
        L = load_library()
        L.orbgpu_fetch_comparisons.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        _check(L.orbgpu_fetch_comparisons(self._h, C.byref(out)))
        return int(out.value)

    def measure_popc_peak(self, variant: int = 0):
        """integer-pipe peak of this GPU (register-only microbenchmark on all SMs) -> dict(popc_per_s, per_clk_sm, sm_mhz)"""
        L = load_library()
        L.orbgpu_measure_popc_peak.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        _check(L.orbgpu_measure_popc_peak(self._h, int(variant), C.byref(a), C.byref(b), C.byref(c)))
        return {"popc_per_s": a.value, "per_clk_sm": b.value, "sm_mhz": c.value}

    def set_knn_engine(self, engine: int):
        _check(load_library().orbgpu_knn2_set_engine(self._h, int(engine)))

    def set_triangulation_engine(self, engine: int):
        _check(load_library().orbgpu_triangulation_set_engine(self._h, int(engine)))

    # ---- a1
    def descriptor_distance(self, a, b) -> np.ndarray:
        a, b = as_u8(a).reshape(-1, 32), as_u8(b).reshape(-1, 32)
        out = np.empty(a.shape[0], dtype=np.int32)
        _check(load_library().orbgpu_descriptor_distance(self._h, a.shape[0], _p(a, u8p), _p(b, u8p), _p(out, i32p)))
        return out

    def compute_three_maxima(self, sizes) -> np.ndarray:
        sizes = as_i32(sizes)
        ind = np.zeros(3, dtype=np.int32)
        _check(load_library().orbgpu_compute_three_maxima(self._h, _p(sizes, i32p), int(sizes.shape[0]), _p(ind, i32p)))
        return ind

    # ---- 8(f) rank 1: Frame::isInFrustum (Frame.cc:676-782) for a list of map points -> dict of the MapPoint members it writes
    def is_in_frustum(self, fr: FrustumHostStruct, world_pos, normal, min_distance, max_distance):
        wp, nm = as_f32(world_pos).reshape(-1, 3), as_f32(normal).reshape(-1, 3)
        mn, mx = as_f32(min_distance), as_f32(max_distance)
        n = wp.shape[0]
        m = max(n, 1)
        out = {"in_view": np.zeros(m, dtype=np.uint8), "proj_xy": np.zeros((m, 2), dtype=np.float32), "proj_xr": np.zeros(m, dtype=np.float32),
               "depth": np.zeros(m, dtype=np.float32), "scale_level": np.zeros(m, dtype=np.int32), "view_cos": np.zeros(m, dtype=np.float32)}
        L = load_library()
        L.orbgpu_is_in_frustum.argtypes = [C.c_void_p, C.POINTER(FrustumHostStruct), C.c_int32, f32p, f32p, f32p, f32p, u8p, f32p, f32p, f32p,
                                           i32p, f32p]
        _check(L.orbgpu_is_in_frustum(self._h, C.byref(fr), n, _p(wp, f32p), _p(nm, f32p), _p(mn, f32p), _p(mx, f32p), _p(out["in_view"], u8p),
                                      _p(out["proj_xy"], f32p), _p(out["proj_xr"], f32p), _p(out["depth"], f32p),
                                      _p(out["scale_level"], i32p), _p(out["view_cos"], f32p)))
        return {k: v[:n] for k, v in out.items()}

    # ---- uploads
    def upload_frame(self, f: HostFrame) -> "DeviceFrame":
        return DeviceFrame(self, f)

    def upload_vocabulary(self, v: HostVoc) -> "DeviceVoc":
        return DeviceVoc(self, v)

    def upload_kfset(self, s: HostKfSet) -> "DeviceKfSet":
        return DeviceKfSet(self, s)

    # coarse stage of Frame::ComputeStereoMatches (Frame.cc:1139-1216) -> (best right keypoint or -1, bestDist) per left keypoint
    def stereo_coarse_match(self, left: HostFrame, right: HostFrame, n_rows: int, mb: float, mbf: float):
        nl, nr = left.n, right.n
        bi = np.full(max(nl, 1), -1, dtype=np.int32)
        bd = np.full(max(nl, 1), TH_HIGH, dtype=np.int32)
        L = load_library()
        L.orbgpu_stereo_coarse_match.argtypes = [C.c_void_p, C.c_int32, u8p, f32p, i32p, C.c_int32, u8p, f32p, i32p, f32p, C.c_int32,
                                                 C.c_int32, C.c_float, C.c_float, i32p, i32p]
        _check(L.orbgpu_stereo_coarse_match(self._h, nl, _p(left.desc, u8p), _p(left.kp_xy, f32p), _p(left.octave, i32p), nr,
                                            _p(right.desc, u8p), _p(right.kp_xy, f32p), _p(right.octave, i32p),
                                            _p(left.scale_factors, f32p), left.scale_factors.shape[0], int(n_rows), float(mb), float(mbf),
                                            _p(bi, i32p), _p(bd, i32p)))
        return bi[:nl], bd[:nl]

    # MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:444-535) for a batch of map points -> (best_idx, best_median)
    def compute_distinctive_descriptors(self, offsets, desc):
        off = as_i32(offsets)
        d = as_u8(desc).reshape(-1, 32)
        n = off.shape[0] - 1
        bi = np.full(max(n, 1), -1, dtype=np.int32)
        bm = np.full(max(n, 1), -1, dtype=np.int32)
        L = load_library()
        L.orbgpu_compute_distinctive_descriptors.argtypes = [C.c_void_p, C.c_int32, i32p, u8p, i32p, i32p]
        _check(L.orbgpu_compute_distinctive_descriptors(self._h, n, _p(off, i32p), _p(d, u8p), _p(bi, i32p), _p(bm, i32p)))
        return bi[:n], bm[:n]

    def upload_bow_database(self, db: HostBowDb) -> "DeviceBowDb":
        return DeviceBowDb(self, db)

    def upload_database(self, db) -> "DeviceDb":
        return DeviceDb(self, host=db)

    def database_from_device(self, ptr: int, nd: int, keepalive=None) -> "DeviceDb":
        return DeviceDb(self, dev_ptr=ptr, nd=nd, keepalive=keepalive)


class DeviceFrame:
    """Device copy of a Frame/KeyFrame (descriptors as uint4 pairs + CSR cell grid)."""

    def __init__(self, ctx: Context, f: HostFrame):
        self.ctx, self.host = ctx, f
        self._h = C.c_void_p()
        s = f.struct()
        _check(load_library().orbgpu_frame_upload(ctx.handle, C.byref(s), C.byref(self._h)))
        self.n = f.n

    def __del__(self):
        try:
            if self._h:
                load_library().orbgpu_frame_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def grid(self):
        cs = np.zeros(self.host.grid_cols * self.host.grid_rows + 1, dtype=np.int32)
        ci = np.full(max(self.n, 1), -1, dtype=np.int32)
        _check(load_library().orbgpu_frame_grid_download(self.ctx.handle, self._h, _p(cs, i32p), _p(ci, i32p)))
        return cs, ci

    def features_in_area(self, x, y, r, min_level, max_level):
        """batched Frame::GetFeaturesInArea -> (offsets[nq+1], idx[total]) in reference order"""
        x, y, r = as_f32(np.atleast_1d(x)), as_f32(np.atleast_1d(y)), as_f32(np.atleast_1d(r))
        nq = x.shape[0]
        mn = as_i32(np.broadcast_to(np.asarray(min_level), (nq,)))
        mx = as_i32(np.broadcast_to(np.asarray(max_level), (nq,)))
        cap = max(1, nq * max(self.n, 1))
        off = np.zeros(nq + 1, dtype=np.int32)
        idx = np.empty(cap, dtype=np.int32)
        total = C.c_int64(0)
        _check(load_library().orbgpu_features_in_area(self.ctx.handle, self._h, nq, _p(x, f32p), _p(y, f32p), _p(r, f32p), _p(mn, i32p),
                                                      _p(mx, i32p), _p(off, i32p), _p(idx, i32p), cap, C.byref(total)))
        return off, idx[:total.value].copy()

    def transform(self, voc: "DeviceVoc", levelsup: int = 4, store_featvec: bool = True):
        """TemplatedVocabulary::transform on the device -> (word_id, node_id, weight) per feature"""
        n = self.n
        w = np.empty(n, dtype=np.uint32)
        nid = np.empty(n, dtype=np.uint32)
        wt = np.empty(n, dtype=np.float64)
        _check(load_library().orbgpu_transform(self.ctx.handle, voc.handle, self._h, int(levelsup), int(store_featvec), _p(w, u32p),
                                               _p(nid, u32p), _p(wt, f64p)))
        return w, nid, wt

    def bowvector(self):
        n = max(self.n, 1)
        words = np.empty(n, dtype=np.uint32)
        vals = np.empty(n, dtype=np.float64)
        m = C.c_int32(0)
        _check(load_library().orbgpu_bowvector_download(self.ctx.handle, self._h, C.byref(m), _p(words, u32p), _p(vals, f64p)))
        return words[:m.value].copy(), vals[:m.value].copy()

    def featvec(self):
        n = max(self.n, 1)
        nodes = np.empty(n, dtype=np.uint32)
        offs = np.empty(n + 1, dtype=np.int32)
        feats = np.empty(n, dtype=np.uint32)
        m = C.c_int32(0)
        _check(load_library().orbgpu_featvec_download(self.ctx.handle, self._h, C.byref(m), _p(nodes, u32p), _p(offs, i32p),
                                                      _p(feats, u32p)))
        k = m.value
        return nodes[:k].copy(), offs[:k + 1].copy(), feats[:offs[k] if k > 0 else 0].copy()


class DeviceVoc:
    def __init__(self, ctx: Context, v: HostVoc):
        self.ctx, self.host = ctx, v
        self._h = C.c_void_p()
        s = v.struct()
        _check(load_library().orbgpu_voc_upload(ctx.handle, C.byref(s), C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                load_library().orbgpu_voc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


class DeviceKfSet:
    def __init__(self, ctx: Context, s: HostKfSet):
        self.ctx, self.host = ctx, s
        self._h = C.c_void_p()
        st = s.struct()
        _check(load_library().orbgpu_kfset_upload(ctx.handle, C.byref(st), C.byref(self._h)))
        self.n_kf, self.n_feat = s.n_kf, s.n_feat

    def transform(self, voc: "DeviceVoc", levelsup: int):
        """KeyFrame::ComputeBoW for the whole set on the device (FeatureVector node ids + CSR rebuild); returns the number of
        descriptor comparisons of the vocabulary descent."""
        L = load_library()
        L.orbgpu_kfset_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _check(L.orbgpu_kfset_transform(self.ctx.handle, voc.handle, self._h, int(levelsup)))
        return self.ctx.last_comparisons

    def __del__(self):
        try:
            if self._h:
                load_library().orbgpu_kfset_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


class DeviceBowDb:
    """Device CSR of the key frames' BowVectors; score() = KeyFrameDatabase's common-word count and L1Scoring::score of a query
    BowVector against every key frame (KeyFrameDatabase.cc:928-943, :970; ScoringObject.cpp:23-68)."""

    def __init__(self, ctx: Context, db: HostBowDb):
        self.ctx, self.n_kf = ctx, db.n_kf
        self._h = C.c_void_p()
        L = load_library()
        L.orbgpu_bowdb_upload.argtypes = [C.c_void_p, C.POINTER(BowDbHostStruct), C.POINTER(C.c_void_p)]
        L.orbgpu_bowdb_destroy.argtypes = [C.c_void_p]
        L.orbgpu_bowdb_destroy.restype = None
        L.orbgpu_bow_score_l1.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, u32p, f64p, i32p, f64p]
        s = db.struct()
        _check(L.orbgpu_bowdb_upload(ctx.handle, C.byref(s), C.byref(self._h)))

    def score(self, q_words, q_values):
        qw = np.ascontiguousarray(q_words, dtype=np.uint32)
        qv = np.ascontiguousarray(q_values, dtype=np.float64)
        common = np.zeros(max(self.n_kf, 1), dtype=np.int32)
        scores = np.zeros(max(self.n_kf, 1), dtype=np.float64)
        _check(load_library().orbgpu_bow_score_l1(self.ctx.handle, self._h, qw.shape[0], _p(qw, u32p), _p(qv, f64p), _p(common, i32p),
                                                  _p(scores, f64p)))
        return common[:self.n_kf], scores[:self.n_kf]

    def __del__(self):
        try:
            if self._h:
                load_library().orbgpu_bowdb_destroy(self._h)
                self._h = None
        except Exception:
            pass


class DeviceDb:
    def __init__(self, ctx: Context, host=None, dev_ptr: int = 0, nd: int = 0, keepalive=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        self._keep = keepalive
        if host is not None:
            d = as_u8(host).reshape(-1, 32)
            self.nd = d.shape[0]
            _check(load_library().orbgpu_db_upload(ctx.handle, self.nd, _p(d, u8p), C.byref(self._h)))
        else:
            self.nd = int(nd)
            _check(load_library().orbgpu_db_from_dev(ctx.handle, self.nd, C.c_void_p(dev_ptr), C.byref(self._h)))

    def invalidate(self):
        """the borrowed device memory changed: drop the cached +-1 fp8 expansion (rebuilt by the next tensor-engine search)"""
        L = load_library()
        L.orbgpu_db_invalidate.argtypes = [C.c_void_p]
        _check(L.orbgpu_db_invalidate(self._h))

    def update(self, host):
        """re-upload the descriptors in place (no reallocation)"""
        d = as_u8(host).reshape(-1, 32)
        L = load_library()
        L.orbgpu_db_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, u8p]
        _check(L.orbgpu_db_update(self.ctx.handle, self._h, d.shape[0], _p(d, u8p)))
        self.nd = d.shape[0]

    def __del__(self):
        try:
            if self._h:
                load_library().orbgpu_db_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


class ORBmatcher:
    """Drop-in mirror of ORB_SLAM3::ORBmatcher (include/ORBmatcher.h:34-99) on the GPU."""

    TH_HIGH, TH_LOW, HISTO_LENGTH = TH_HIGH, TH_LOW, HISTO_LENGTH

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, ctx: Optional[Context] = None):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.ctx = ctx if ctx is not None else Context(0)

    # ORBmatcher.h:40
    def DescriptorDistance(self, a, b):
        return self.ctx.descriptor_distance(a, b)

    # ORBmatcher.h:69 -> (nmatches, vnMatches12, vbPrevMatched)
    def SearchForInitialization(self, F1: DeviceFrame, F2: DeviceFrame, vbPrevMatched, windowSize: int = 10):
        prev = as_f32(vbPrevMatched).copy().reshape(-1, 2)
        m = np.empty(F1.n, dtype=np.int32)
        nm = C.c_int32(0)
        _check(load_library().orbgpu_search_for_initialization(self.ctx.handle, F1.handle, F2.handle, _p(prev, f32p), int(windowSize),
                                                               self.mfNNratio, int(self.mbCheckOrientation), _p(m, i32p), C.byref(nm)))
        return nm.value, m, prev

    # ORBmatcher.h:44 -> (nmatches, F.mvpMapPoints as indices into vpMapPoints)
    def SearchByProjection(self, F: DeviceFrame, vpMapPoints: HostMapPoints, th: float = 3.0, bFarPoints: bool = False,
                           thFarPoints: float = 50.0, kp_prior_obs=None, kp_mp=None):
        n = F.n
        prior = as_i32(kp_prior_obs) if kp_prior_obs is not None else np.zeros(n, dtype=np.int32)
        out = as_i32(kp_mp).copy() if kp_mp is not None else np.full(n, -1, dtype=np.int32)
        nm = C.c_int32(0)
        s = vpMapPoints.struct()
        _check(load_library().orbgpu_search_by_projection_local(self.ctx.handle, F.handle, C.byref(s), float(th), int(bFarPoints),
                                                                float(thFarPoints), self.mfNNratio, _p(prior, i32p), _p(out, i32p),
                                                                C.byref(nm)))
        return nm.value, out

    # Tracking::SearchLocalPoints: Frame::isInFrustum of every local map point + SearchByProjection (ORBmatcher.h:44), fused on the
    # device -> (nmatches, F.mvpMapPoints as indices into the local points, mbTrackInView per point)
    def SearchLocalPoints(self, F: DeviceFrame, fr: FrustumHostStruct, pts: HostLocalPoints, th: float = 1.0, bFarPoints: bool = False,
                          thFarPoints: float = 50.0, kp_prior_obs=None, kp_mp=None):
        n = F.n
        prior = as_i32(kp_prior_obs) if kp_prior_obs is not None else np.zeros(n, dtype=np.int32)
        out = as_i32(kp_mp).copy() if kp_mp is not None else np.full(n, -1, dtype=np.int32)
        inv = np.zeros(max(pts.n, 1), dtype=np.uint8)
        nm = C.c_int32(0)
        s = pts.struct()
        L = load_library()
        L.orbgpu_search_local_points.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(FrustumHostStruct), C.POINTER(LocalPointsHostStruct),
                                                 C.c_float, C.c_int32, C.c_float, C.c_float, i32p, i32p, u8p, C.POINTER(C.c_int32)]
        _check(L.orbgpu_search_local_points(self.ctx.handle, F.handle, C.byref(fr), C.byref(s), float(th), int(bFarPoints),
                                            float(thFarPoints), self.mfNNratio, _p(prior, i32p), _p(out, i32p), _p(inv, u8p), C.byref(nm)))
        return nm.value, out, inv[:pts.n]

    # search core of the self-projecting overloads (ORBmatcher.h:48-60, 78-84): points already projected by the caller
    # -> (nmatches, best_idx per point, best_dist per point, owner point per keypoint)
    def SearchProjected(self, F: DeviceFrame, pts: HostProjPoints, max_dist: float, ordered: bool, kp_locked=None,
                        stereo_gate: bool = False, chi2_gate: bool = False, inv_level_sigma2=None):
        prm = proj_params(max_dist, ordered, stereo_gate, chi2_gate, self.mbCheckOrientation and pts.angle is not None, inv_level_sigma2)
        s = pts.struct()
        kl = as_u8(kp_locked) if kp_locked is not None else None
        bi = np.full(max(pts.n, 1), -1, dtype=np.int32)
        bd = np.full(max(pts.n, 1), 256, dtype=np.int32)
        own = np.full(max(F.n, 1), -1, dtype=np.int32)
        nm = C.c_int32(0)
        _check(load_library().orbgpu_search_projected(self.ctx.handle, F.handle, C.byref(s), C.byref(prm),
                                                      _p(kl, u8p) if kl is not None else C.cast(None, u8p), _p(bi, i32p), _p(bd, i32p),
                                                      _p(own, i32p), C.byref(nm)))
        return nm.value, bi[:pts.n], bd[:pts.n], own[:F.n]

    # ORBmatcher.h:65 -> (nmatches, vpMapPointMatches as KF feature indices, indexed by F feature)
    def SearchByBoW(self, KF: DeviceFrame, F: DeviceFrame, kf_mp_valid, f_mp_valid=None):
        v1 = as_u8(kf_mp_valid)
        nm = C.c_int32(0)
        if f_mp_valid is None:  # KeyFrame <-> Frame
            out = np.empty(F.n, dtype=np.int32)
            _check(load_library().orbgpu_search_by_bow_kf_f(self.ctx.handle, KF.handle, F.handle, _p(v1, u8p), self.mfNNratio,
                                                            int(self.mbCheckOrientation), _p(out, i32p), C.byref(nm)))
        else:  # KeyFrame <-> KeyFrame (ORBmatcher.h:66)
            v2 = as_u8(f_mp_valid)
            out = np.empty(KF.n, dtype=np.int32)
            _check(load_library().orbgpu_search_by_bow_kf_kf(self.ctx.handle, KF.handle, F.handle, _p(v1, u8p), _p(v2, u8p),
                                                             self.mfNNratio, int(self.mbCheckOrientation), _p(out, i32p), C.byref(nm)))
        return nm.value, out

    # one frame against K candidate key frames (Tracking::Relocalization's loop over SearchByBoW, Tracking.cc:4469-4495) in one call
    # -> (nmatches[K], vpMapPointMatches[K, F.N] as KF feature indices)
    def SearchByBoWBatch(self, KFs, F: DeviceFrame, kf_mp_valids):
        K = len(KFs)
        vs = [as_u8(v) for v in kf_mp_valids]
        out = np.full((max(K, 1), max(F.n, 1)), -1, dtype=np.int32)
        nm = np.zeros(max(K, 1), dtype=np.int32)
        hs = (C.c_void_p * max(K, 1))(*[k.handle for k in KFs])
        ps = (u8p * max(K, 1))(*[_p(v, u8p) for v in vs])
        L = load_library()
        L.orbgpu_search_by_bow_kf_f_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(u8p), C.c_float,
                                                      C.c_int32, i32p, i32p]
        _check(L.orbgpu_search_by_bow_kf_f_batch(self.ctx.handle, K, hs, F.handle, ps, self.mfNNratio, int(self.mbCheckOrientation),
                                                 _p(out, i32p), _p(nm, i32p)))
        return nm[:K], out[:K, :F.n]

    # the current key frame against the K key frames of a candidate's covisibility window (LoopClosing.cc:909-925) in one call
    # -> (nmatches[K], vpMatches12[K, KF1.N] as KF2 feature indices)
    def SearchByBoWBatchKF(self, KF1: DeviceFrame, kf1_mp_valid, KF2s, kf2_mp_valids):
        K = len(KF2s)
        v1 = as_u8(kf1_mp_valid)
        vs = [as_u8(v) for v in kf2_mp_valids]
        out = np.full((max(K, 1), max(KF1.n, 1)), -1, dtype=np.int32)
        nm = np.zeros(max(K, 1), dtype=np.int32)
        hs = (C.c_void_p * max(K, 1))(*[k.handle for k in KF2s])
        ps = (u8p * max(K, 1))(*[_p(v, u8p) for v in vs])
        L = load_library()
        L.orbgpu_search_by_bow_kf_kf_batch.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(u8p),
                                                       C.c_float, C.c_int32, i32p, i32p]
        _check(L.orbgpu_search_by_bow_kf_kf_batch(self.ctx.handle, KF1.handle, _p(v1, u8p), K, hs, ps, self.mfNNratio,
                                                  int(self.mbCheckOrientation), _p(out, i32p), _p(nm, i32p)))
        return nm[:K], out[:K, :KF1.n]

    # ORBmatcher.h:72, batched over pairs -> (nmatches[P], vMatches12[P, n_feat])
    def SearchForTriangulation(self, kfs: DeviceKfSet, kf1, kf2, ep, f12, bOnlyStereo: bool = False, bCoarse: bool = False, out=None):
        """out: optional caller-owned (vMatches12[P, n_feat] int32, nmatches[P] int32) host buffers -- pinned memory makes the
        device-to-host copies asynchronous DMA transfers."""
        kf1, kf2 = as_i32(kf1), as_i32(kf2)
        ep, f12 = as_f32(ep), as_f32(f12)
        P = kf1.shape[0]
        if out is not None:
            m, nm = out
            assert m.dtype == np.int32 and m.shape == (P, kfs.n_feat) and m.flags["C_CONTIGUOUS"]
            assert nm.dtype == np.int32 and nm.shape == (P,)
            _check(load_library().orbgpu_search_for_triangulation_batch(self.ctx.handle, kfs.handle, P, _p(kf1, i32p), _p(kf2, i32p),
                                                                        _p(ep, f32p), _p(f12, f32p), int(bOnlyStereo), int(bCoarse),
                                                                        int(self.mbCheckOrientation), _p(m, i32p), _p(nm, i32p)))
            return nm, m
        m = np.empty((P, kfs.n_feat), dtype=np.int32)
        nm = np.empty(P, dtype=np.int32)
        _check(load_library().orbgpu_search_for_triangulation_batch(self.ctx.handle, kfs.handle, P, _p(kf1, i32p), _p(kf2, i32p),
                                                                    _p(ep, f32p), _p(f12, f32p), int(bOnlyStereo), int(bCoarse),
                                                                    int(self.mbCheckOrientation), _p(m, i32p), _p(nm, i32p)))
        return nm, m

    # same search, vMatchedPairs form: (pair_offsets[P+1], pairs[total, 2]) -- only the matched pairs cross the bus
    def SearchForTriangulationPairs(self, kfs: DeviceKfSet, kf1, kf2, ep, f12, bOnlyStereo: bool = False, bCoarse: bool = False,
                                    out=None):
        kf1, kf2 = as_i32(kf1), as_i32(kf2)
        ep, f12 = as_f32(ep), as_f32(f12)
        P = kf1.shape[0]
        if out is None:
            out = (np.empty(P + 1, dtype=np.int32), np.empty((P * kfs.n_feat, 2), dtype=np.int32))
        offs, pairs = out
        total = C.c_int64(0)
        L = load_library()
        L.orbgpu_search_for_triangulation_batch_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, i32p, i32p, f32p, f32p, C.c_int32,
                                                                  C.c_int32, C.c_int32, i32p, i32p, C.c_int64, C.POINTER(C.c_int64)]
        _check(L.orbgpu_search_for_triangulation_batch_pairs(self.ctx.handle, kfs.handle, P, _p(kf1, i32p), _p(kf2, i32p), _p(ep, f32p),
                                                             _p(f12, f32p), int(bOnlyStereo), int(bCoarse), int(self.mbCheckOrientation),
                                                             _p(offs, i32p), _p(pairs, i32p), pairs.shape[0], C.byref(total)))
        return offs, pairs[:total.value]

    def SearchForTriangulation_dev(self, kfs: DeviceKfSet, n_pairs, kf1_ptr, kf2_ptr, ep_ptr, f12_ptr, matches_ptr, nmatches_ptr,
                                   bOnlyStereo=False, bCoarse=False):
        vp = C.c_void_p
        _check(load_library().orbgpu_search_for_triangulation_batch_dev(self.ctx.handle, kfs.handle, int(n_pairs), vp(kf1_ptr), vp(kf2_ptr),
                                                                        vp(ep_ptr), vp(f12_ptr), int(bOnlyStereo), int(bCoarse),
                                                                        int(self.mbCheckOrientation), vp(matches_ptr), vp(nmatches_ptr)))

    # fused search + all-gather: rows / counts of this rank's pairs stored into every rank's result buffers (peer memory)
    def SearchForTriangulation_peers_dev(self, kfs: DeviceKfSet, n_pairs, kf1_ptr, kf2_ptr, ep_ptr, f12_ptr, target_matches, target_nmatches,
                                         pair_offset: int, rows_preset=True, bCoarse=False):
        vp = C.c_void_p
        n = len(target_matches)
        tm = (C.c_void_p * n)(*[int(x) for x in target_matches])
        tn = (C.c_void_p * n)(*[int(x) for x in target_nmatches])
        L = load_library()
        L.orbgpu_search_for_triangulation_batch_peers_dev.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32,
                                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64, C.c_int32]
        _check(L.orbgpu_search_for_triangulation_batch_peers_dev(self.ctx.handle, kfs.handle, int(n_pairs), vp(kf1_ptr), vp(kf2_ptr),
                                                                 vp(ep_ptr), vp(f12_ptr), int(bCoarse), int(self.mbCheckOrientation), n, tm, tn,
                                                                 int(pair_offset), int(rows_preset)))

    # fused search + all-gather in the vMatchedPairs form: compact (idx1 << 16 | idx2) entries + counts into every rank's buffers,
    # epoch flags instead of a barrier (orbgpu_search_for_triangulation_batch_gather_dev)
    def SearchForTriangulation_gather_dev(self, kfs: DeviceKfSet, n_pairs, kf1_ptr, kf2_ptr, ep_ptr, f12_ptr, gather: TriGatherStruct,
                                          pair_offset: int, bCoarse=False):
        vp = C.c_void_p
        L = load_library()
        L.orbgpu_search_for_triangulation_batch_gather_dev.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32,
                                                                       C.POINTER(TriGatherStruct), C.c_int64]
        _check(L.orbgpu_search_for_triangulation_batch_gather_dev(self.ctx.handle, kfs.handle, int(n_pairs), vp(kf1_ptr), vp(kf2_ptr),
                                                                  vp(ep_ptr), vp(f12_ptr), int(bCoarse), int(self.mbCheckOrientation),
                                                                  C.byref(gather), int(pair_offset)))

    # the gathered compact result on the host as (pair_offsets[P+1], pairs[total, 2]) -- offsets scan + packing on the device
    def TriangulationGatherDownload(self, n_pairs: int, n_feat: int, counts_ptr: int, entries_ptr: int, out=None):
        if out is None:
            out = (np.empty(n_pairs + 1, dtype=np.int32), np.empty((n_pairs * 512, 2), dtype=np.int32))
        offs, pairs = out
        total = C.c_int64(0)
        L = load_library()
        L.orbgpu_tri_gather_download.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, i32p, i32p, C.c_int64,
                                                 C.POINTER(C.c_int64)]
        _check(L.orbgpu_tri_gather_download(self.ctx.handle, int(n_pairs), int(n_feat), C.c_void_p(counts_ptr), C.c_void_p(entries_ptr),
                                            _p(offs, i32p), _p(pairs, i32p), pairs.shape[0], C.byref(total)))
        return offs, pairs[:total.value]

    # brute-force 2-NN + ratio test ("SearchByNN" of BASELINE.json) -> best_idx, best_dist, second_dist, match
    def SearchByNN(self, db: DeviceDb, q, th_low: int = TH_LOW, database=None):
        """database: new host descriptors for `db` (an owned database of at least that many rows) -- uploaded in chunks on the database's
        own stream and searched chunk by chunk as they arrive (orbgpu_knn2_ratio_update); None: search the resident database"""
        q = as_u8(q).reshape(-1, 32)
        nq = q.shape[0]
        bi = np.empty(nq, dtype=np.int32)
        bd = np.empty(nq, dtype=np.int32)
        sd = np.empty(nq, dtype=np.int32)
        mt = np.empty(nq, dtype=np.int32)
        L = load_library()
        if database is not None:
            d = as_u8(database).reshape(-1, 32)
            L.orbgpu_knn2_ratio_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, u8p, C.c_int64, u8p, C.c_int32, C.c_float, i32p, i32p,
                                                   i32p, i32p]
            _check(L.orbgpu_knn2_ratio_update(self.ctx.handle, db.handle, d.shape[0], _p(d, u8p), nq, _p(q, u8p), int(th_low), self.mfNNratio,
                                              _p(bi, i32p), _p(bd, i32p), _p(sd, i32p), _p(mt, i32p)))
            db.nd = d.shape[0]
            return bi, bd, sd, mt
        _check(L.orbgpu_knn2_ratio(self.ctx.handle, db.handle, nq, _p(q, u8p), int(th_low), self.mfNNratio, _p(bi, i32p),
                                   _p(bd, i32p), _p(sd, i32p), _p(mt, i32p)))
        return bi, bd, sd, mt

    def SearchByNN_dev(self, db: DeviceDb, nq: int, q_ptr: int, best_idx_ptr: int, best_dist_ptr: int, second_ptr: int, match_ptr: int,
                       th_low: int = TH_LOW):
        vp = C.c_void_p
        _check(load_library().orbgpu_knn2_ratio_dev(self.ctx.handle, db.handle, int(nq), vp(q_ptr), int(th_low), self.mfNNratio,
                                                    vp(best_idx_ptr), vp(best_dist_ptr), vp(second_ptr), vp(match_ptr)))
