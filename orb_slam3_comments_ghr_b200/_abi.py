"""ctypes mirror of include/orbmatch_b200.h (POD structs only).

The structs are the flat views of the reference members the matcher reads
(Frame.h:219-296, KeyFrame.h:378-406, MapPoint.h:166-239, TemplatedVocabulary.h:361-435).
`HostFrame` & co. keep the numpy arrays alive for as long as the struct is used.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    assert a.flags["C_CONTIGUOUS"], "arrays crossing the ABI must be C-contiguous"
    return a.ctypes.data_as(typ)


def as_u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class FrameHostStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("desc", u8p),
        ("kp_xy", f32p),
        ("octave", i32p),
        ("angle", f32p),
        ("u_right", f32p),
        ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
        ("grid_inv_w", C.c_float), ("grid_inv_h", C.c_float),
        ("grid_cols", C.c_int32), ("grid_rows", C.c_int32),
        ("n_levels", C.c_int32),
        ("scale_factors", f32p),
        ("level_sigma2", f32p),
        ("fv_n_nodes", C.c_int32),
        ("fv_node_ids", u32p),
        ("fv_offsets", i32p),
        ("fv_features", u32p),
    ]


class VocHostStruct(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("L", C.c_int32), ("n_nodes", C.c_int32),
        ("node_desc", u8p),
        ("child_offsets", i32p),
        ("child_ids", u32p),
        ("weight", f64p),
        ("word_id", u32p),
    ]


class MapPointsHostStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("desc", u8p),
        ("proj_xy", f32p),
        ("proj_xr", f32p),
        ("scale_level", i32p),
        ("view_cos", f32p),
        ("depth", f32p),
        ("in_view", u8p),
        ("bad", u8p),
        ("n_obs", i32p),
    ]


class ProjPointsHostStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("desc", u8p),
        ("uv", f32p),
        ("radius", f32p),
        ("min_level", i32p),
        ("max_level", i32p),
        ("ur", f32p),
        ("active", u8p),
        ("locks", u8p),
        ("angle", f32p),
    ]


class FrustumHostStruct(C.Structure):
    """orbgpu_frustum_host: the Frame members Frame::isInFrustum reads (Frame.cc:676-782)"""
    _fields_ = [
        ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("K", C.c_float * 4), ("mbf", C.c_float),
        ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
        ("viewing_cos_limit", C.c_float), ("log_scale_factor", C.c_float), ("n_levels", C.c_int32),
    ]


def frustum_struct(Rcw, tcw, Ow, K, mbf, bounds, viewing_cos_limit, log_scale_factor, n_levels) -> FrustumHostStruct:
    f = FrustumHostStruct()
    f.Rcw[:] = [float(x) for x in np.asarray(Rcw, dtype=np.float32).reshape(9)]
    f.tcw[:] = [float(x) for x in np.asarray(tcw, dtype=np.float32).reshape(3)]
    f.Ow[:] = [float(x) for x in np.asarray(Ow, dtype=np.float32).reshape(3)]
    f.K[:] = [float(x) for x in np.asarray(K, dtype=np.float32).reshape(4)]
    f.mbf = float(mbf)
    f.min_x, f.min_y, f.max_x, f.max_y = [float(x) for x in bounds]
    f.viewing_cos_limit, f.log_scale_factor, f.n_levels = float(viewing_cos_limit), float(log_scale_factor), int(n_levels)
    return f


class LocalPointsHostStruct(C.Structure):
    """orbgpu_localpoints_host: the map points of Tracking::SearchLocalPoints"""
    _fields_ = [("n", C.c_int32), ("desc", u8p), ("world_pos", f32p), ("normal", f32p), ("min_distance", f32p), ("max_distance", f32p),
                ("skip", u8p), ("bad", u8p), ("n_obs", i32p)]


@dataclass
class HostLocalPoints:
    desc: np.ndarray
    world_pos: np.ndarray
    normal: np.ndarray
    min_distance: np.ndarray
    max_distance: np.ndarray
    bad: np.ndarray
    n_obs: np.ndarray
    skip: Optional[np.ndarray] = None

    @property
    def n(self) -> int:
        return int(self.desc.shape[0])

    def struct(self) -> LocalPointsHostStruct:
        self.desc, self.bad = as_u8(self.desc), as_u8(self.bad)
        self.world_pos, self.normal = as_f32(self.world_pos), as_f32(self.normal)
        self.min_distance, self.max_distance = as_f32(self.min_distance), as_f32(self.max_distance)
        self.n_obs = as_i32(self.n_obs)
        if self.skip is not None:
            self.skip = as_u8(self.skip)
        return LocalPointsHostStruct(self.n, _ptr(self.desc, u8p), _ptr(self.world_pos, f32p), _ptr(self.normal, f32p),
                                     _ptr(self.min_distance, f32p), _ptr(self.max_distance, f32p), _ptr(self.skip, u8p),
                                     _ptr(self.bad, u8p), _ptr(self.n_obs, i32p))


class ProjSearchParamsStruct(C.Structure):
    _fields_ = [
        ("max_dist", C.c_float),
        ("ordered", C.c_int32),
        ("stereo_gate", C.c_int32),
        ("chi2_gate", C.c_int32),
        ("check_ori", C.c_int32),
        ("inv_level_sigma2", f32p),
    ]


class BowDbHostStruct(C.Structure):
    _fields_ = [("n_kf", C.c_int32), ("offsets", i32p), ("words", u32p), ("values", f64p)]


class KfSetHostStruct(C.Structure):
    _fields_ = [
        ("n_kf", C.c_int32), ("n_feat", C.c_int32),
        ("desc", u8p),
        ("kp_xy", f32p),
        ("octave", i32p),
        ("angle", f32p),
        ("has_mp", u8p),
        ("u_right", f32p),
        ("node_id", u32p),
        ("n_levels", C.c_int32),
        ("scale_factors", f32p),
        ("level_sigma2", f32p),
    ]


def orb_scale_tables(n_levels: int = 8, scale_factor: float = 1.2):
    """mvScaleFactors / mvLevelSigma2 exactly as ORBextractor builds them in fp32
    (ORBextractor.cc:488-505): sf[0]=1, sf[i]=sf[i-1]*scaleFactor, sigma2=sf*sf."""
    sf = np.empty(n_levels, dtype=np.float32)
    s2 = np.empty(n_levels, dtype=np.float32)
    sf[0] = np.float32(1.0)
    s2[0] = np.float32(1.0)
    f = np.float32(scale_factor)
    for i in range(1, n_levels):
        sf[i] = np.float32(sf[i - 1] * f)
        s2[i] = np.float32(sf[i] * sf[i])
    return sf, s2


@dataclass
class HostFrame:
    """Flat host view of one Frame / KeyFrame (everything the matcher reads)."""
    desc: np.ndarray            # [n,32] u8
    kp_xy: np.ndarray           # [n,2] f32
    octave: np.ndarray          # [n] i32
    angle: np.ndarray           # [n] f32
    u_right: Optional[np.ndarray] = None
    min_x: float = 0.0
    min_y: float = 0.0
    max_x: float = 640.0
    max_y: float = 480.0
    grid_cols: int = 64
    grid_rows: int = 48
    scale_factors: np.ndarray = field(default_factory=lambda: orb_scale_tables()[0])
    level_sigma2: np.ndarray = field(default_factory=lambda: orb_scale_tables()[1])
    fv_node_ids: Optional[np.ndarray] = None
    fv_offsets: Optional[np.ndarray] = None
    fv_features: Optional[np.ndarray] = None

    def __post_init__(self):
        self.desc = as_u8(self.desc).reshape(-1, 32)
        self.kp_xy = as_f32(self.kp_xy).reshape(-1, 2)
        self.octave = as_i32(self.octave)
        self.angle = as_f32(self.angle)
        if self.u_right is not None:
            self.u_right = as_f32(self.u_right)
        self.scale_factors = as_f32(self.scale_factors)
        self.level_sigma2 = as_f32(self.level_sigma2)
        if self.fv_node_ids is not None:
            self.fv_node_ids = as_u32(self.fv_node_ids)
            self.fv_offsets = as_i32(self.fv_offsets)
            self.fv_features = as_u32(self.fv_features)
        # mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX) in fp32 (Frame.cc:420-422)
        self.grid_inv_w = float(np.float32(self.grid_cols) / (np.float32(self.max_x) - np.float32(self.min_x)))
        self.grid_inv_h = float(np.float32(self.grid_rows) / (np.float32(self.max_y) - np.float32(self.min_y)))

    @property
    def n(self) -> int:
        return int(self.desc.shape[0])

    def with_featvec(self, node_ids, offsets, features) -> "HostFrame":
        import copy
        f = copy.copy(self)
        f.fv_node_ids = as_u32(node_ids)
        f.fv_offsets = as_i32(offsets)
        f.fv_features = as_u32(features)
        return f

    def struct(self) -> FrameHostStruct:
        s = FrameHostStruct()
        s.n = self.n
        s.desc = _ptr(self.desc, u8p)
        s.kp_xy = _ptr(self.kp_xy, f32p)
        s.octave = _ptr(self.octave, i32p)
        s.angle = _ptr(self.angle, f32p)
        s.u_right = _ptr(self.u_right, f32p)
        s.min_x, s.min_y, s.max_x, s.max_y = self.min_x, self.min_y, self.max_x, self.max_y
        s.grid_inv_w, s.grid_inv_h = self.grid_inv_w, self.grid_inv_h
        s.grid_cols, s.grid_rows = self.grid_cols, self.grid_rows
        s.n_levels = int(self.scale_factors.shape[0])
        s.scale_factors = _ptr(self.scale_factors, f32p)
        s.level_sigma2 = _ptr(self.level_sigma2, f32p)
        if self.fv_node_ids is not None:
            s.fv_n_nodes = int(self.fv_node_ids.shape[0])
            s.fv_node_ids = _ptr(self.fv_node_ids, u32p)
            s.fv_offsets = _ptr(self.fv_offsets, i32p)
            s.fv_features = _ptr(self.fv_features, u32p)
        else:
            s.fv_n_nodes = 0
        s._keep = self  # keep arrays alive
        return s


@dataclass
class HostVoc:
    """Flat host view of a DBoW2 vocabulary tree (node 0 = root)."""
    k: int
    L: int
    node_desc: np.ndarray      # [n_nodes,32] u8
    child_offsets: np.ndarray  # [n_nodes+1] i32
    child_ids: np.ndarray      # [..] u32
    weight: np.ndarray         # [n_nodes] f64
    word_id: np.ndarray        # [n_nodes] u32

    def __post_init__(self):
        self.node_desc = as_u8(self.node_desc).reshape(-1, 32)
        self.child_offsets = as_i32(self.child_offsets)
        self.child_ids = as_u32(self.child_ids)
        self.weight = as_f64(self.weight)
        self.word_id = as_u32(self.word_id)

    @property
    def n_nodes(self) -> int:
        return int(self.node_desc.shape[0])

    def struct(self) -> VocHostStruct:
        s = VocHostStruct()
        s.k, s.L, s.n_nodes = self.k, self.L, self.n_nodes
        s.node_desc = _ptr(self.node_desc, u8p)
        s.child_offsets = _ptr(self.child_offsets, i32p)
        s.child_ids = _ptr(self.child_ids, u32p)
        s.weight = _ptr(self.weight, f64p)
        s.word_id = _ptr(self.word_id, u32p)
        s._keep = self
        return s

    def node_levels(self) -> np.ndarray:
        lev = np.zeros(self.n_nodes, dtype=np.int32)
        for i in range(self.n_nodes):  # parents precede children in DBoW2's creation order? not assumed:
            pass
        # BFS from the root
        frontier = [0]
        while frontier:
            nxt = []
            for p in frontier:
                for c in self.child_ids[self.child_offsets[p]:self.child_offsets[p + 1]]:
                    lev[c] = lev[p] + 1
                    nxt.append(int(c))
            frontier = nxt
        return lev

    def save(self, path):
        np.savez_compressed(path, k=self.k, L=self.L, node_desc=self.node_desc, child_offsets=self.child_offsets,
                            child_ids=self.child_ids, weight=self.weight, word_id=self.word_id)

    @staticmethod
    def load(path) -> "HostVoc":
        z = np.load(path)
        return HostVoc(int(z["k"]), int(z["L"]), z["node_desc"], z["child_offsets"], z["child_ids"], z["weight"],
                       z["word_id"])


@dataclass
class HostMapPoints:
    """Flat host view of the local map points SearchByProjection(F, vpMapPoints) reads."""
    desc: np.ndarray
    proj_xy: np.ndarray
    scale_level: np.ndarray
    view_cos: np.ndarray
    depth: np.ndarray
    in_view: np.ndarray
    bad: np.ndarray
    n_obs: np.ndarray
    proj_xr: Optional[np.ndarray] = None

    def __post_init__(self):
        self.desc = as_u8(self.desc).reshape(-1, 32)
        self.proj_xy = as_f32(self.proj_xy).reshape(-1, 2)
        self.scale_level = as_i32(self.scale_level)
        self.view_cos = as_f32(self.view_cos)
        self.depth = as_f32(self.depth)
        self.in_view = as_u8(self.in_view)
        self.bad = as_u8(self.bad)
        self.n_obs = as_i32(self.n_obs)
        if self.proj_xr is not None:
            self.proj_xr = as_f32(self.proj_xr)

    @property
    def n(self) -> int:
        return int(self.desc.shape[0])

    def struct(self) -> MapPointsHostStruct:
        s = MapPointsHostStruct()
        s.n = self.n
        s.desc = _ptr(self.desc, u8p)
        s.proj_xy = _ptr(self.proj_xy, f32p)
        s.proj_xr = _ptr(self.proj_xr, f32p)
        s.scale_level = _ptr(self.scale_level, i32p)
        s.view_cos = _ptr(self.view_cos, f32p)
        s.depth = _ptr(self.depth, f32p)
        s.in_view = _ptr(self.in_view, u8p)
        s.bad = _ptr(self.bad, u8p)
        s.n_obs = _ptr(self.n_obs, i32p)
        s._keep = self
        return s


@dataclass
class HostProjPoints:
    """Points already projected into the target frame: the inputs of the search core shared by the self-projecting
    overloads (SearchByProjection Cur/Last, Cur/KF, KF/Sim3; Fuse; SearchBySim3)."""
    desc: np.ndarray
    uv: np.ndarray
    radius: np.ndarray
    min_level: np.ndarray
    max_level: np.ndarray
    active: np.ndarray
    ur: Optional[np.ndarray] = None
    locks: Optional[np.ndarray] = None
    angle: Optional[np.ndarray] = None

    def __post_init__(self):
        self.desc = as_u8(self.desc).reshape(-1, 32)
        self.uv = as_f32(self.uv).reshape(-1, 2)
        self.radius = as_f32(self.radius)
        self.min_level = as_i32(self.min_level)
        self.max_level = as_i32(self.max_level)
        self.active = as_u8(self.active)
        if self.ur is not None:
            self.ur = as_f32(self.ur)
        if self.locks is not None:
            self.locks = as_u8(self.locks)
        if self.angle is not None:
            self.angle = as_f32(self.angle)

    @property
    def n(self) -> int:
        return int(self.desc.shape[0])

    def struct(self) -> ProjPointsHostStruct:
        s = ProjPointsHostStruct()
        s.n = self.n
        s.desc = _ptr(self.desc, u8p)
        s.uv = _ptr(self.uv, f32p)
        s.radius = _ptr(self.radius, f32p)
        s.min_level = _ptr(self.min_level, i32p)
        s.max_level = _ptr(self.max_level, i32p)
        s.ur = _ptr(self.ur, f32p)
        s.active = _ptr(self.active, u8p)
        s.locks = _ptr(self.locks, u8p)
        s.angle = _ptr(self.angle, f32p)
        s._keep = self
        return s


def proj_params(max_dist, ordered, stereo_gate=0, chi2_gate=0, check_ori=0, inv_level_sigma2=None) -> ProjSearchParamsStruct:
    p = ProjSearchParamsStruct()
    p.max_dist = float(max_dist)
    p.ordered, p.stereo_gate, p.chi2_gate, p.check_ori = int(ordered), int(stereo_gate), int(chi2_gate), int(check_ori)
    inv = as_f32(inv_level_sigma2) if inv_level_sigma2 is not None else None
    p.inv_level_sigma2 = _ptr(inv, f32p)
    p._keep = inv
    return p


@dataclass
class HostBowDb:
    """CSR of the key frames' BowVectors (words ascending inside a key frame): the database KeyFrameDatabase indexes."""
    offsets: np.ndarray
    words: np.ndarray
    values: np.ndarray

    def __post_init__(self):
        self.offsets = as_i32(self.offsets)
        self.words = as_u32(self.words)
        self.values = np.ascontiguousarray(self.values, dtype=np.float64)

    @property
    def n_kf(self) -> int:
        return int(self.offsets.shape[0] - 1)

    def struct(self) -> BowDbHostStruct:
        s = BowDbHostStruct()
        s.n_kf = self.n_kf
        s.offsets = _ptr(self.offsets, i32p)
        s.words = _ptr(self.words, u32p)
        s.values = _ptr(self.values, f64p)
        s._keep = self
        return s


@dataclass
class HostKfSet:
    """Flat host view of a batch of keyframes with n_feat features each (config C4)."""
    desc: np.ndarray      # [n_kf,n_feat,32]
    kp_xy: np.ndarray     # [n_kf,n_feat,2]
    octave: np.ndarray    # [n_kf,n_feat]
    angle: np.ndarray     # [n_kf,n_feat]
    has_mp: np.ndarray    # [n_kf,n_feat] u8
    node_id: np.ndarray   # [n_kf,n_feat] u32 (0xFFFFFFFF == not in the FeatureVector)
    u_right: Optional[np.ndarray] = None
    scale_factors: np.ndarray = field(default_factory=lambda: orb_scale_tables()[0])
    level_sigma2: np.ndarray = field(default_factory=lambda: orb_scale_tables()[1])

    def __post_init__(self):
        self.desc = as_u8(self.desc)
        assert self.desc.ndim == 3 and self.desc.shape[2] == 32
        self.kp_xy = as_f32(self.kp_xy)
        self.octave = as_i32(self.octave)
        self.angle = as_f32(self.angle)
        self.has_mp = as_u8(self.has_mp)
        if self.node_id is not None:  # None: the FeatureVectors are computed on the device (DeviceKfSet.transform)
            self.node_id = as_u32(self.node_id)
        if self.u_right is not None:
            self.u_right = as_f32(self.u_right)
        self.scale_factors = as_f32(self.scale_factors)
        self.level_sigma2 = as_f32(self.level_sigma2)

    @property
    def n_kf(self) -> int:
        return int(self.desc.shape[0])

    @property
    def n_feat(self) -> int:
        return int(self.desc.shape[1])

    def struct(self) -> KfSetHostStruct:
        s = KfSetHostStruct()
        s.n_kf, s.n_feat = self.n_kf, self.n_feat
        s.desc = _ptr(self.desc, u8p)
        s.kp_xy = _ptr(self.kp_xy, f32p)
        s.octave = _ptr(self.octave, i32p)
        s.angle = _ptr(self.angle, f32p)
        s.has_mp = _ptr(self.has_mp, u8p)
        s.u_right = _ptr(self.u_right, f32p)
        s.node_id = _ptr(self.node_id, u32p)
        s.n_levels = int(self.scale_factors.shape[0])
        s.scale_factors = _ptr(self.scale_factors, f32p)
        s.level_sigma2 = _ptr(self.level_sigma2, f32p)
        s._keep = self
        return s
