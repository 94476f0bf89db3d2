"""shared helpers of the test-suite (test infrastructure)"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from orb_slam3_comments_ghr_b200 import synth  # noqa: E402
from orb_slam3_comments_ghr_b200._abi import HostVoc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(*arrays) -> np.ndarray:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def golden_outputs():
    return np.load(os.path.join(GOLD, "reference_outputs.npz"))


def golden_voc() -> HostVoc:
    return HostVoc.load(os.path.join(GOLD, "voc_k10_L4.npz"))


def golden_cases():
    from make_golden import golden_cases as gc
    return gc()


def projected_cases():
    from make_golden import projected_cases as pc
    return pc()


def projected_golden():
    return np.load(os.path.join(GOLD, "projected_outputs.npz"))


def attach_featvec(oracle, voc, frame, levelsup):
    w, nid, wt = oracle.voc_transform(voc, frame.desc, levelsup)
    fn, fo, ff = oracle.featvec(nid, wt)
    return frame.with_featvec(fn, fo, ff)


def add_stereo(tc, seed, frac=0.5):
    """gives a triangulation case right-image coordinates for a fraction of its features (mvuRight >= 0 == stereo keypoint)"""
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(tc.kfs.octave.shape) < frac, tc.kfs.kp_xy[..., 0] - rng.uniform(1, 30, tc.kfs.octave.shape), -1.0)
    tc.kfs.u_right = np.ascontiguousarray(ur, dtype=np.float32)
    return tc


def stereo_projection_case(seed, th):
    """C2 with a stereo frame: mvuRight on 60 % of the keypoints and mTrackProjXR on the map points (:107-117)"""
    from orb_slam3_comments_ghr_b200._abi import HostFrame
    c = synth.make_projection_case(seed, th=th)
    rng = np.random.default_rng(seed + 1)
    f = c.frame
    ur = np.where(rng.random(f.n) < 0.6, np.maximum(f.kp_xy[:, 0] - rng.uniform(1, 30, f.n), 0.25), -1.0).astype(np.float32)
    c.frame = HostFrame(f.desc, f.kp_xy, f.octave, f.angle, u_right=ur, scale_factors=f.scale_factors, level_sigma2=f.level_sigma2)
    c.mps.proj_xr = (c.mps.proj_xy[:, 0] - rng.uniform(1, 30, c.mps.n)).astype(np.float32)
    return c
