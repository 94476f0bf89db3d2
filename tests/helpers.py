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
