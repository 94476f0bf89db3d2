"""Seeded synthetic inputs for the five BASELINE.json configs.

ORBvoc and the datasets are not available offline, so every workload is synthetic:
random 256-bit descriptors with planted bit-flipped matches, keypoints on a 640x480
image quantised to 1/4 px (so fp32 grid math has exact-tie cases), octaves drawn from
ORBextractor's per-level share (ORBextractor.cc:512-532), angles with a planted global
rotation, pinhole intrinsics from Examples/ROS/ORB_SLAM3/orbbec335L_rgbd.yaml:11-14.
All generators are pure numpy and deterministic in `seed`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from ._abi import HostBowDb, HostFrame, HostKfSet, HostMapPoints, HostProjPoints, HostVoc, orb_scale_tables

IMG_W, IMG_H = 640.0, 480.0
# orbbec335L_rgbd.yaml:11-14
FX, FY, CX, CY = 368.05096, 368.05399, 317.11264, 236.39537
K_PINHOLE = np.array([FX, FY, CX, CY], dtype=np.float32)


def random_descriptors(rng: np.random.Generator, n: int) -> np.ndarray:
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def flip_mask(rng: np.random.Generator, n: int, log2_inv_density) -> np.ndarray:
    """[n,32] u8 masks whose bits are set with probability 2**-k (k per row or scalar):
    XOR-ing one onto a descriptor flips ~256*2**-k bits.  k >= 9 rows stay all-zero."""
    k = np.broadcast_to(np.asarray(log2_inv_density, dtype=np.int64), (n,))
    m = np.full((n, 32), 255, dtype=np.uint8)
    for level in range(1, int(k.max(initial=0)) + 1):
        r = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        sel = (k >= level)[:, None]
        m = np.where(sel, m & r, m)
    m[k <= 0] = 0
    m[k >= 9] = 0
    return m


def planted_copies(rng: np.random.Generator, base: np.ndarray) -> np.ndarray:
    """copies of `base` with ~0..40 flipped bits (densities 1/8, 1/16, 1/32, or exact copy)."""
    n = base.shape[0]
    k = rng.choice(np.array([3, 4, 4, 5, 5, 9]), size=n)
    return base ^ flip_mask(rng, n, k)


def octave_shares(n_levels: int = 8, scale_factor: float = 1.2) -> np.ndarray:
    """per-level feature share of ORBextractor (ORBextractor.cc:515-532): geometric in 1/scaleFactor."""
    f = 1.0 / scale_factor
    w = f ** np.arange(n_levels)
    return w / w.sum()


def random_octaves(rng, n, n_levels=8):
    return rng.choice(n_levels, size=n, p=octave_shares(n_levels)).astype(np.int32)


def quantise(xy):
    return (np.round(np.asarray(xy, dtype=np.float64) * 4.0) / 4.0).astype(np.float32)


def random_keypoints(rng, n):
    xy = np.stack([rng.uniform(0, IMG_W, n), rng.uniform(0, IMG_H, n)], axis=1)
    xy = quantise(xy)
    xy[:, 0] = np.clip(xy[:, 0], 0, IMG_W - 0.25)
    xy[:, 1] = np.clip(xy[:, 1], 0, IMG_H - 0.25)
    return xy


def make_frame(rng, n, n_levels=8) -> HostFrame:
    sf, s2 = orb_scale_tables(n_levels)
    return HostFrame(desc=random_descriptors(rng, n), kp_xy=random_keypoints(rng, n), octave=random_octaves(rng, n, n_levels),
                     angle=quantise(rng.uniform(0, 360, n) % 360.0), scale_factors=sf, level_sigma2=s2)


# ---------------------------------------------------------------- C1
@dataclass
class InitCase:
    f1: HostFrame
    f2: HostFrame
    prev_matched: np.ndarray
    window_size: int = 100
    nnratio: float = 0.9
    check_ori: int = 1


def make_init_case(seed: int, n: int = 1000, window_size: int = 100, max_disp: float = 60.0, dup_frac: float = 0.05) -> InitCase:
    """C1: SearchForInitialization on two n-keypoint frames; F2 = F1 displaced by U(-max_disp,max_disp) px,
    planted descriptors, global rotation + N(0,5 deg); vbPrevMatched = F1 points (Tracking.cc:2916).
    A fraction of F2 features are near-duplicates of each other so that displaced matches
    (ORBmatcher.cc:813-817) and ratio-test failures occur."""
    rng = np.random.default_rng(seed)
    f1 = make_frame(rng, n)
    perm = rng.permutation(n)
    n_pl = int(0.7 * n)
    src = perm[:n_pl]
    desc2 = random_descriptors(rng, n)
    xy2 = random_keypoints(rng, n)
    oct2 = random_octaves(rng, n)
    ang2 = quantise(rng.uniform(0, 360, n) % 360.0)
    dst = rng.permutation(n)[:n_pl]
    desc2[dst] = planted_copies(rng, f1.desc[src])
    disp = rng.uniform(-max_disp, max_disp, size=(n_pl, 2))
    xy2[dst] = quantise(np.clip(f1.kp_xy[src] + disp, [0, 0], [IMG_W - 0.25, IMG_H - 0.25]))
    oct2[dst] = f1.octave[src]
    rot = 37.0
    ang2[dst] = quantise((f1.angle[src] - rot + rng.normal(0, 5.0, n_pl)) % 360.0)
    # near duplicates: another F2 feature close by with almost the same descriptor
    n_dup = int(dup_frac * n)
    if n_dup > 0:
        a = dst[rng.integers(0, n_pl, n_dup)]
        free = np.setdiff1d(np.arange(n), dst)
        b = free[rng.permutation(free.shape[0])[:n_dup]]
        a = a[:b.shape[0]]
        desc2[b] = desc2[a] ^ flip_mask(rng, b.shape[0], 5)
        xy2[b] = quantise(np.clip(xy2[a] + rng.uniform(-20, 20, size=(b.shape[0], 2)), [0, 0], [IMG_W - 0.25, IMG_H - 0.25]))
        oct2[b] = oct2[a]
        ang2[b] = ang2[a]
    f2 = HostFrame(desc=desc2, kp_xy=xy2, octave=oct2, angle=ang2)
    return InitCase(f1, f2, f1.kp_xy.copy(), window_size)


# ---------------------------------------------------------------- C2
@dataclass
class ProjectionCase:
    frame: HostFrame
    mps: HostMapPoints
    kp_prior_obs: np.ndarray
    kp_mp: np.ndarray
    th: float = 1.0
    far_points: int = 0
    th_far: float = 50.0
    nnratio: float = 0.8


def make_projection_case(seed: int, n_kp: int = 2000, n_mp: int = 5000, th: float = 1.0, n_planted: Optional[int] = None,
                         far_points: int = 0) -> ProjectionCase:
    """C2: SearchByProjection(Frame, local MapPoints).  n_planted map points project onto keypoints
    (+N(0,1.5 px)) with predicted level = kp.octave or +1; the rest are random in-image.  Several map
    points are planted on the same keypoint so that the sequential skip rule (ORBmatcher.cc:102-104)
    and overwrites matter; 10 % of keypoints already hold an observed map point."""
    rng = np.random.default_rng(seed)
    frame = make_frame(rng, n_kp)
    if n_planted is None:
        n_planted = int(0.3 * n_mp)
    n_levels = frame.scale_factors.shape[0]
    desc = random_descriptors(rng, n_mp)
    proj = random_keypoints(rng, n_mp).astype(np.float32)
    level = random_octaves(rng, n_mp)
    tgt = rng.integers(0, n_kp, n_planted)  # with repetition: several MPs per keypoint
    who = rng.permutation(n_mp)[:n_planted]
    desc[who] = planted_copies(rng, frame.desc[tgt])
    proj[who] = quantise(frame.kp_xy[tgt] + rng.normal(0, 1.5, size=(n_planted, 2)))
    level[who] = np.clip(frame.octave[tgt] + rng.integers(0, 2, n_planted), 0, n_levels - 1)
    view_cos = rng.uniform(0.9, 1.0, n_mp).astype(np.float32)
    view_cos[rng.random(n_mp) < 0.2] = np.float32(0.9995)
    depth = rng.uniform(0.5, 80.0, n_mp).astype(np.float32)
    in_view = (rng.random(n_mp) < 0.9).astype(np.uint8)
    bad = (rng.random(n_mp) < 0.03).astype(np.uint8)
    n_obs = rng.integers(1, 6, n_mp).astype(np.int32)
    n_obs[rng.random(n_mp) < 0.1] = 0  # unobserved points may be overwritten later
    mps = HostMapPoints(desc, proj, level, view_cos, depth, in_view, bad, n_obs)
    prior = np.zeros(n_kp, dtype=np.int32)
    pre = rng.random(n_kp) < 0.1
    prior[pre] = rng.integers(0, 4, int(pre.sum()))
    kp_mp = np.full(n_kp, -1, dtype=np.int32)
    return ProjectionCase(frame, mps, prior, kp_mp, th=th, far_points=far_points)


def make_projected_case(seed: int, n_kp: int = 2000, n_pts: int = 3000, th: float = 7.0, stereo: bool = False,
                        level_mode: str = "pm1", planted_frac: float = 0.5, lock_frac: float = 0.9):
    """Inputs of the search core shared by the self-projecting overloads (row a6): points already projected into a
    frame.  planted points sit on keypoints (+N(0,2 px)) with a bit-flipped copy of the keypoint's descriptor, several
    per keypoint so that the ordered skip rule matters; the others are random in-image.  level_mode: 'pm1' = window
    [l-1, l+1] (Cur/Last, :2041), 'fwd' = (l, -1) (:2035), 'bwd' = (0, l) (:2038), 'pred' = (l-1, l) (KeyFrame callers)."""
    rng = np.random.default_rng(seed)
    frame = make_frame(rng, n_kp)
    n_levels = frame.scale_factors.shape[0]
    if stereo:
        ur = quantise(np.maximum(frame.kp_xy[:, 0] - rng.uniform(2, 40, n_kp), 0.25)).astype(np.float32)
        ur[rng.random(n_kp) >= 0.6] = np.float32(-1.0)  # keypoints without a stereo match (mvuRight = -1)
        frame = HostFrame(frame.desc, frame.kp_xy, frame.octave, frame.angle, u_right=ur,
                          scale_factors=frame.scale_factors, level_sigma2=frame.level_sigma2)
    desc = random_descriptors(rng, n_pts)
    uv = random_keypoints(rng, n_pts).astype(np.float32)
    level = random_octaves(rng, n_pts)
    n_pl = int(planted_frac * n_pts)
    tgt = rng.integers(0, n_kp, n_pl)
    who = rng.permutation(n_pts)[:n_pl]
    desc[who] = planted_copies(rng, frame.desc[tgt])
    uv[who] = quantise(frame.kp_xy[tgt] + rng.normal(0, 2.0, size=(n_pl, 2)))
    level[who] = np.clip(frame.octave[tgt] + rng.integers(-1, 2, n_pl), 0, n_levels - 1)
    radius = (np.float32(th) * frame.scale_factors[level]).astype(np.float32)
    if level_mode == "pm1":
        lo, hi = level - 1, level + 1
    elif level_mode == "fwd":
        lo, hi = level, np.full(n_pts, -1)
    elif level_mode == "bwd":
        lo, hi = np.zeros(n_pts, dtype=np.int64), level
    else:
        lo, hi = level - 1, level
    pur = None
    if stereo:
        pur = (uv[:, 0] - rng.uniform(2, 40, n_pts)).astype(np.float32)
        has = frame.u_right[tgt] > 0
        pur[who[has]] = frame.u_right[tgt[has]] + rng.normal(0, 3.0, int(has.sum())).astype(np.float32)
    active = (rng.random(n_pts) < 0.9).astype(np.uint8)
    locks = (rng.random(n_pts) < lock_frac).astype(np.uint8)
    angle = quantise(rng.uniform(0, 360, n_pts) % 360.0).astype(np.float32)
    angle[who] = quantise((frame.angle[tgt] + 20.0 + rng.normal(0, 4.0, n_pl)) % 360.0)
    odd = who[rng.random(n_pl) < 0.15]  # wrong rotation: these matches fall outside the three dominant bins
    angle[odd] = quantise(rng.uniform(0, 360, odd.size) % 360.0)
    pts = HostProjPoints(desc, uv, radius, lo, hi, active, ur=pur, locks=locks, angle=angle)
    kp_locked = (rng.random(n_kp) < 0.1).astype(np.uint8)
    return frame, pts, kp_locked


def make_stereo_case(seed: int, n: int = 2000, planted_frac: float = 0.7):
    """A rectified stereo pair: right keypoints are left keypoints shifted by a disparity in [0, mbf/mb] on (nearly) the same
    row, with bit-flipped descriptors and octave within +-1; plus clutter, keypoints on the image border rows and ties."""
    rng = np.random.default_rng(seed)
    left, right = make_frame(rng, n), make_frame(rng, n)
    mb, mbf = 0.11, 40.0
    n_pl = int(planted_frac * n)
    src = rng.permutation(n)[:n_pl]
    dst = rng.permutation(n)[:n_pl]
    right.desc[dst] = planted_copies(rng, left.desc[src])
    disp = rng.uniform(-5, mbf / mb + 20, n_pl)  # some outside [0, maxD]
    right.kp_xy[dst, 0] = quantise(np.clip(left.kp_xy[src, 0] - disp, 0, IMG_W - 0.25))
    right.kp_xy[dst, 1] = quantise(np.clip(left.kp_xy[src, 1] + rng.normal(0, 1.5, n_pl), 0, IMG_H - 0.25))
    right.octave[dst] = np.clip(left.octave[src] + rng.integers(-2, 3, n_pl), 0, 7)
    k = min(8, n)
    left.kp_xy[:k, 1] = np.array([0.0, 0.25, 479.75, 479.0, 1.0, 478.5, 0.75, 2.0], dtype=np.float32)[:k]  # border rows
    right.desc[dst[:20]] = left.desc[src[:20]]  # exact copies: distance-0 ties with duplicates below
    right.desc[(dst[:20] + 1) % n] = left.desc[src[:20]]
    right.kp_xy[(dst[:20] + 1) % n] = right.kp_xy[dst[:20]]
    right.octave[(dst[:20] + 1) % n] = right.octave[dst[:20]]
    return left, right, IMG_H, mb, mbf


def make_distinctive_case(seed: int, n_mp: int = 3000, max_obs: int = 40):
    """Observation descriptors of n_mp map points (CSR): noisy copies of a per-point prototype, a few outliers, some points with
    1-2 observations (median index 0), a few empty lists, duplicates (distance 0 ties) and one long list."""
    rng = np.random.default_rng(seed)
    offs, descs = [0], []
    for p in range(n_mp):
        n = 0 if p % 211 == 7 else int(rng.integers(1, max_obs + 1))
        if p == n_mp // 2:
            n = 1500  # longer than the shared-memory staging of the kernel
        proto = random_descriptors(rng, 1)
        d = np.repeat(proto, n, axis=0) ^ flip_mask(rng, n, rng.choice(np.array([3, 4, 5]), size=n)) if n else np.zeros((0, 32), np.uint8)
        if n > 3:
            d[rng.integers(0, n)] = random_descriptors(rng, 1)[0]      # an outlier observation
            if p % 5 == 0:
                d[rng.integers(0, n)] = d[rng.integers(0, n)]         # exact duplicates
        descs.append(d)
        offs.append(offs[-1] + n)
    return np.array(offs, dtype=np.int32), np.concatenate(descs, axis=0)


def make_bowdb_case(seed: int, n_kf: int = 2000, n_words_voc: int = 10000, words_per_kf: int = 800):
    """A key-frame database of L1-normalised BowVectors over a vocabulary of n_words_voc words plus a query that shares many
    words with a few key frames (the relocalisation / loop candidates) and few with the rest; sizes vary per key frame and a
    few are empty or disjoint from the query."""
    rng = np.random.default_rng(seed)
    offs, words, vals = [0], [], []
    pop = rng.zipf(1.3, size=n_words_voc).astype(np.float64)  # popular words, as in a real vocabulary
    pop /= pop.sum()

    def bow(n):
        w = np.unique(rng.choice(n_words_voc, size=n, p=pop))
        v = rng.gamma(2.0, 1.0, size=w.shape[0])
        return w.astype(np.uint32), v / v.sum()

    qw, qv = bow(words_per_kf + 200)
    for k in range(n_kf):
        n = int(rng.integers(1, 2 * words_per_kf)) if k % 97 else 0
        w, v = bow(n) if n else (np.zeros(0, np.uint32), np.zeros(0))
        if k % 50 == 3 and w.size:  # a near-duplicate of the query: perturbed weights on a subset of its words
            keep = rng.random(qw.shape[0]) < 0.8
            w, v = qw[keep].copy(), qv[keep] * rng.uniform(0.5, 1.5, int(keep.sum()))
            v = v / v.sum()
        words.append(w); vals.append(v); offs.append(offs[-1] + w.shape[0])
    return HostBowDb(np.array(offs), np.concatenate(words), np.concatenate(vals)), qw, qv


# ---------------------------------------------------------------- vocabulary
def random_vocabulary(seed: int, k: int = 10, L: int = 4, stop_frac: float = 0.02, ragged: bool = False) -> HostVoc:
    """A synthetic k-ary, depth-L Hamming tree in DBoW2's flat layout (node 0 = root).  Children
    descriptors are their parent's with ~1/8 of the bits flipped (so descents are decisive but
    ties still occur); leaf weights are idf-like positive doubles, a fraction exactly 0 (stopped
    words, TemplatedVocabulary.h:1157).  ragged=True drops some children (k is only an upper bound,
    TemplatedVocabulary.h:1237-1248 walks whatever `children` holds).
    For the tree built by the reference's own create() see scripts/make_golden.py."""
    rng = np.random.default_rng(seed)
    descs = [np.zeros((1, 32), dtype=np.uint8)]
    children = [[]]
    frontier = [0]
    root_desc = random_descriptors(rng, 1)
    descs[0] = root_desc
    n_nodes = 1
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            kk = k if not ragged else int(rng.integers(max(2, k - 3), k + 1))
            pd = descs[p] if level > 1 else random_descriptors(rng, 1)
            if level == 1:
                cd = random_descriptors(rng, kk)
            else:
                cd = pd ^ flip_mask(rng, kk, 3)
            for j in range(kk):
                descs.append(cd[j:j + 1])
                children.append([])
                children[p].append(n_nodes)
                nxt.append(n_nodes)
                n_nodes += 1
        frontier = nxt
    node_desc = np.concatenate(descs, axis=0)
    off = np.zeros(n_nodes + 1, dtype=np.int32)
    ch = []
    for i in range(n_nodes):
        off[i + 1] = off[i] + len(children[i])
        ch.extend(children[i])
    child_ids = np.asarray(ch, dtype=np.uint32)
    weight = np.zeros(n_nodes, dtype=np.float64)
    word_id = np.zeros(n_nodes, dtype=np.uint32)
    leaves = np.asarray(frontier, dtype=np.int64)
    weight[leaves] = np.log(rng.uniform(1.5, 400.0, leaves.shape[0]))
    weight[leaves[rng.random(leaves.shape[0]) < stop_frac]] = 0.0
    word_id[leaves] = np.arange(leaves.shape[0], dtype=np.uint32)
    return HostVoc(k, L, node_desc, off, child_ids, weight, word_id)


def descriptors_near_words(rng, voc: HostVoc, n: int, n_distinct: Optional[int] = None) -> np.ndarray:
    """descriptors scattered around the vocabulary's leaves (so they land in many different nodes)."""
    leaves = np.nonzero(np.diff(voc.child_offsets) == 0)[0]
    leaves = leaves[leaves > 0]
    pick = leaves[rng.integers(0, leaves.shape[0], n)]
    return voc.node_desc[pick] ^ flip_mask(rng, n, rng.choice(np.array([3, 4, 5]), size=n))


# ---------------------------------------------------------------- C3
@dataclass
class BowCase:
    kf: HostFrame
    f: HostFrame
    kf_mp_valid: np.ndarray
    f_mp_valid: np.ndarray
    nnratio: float = 0.7
    check_ori: int = 1


def make_bow_case(seed: int, voc: HostVoc, n: int = 2000, valid_frac: float = 0.6) -> BowCase:
    """C3: a keyframe and a frame with n features each; 60 % of the frame's features are planted copies
    of keyframe features (global rotation + noise on the angles); FeatureVectors are NOT filled here --
    run transform (GPU or oracle) and attach with HostFrame.with_featvec."""
    rng = np.random.default_rng(seed)
    kf = make_frame(rng, n)
    kf.desc[:] = descriptors_near_words(rng, voc, n)
    f = make_frame(rng, n)
    f.desc[:] = descriptors_near_words(rng, voc, n)
    n_pl = int(0.6 * n)
    src = rng.permutation(n)[:n_pl]
    dst = rng.permutation(n)[:n_pl]
    f.desc[dst] = planted_copies(rng, kf.desc[src])
    f.angle[dst] = quantise((kf.angle[src] - 21.0 + rng.normal(0, 5.0, n_pl)) % 360.0)
    # duplicates inside the frame to trigger ratio failures and "already matched" skips
    n_dup = n // 20
    a = dst[rng.integers(0, n_pl, n_dup)]
    b = rng.integers(0, n, n_dup)
    f.desc[b] = f.desc[a] ^ flip_mask(rng, n_dup, 5)
    kf_valid = (rng.random(n) < valid_frac).astype(np.uint8)
    f_valid = (rng.random(n) < valid_frac).astype(np.uint8)
    return BowCase(kf, f, kf_valid, f_valid)


def make_bow_conflict_case(seed: int, n1: int = 1200, n2: int = 1500, group: int = 10, layout: str = "root") -> BowCase:
    """SearchByBoW stress for the "partner already matched" rule (:335 / :962): the frame holds groups of `group` near-duplicate
    descriptors and many keyframe features sit near each group's prototype, so later keyframe features find their best
    partners taken and walk down their candidate lists (past 8 entries for the big groups).  FeatureVectors are attached here:
    layout "root" = one node holding everything (levelsup >= L), "mixed" = one big node, one mid-size node and small ones."""
    rng = np.random.default_rng(seed)
    n_groups = max(n2 // group, 1)
    proto = random_descriptors(rng, n_groups)
    f = make_frame(rng, n2)
    gid2 = np.arange(n2) % n_groups
    f.desc[:] = proto[gid2] ^ flip_mask(rng, n2, rng.choice(np.array([4, 5, 6]), size=n2))
    kf = make_frame(rng, n1)
    gid1 = rng.integers(0, n_groups, n1)
    gid1[: n1 // 8] = 0  # one crowded group: more contenders than members
    kf.desc[:] = proto[gid1] ^ flip_mask(rng, n1, rng.choice(np.array([5, 6, 7, 9]), size=n1))
    kf.angle[:] = quantise((f.angle[rng.integers(0, n2, n1)] + 33.0 + rng.normal(0, 4.0, n1)) % 360.0)
    kf_valid = (rng.random(n1) < 0.8).astype(np.uint8)
    f_valid = (rng.random(n2) < 0.8).astype(np.uint8)

    def featvec(n, cuts):
        order = rng.permutation(n).astype(np.uint32)
        bounds = [0] + [int(c * n) for c in cuts] + [n]
        feats = np.concatenate([np.sort(order[bounds[i]:bounds[i + 1]]) for i in range(len(bounds) - 1)])  # ascending in a node
        keep = [i for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]
        return (np.array([10 + 3 * i for i in keep], dtype=np.uint32),
                np.array([bounds[i] for i in keep] + [n], dtype=np.int32), feats)

    cuts = [] if layout == "root" else [0.55, 0.8, 0.82, 0.9]
    kf = kf.with_featvec(*featvec(n1, cuts))
    f = f.with_featvec(*featvec(n2, cuts))
    return BowCase(kf, f, kf_valid, f_valid)


# ---------------------------------------------------------------- C4
@dataclass
class TriangulationCase:
    kfs: HostKfSet
    kf1: np.ndarray       # [P] i32
    kf2: np.ndarray       # [P] i32
    T1w: np.ndarray       # [P,12] f32  (R row-major | t)
    T2w: np.ndarray       # [P,12] f32
    K: np.ndarray         # [4] fx fy cx cy
    ep: Optional[np.ndarray] = None    # [P,2]  filled by geometry()
    f12: Optional[np.ndarray] = None   # [P,9]
    nnratio: float = 0.6
    check_ori: int = 0


def _small_rotation(rng, n, max_angle=0.08):
    w = rng.normal(0, 1, size=(n, 3))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    th = rng.uniform(0, max_angle, n)
    Kx = np.zeros((n, 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -w[:, 2], w[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = w[:, 2], -w[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -w[:, 1], w[:, 0]
    I = np.eye(3)[None]
    return I + np.sin(th)[:, None, None] * Kx + (1 - np.cos(th))[:, None, None] * (Kx @ Kx)


def make_triangulation_case(seed: int, n_pairs: int = 64, n_feat: int = 2000, n_nodes: int = 100,
                            mp_frac: float = 0.5, node_base: int = 11) -> TriangulationCase:
    """C4: n_pairs independent keyframe pairs (2*n_pairs keyframes).  Each pair observes a common set of
    3-D landmarks (so planted matches satisfy the epipolar constraint up to the 1/4 px quantisation +
    N(0,0.7 px) noise), padded with random features.  50 % of the features already have a map point.
    FeatureVector node ids: landmarks carry a node id in [node_base, node_base+n_nodes) (level-2 ids of a
    k=10 tree start at 11); 90 % of the observations keep it (quantisation to different words otherwise)."""
    rng = np.random.default_rng(seed)
    P = n_pairs
    nk = 2 * P
    sf, s2 = orb_scale_tables(8)
    desc = random_descriptors(rng, nk * n_feat).reshape(nk, n_feat, 32)
    xy = random_keypoints(rng, nk * n_feat).reshape(nk, n_feat, 2)
    octv = random_octaves(rng, nk * n_feat).reshape(nk, n_feat)
    ang = quantise(rng.uniform(0, 360, nk * n_feat) % 360.0).reshape(nk, n_feat)
    node = (node_base + rng.integers(0, n_nodes, nk * n_feat)).astype(np.uint32).reshape(nk, n_feat)
    has_mp = (rng.random((nk, n_feat)) < mp_frac).astype(np.uint8)
    # poses: camera 1 near identity, camera 2 a small baseline away
    R1 = _small_rotation(rng, P)
    R2 = _small_rotation(rng, P)
    t1 = rng.normal(0, 0.05, size=(P, 3))
    t2 = t1 + rng.normal(0, 0.25, size=(P, 3))
    n_lm = int(0.6 * n_feat)
    X = np.stack([rng.uniform(-4, 4, (P, n_lm)), rng.uniform(-3, 3, (P, n_lm)), rng.uniform(2.5, 12, (P, n_lm))], axis=2)
    lm_desc = random_descriptors(rng, P * n_lm).reshape(P, n_lm, 32)
    lm_node = (node_base + rng.integers(0, n_nodes, (P, n_lm))).astype(np.uint32)
    lm_oct = random_octaves(rng, P * n_lm).reshape(P, n_lm)
    lm_ang = rng.uniform(0, 360, (P, n_lm))
    for cam, (R, t) in enumerate(((R1, t1), (R2, t2))):
        Xc = np.einsum("pij,pnj->pni", R, X) + t[:, None, :]
        u = FX * Xc[..., 0] / Xc[..., 2] + CX + rng.normal(0, 0.7, (P, n_lm))
        v = FY * Xc[..., 1] / Xc[..., 2] + CY + rng.normal(0, 0.7, (P, n_lm))
        ok = (Xc[..., 2] > 0.5) & (u >= 0) & (u < IMG_W - 0.25) & (v >= 0) & (v < IMG_H - 0.25)
        slot = np.argsort(rng.random((P, n_feat)), axis=1)[:, :n_lm]  # where each landmark lands in the KF
        kf_idx = (2 * np.arange(P) + cam)[:, None].repeat(n_lm, 1)
        sel = ok
        d = lm_desc ^ flip_mask(rng, P * n_lm, rng.choice(np.array([3, 4, 4, 5, 5, 9]), size=P * n_lm)).reshape(P, n_lm, 32)
        desc[kf_idx[sel], slot[sel]] = d[sel]
        xy[kf_idx[sel], slot[sel], 0] = quantise(u[sel])
        xy[kf_idx[sel], slot[sel], 1] = quantise(v[sel])
        octv[kf_idx[sel], slot[sel]] = np.clip(lm_oct[sel] + rng.integers(-1, 2, int(sel.sum())), 0, 7)
        ang[kf_idx[sel], slot[sel]] = quantise((lm_ang[sel] + cam * 15.0 + rng.normal(0, 4.0, int(sel.sum()))) % 360.0)
        keep = rng.random((P, n_lm)) < 0.9
        node[kf_idx[sel & keep], slot[sel & keep]] = lm_node[sel & keep]
    # a few features are not in the FeatureVector at all (stopped words)
    node[rng.random((nk, n_feat)) < 0.01] = np.uint32(0xFFFFFFFF)
    kfs = HostKfSet(desc, xy, octv, ang, has_mp, node, scale_factors=sf, level_sigma2=s2)
    T1w = np.concatenate([R1.reshape(P, 9), t1], axis=1).astype(np.float32)
    T2w = np.concatenate([R2.reshape(P, 9), t2], axis=1).astype(np.float32)
    kf1 = (2 * np.arange(P)).astype(np.int32)
    kf2 = (2 * np.arange(P) + 1).astype(np.int32)
    return TriangulationCase(kfs, kf1, kf2, T1w, T2w, K_PINHOLE.copy())


def make_triangulation_case_shared(seed: int, n_kf: int = 512, n_neighbours: int = 8, n_feat: int = 2000, n_nodes: int = 100,
                                   mp_frac: float = 0.5, node_base: int = 11) -> TriangulationCase:
    """C4, shared-key-frame variant: the shape of LocalMapping::CreateNewMapPoints (LocalMapping.cc:556-630) -- every
    key frame is matched against its n_neighbours best covisible key frames.  n_kf key frames along a trajectory observe a common
    landmark cloud (a sliding window of it each, so neighbours share most of their landmarks); pair list = (i, i + j) for j = 1 ..
    n_neighbours (wrapping around), i.e. n_kf * n_neighbours pairs over only n_kf key frames: the key-frame set is small (32 MB at
    512 x 2000) and every key frame is read 2 * n_neighbours times per step."""
    rng = np.random.default_rng(seed)
    nk = n_kf
    sf, s2 = orb_scale_tables(8)
    desc = random_descriptors(rng, nk * n_feat).reshape(nk, n_feat, 32)
    xy = random_keypoints(rng, nk * n_feat).reshape(nk, n_feat, 2)
    octv = random_octaves(rng, nk * n_feat).reshape(nk, n_feat)
    ang = quantise(rng.uniform(0, 360, nk * n_feat) % 360.0).reshape(nk, n_feat)
    node = (node_base + rng.integers(0, n_nodes, nk * n_feat)).astype(np.uint32).reshape(nk, n_feat)
    has_mp = (rng.random((nk, n_feat)) < mp_frac).astype(np.uint8)
    # trajectory: the camera drifts sideways through a long landmark cloud, 0.12 m per key frame, small rotations
    R = _small_rotation(rng, nk, max_angle=0.05)
    cx_world = 0.12 * np.arange(nk)
    n_win = int(0.6 * n_feat)          # landmarks seen by one key frame
    step = 30                          # new landmarks per key frame -> neighbours j apart share n_win - 30 j of them
    n_lm = n_win + step * nk
    lm_x0 = np.arange(n_lm) / step * 0.12  # a landmark enters the view when the camera reaches it
    X = np.stack([lm_x0 + rng.uniform(-4, 4, n_lm), rng.uniform(-3, 3, n_lm), rng.uniform(2.5, 12, n_lm)], axis=1)
    lm_desc = random_descriptors(rng, n_lm)
    lm_node = (node_base + rng.integers(0, n_nodes, n_lm)).astype(np.uint32)
    lm_oct = random_octaves(rng, n_lm)
    lm_ang = rng.uniform(0, 360, n_lm)
    t_all = np.zeros((nk, 3))
    for k in range(nk):
        Cw = np.array([cx_world[k] + 0.5 * n_win / step * 0.12, 0.0, 0.0]) + rng.normal(0, 0.03, 3)
        t = -R[k] @ Cw
        t_all[k] = t
        lm = np.arange(k * step, k * step + n_win)
        Xc = X[lm] @ R[k].T + t
        u = FX * Xc[:, 0] / Xc[:, 2] + CX + rng.normal(0, 0.7, n_win)
        v = FY * Xc[:, 1] / Xc[:, 2] + CY + rng.normal(0, 0.7, n_win)
        ok = (Xc[:, 2] > 0.5) & (u >= 0) & (u < IMG_W - 0.25) & (v >= 0) & (v < IMG_H - 0.25)
        slot = rng.permutation(n_feat)[:n_win]
        d = lm_desc[lm] ^ flip_mask(rng, n_win, rng.choice(np.array([3, 4, 4, 5, 5, 9]), size=n_win))
        desc[k, slot[ok]] = d[ok]
        xy[k, slot[ok], 0] = quantise(u[ok])
        xy[k, slot[ok], 1] = quantise(v[ok])
        octv[k, slot[ok]] = np.clip(lm_oct[lm][ok] + rng.integers(-1, 2, int(ok.sum())), 0, 7)
        ang[k, slot[ok]] = quantise((lm_ang[lm][ok] + rng.normal(0, 4.0, int(ok.sum()))) % 360.0)
        keep = ok & (rng.random(n_win) < 0.9)
        node[k, slot[keep]] = lm_node[lm][keep]
    node[rng.random((nk, n_feat)) < 0.01] = np.uint32(0xFFFFFFFF)
    kfs = HostKfSet(desc, xy, octv, ang, has_mp, node, scale_factors=sf, level_sigma2=s2)
    kf1 = np.repeat(np.arange(nk), n_neighbours).astype(np.int32)
    kf2 = ((kf1 + np.tile(np.arange(1, n_neighbours + 1), nk)) % nk).astype(np.int32)
    Tw = np.concatenate([R.reshape(nk, 9), t_all], axis=1).astype(np.float32)
    return TriangulationCase(kfs, kf1, kf2, Tw[kf1], Tw[kf2], K_PINHOLE.copy())


def triangulation_geometry_numpy(T1w, T2w, K1, K2):
    """fp32 restatement of the host-side pose algebra of ORBmatcher.cc:1053-1071 + Pinhole.cpp:194-197 in
    plain left-to-right matrix form: ep = project(T2w * Cw), F12 = K1^-T [t12]x R12 K2^-1.
    (Sophus/Eigen are un-vendored: their exact fp32 rounding is not pinned -- DESIGN.md; the product takes
    ep and F12 as INPUTS so whatever the host computes is what both CPU and GPU consume.)"""
    f = np.float32

    def mat(T):
        return np.asarray(T[:9], dtype=f).reshape(3, 3), np.asarray(T[9:12], dtype=f)

    def mm(A, B):
        C = np.zeros((3, 3), dtype=f)
        for r in range(3):
            for c in range(3):
                C[r, c] = f(f(f(A[r, 0] * B[0, c]) + f(A[r, 1] * B[1, c])) + f(A[r, 2] * B[2, c]))
        return C

    def mv(A, x):
        return np.array([f(f(f(A[r, 0] * x[0]) + f(A[r, 1] * x[1])) + f(A[r, 2] * x[2])) for r in range(3)], dtype=f)

    def inv3(m):
        c = np.zeros((3, 3), dtype=f)
        c[0, 0] = f(f(m[1, 1] * m[2, 2]) - f(m[1, 2] * m[2, 1])); c[0, 1] = f(f(m[0, 2] * m[2, 1]) - f(m[0, 1] * m[2, 2]))
        c[0, 2] = f(f(m[0, 1] * m[1, 2]) - f(m[0, 2] * m[1, 1])); c[1, 0] = f(f(m[1, 2] * m[2, 0]) - f(m[1, 0] * m[2, 2]))
        c[1, 1] = f(f(m[0, 0] * m[2, 2]) - f(m[0, 2] * m[2, 0])); c[1, 2] = f(f(m[0, 2] * m[1, 0]) - f(m[0, 0] * m[1, 2]))
        c[2, 0] = f(f(m[1, 0] * m[2, 1]) - f(m[1, 1] * m[2, 0])); c[2, 1] = f(f(m[0, 1] * m[2, 0]) - f(m[0, 0] * m[2, 1]))
        c[2, 2] = f(f(m[0, 0] * m[1, 1]) - f(m[0, 1] * m[1, 0]))
        det = f(f(f(m[0, 0] * c[0, 0]) + f(m[0, 1] * c[1, 0])) + f(m[0, 2] * c[2, 0]))
        return (c * f(f(1.0) / det)).astype(f)

    R1, t1 = mat(T1w)
    R2, t2 = mat(T2w)
    R1t = R1.T.copy()
    Cw = (-mv(R1t, t1)).astype(f)               # GetCameraCenter = Tcw^-1 translation
    C2 = (mv(R2, Cw) + t2).astype(f)            # T2w * Cw
    ep = np.array([f(f(f(K2[0] * C2[0]) / C2[2]) + K2[2]), f(f(f(K2[1] * C2[1]) / C2[2]) + K2[3])], dtype=f)
    R2t = R2.T.copy()
    tw2 = (-mv(R2t, t2)).astype(f)              # Tw2 = T2w^-1
    R12 = mm(R1, R2t)
    t12 = (mv(R1, tw2) + t1).astype(f)
    tx = np.zeros((3, 3), dtype=f)
    tx[0, 1], tx[0, 2], tx[1, 0], tx[1, 2], tx[2, 0], tx[2, 1] = -t12[2], t12[1], t12[2], -t12[0], -t12[1], t12[0]
    Km1 = np.array([[K1[0], 0, K1[2]], [0, K1[1], K1[3]], [0, 0, 1]], dtype=f)
    Km2 = np.array([[K2[0], 0, K2[2]], [0, K2[1], K2[3]], [0, 0, 1]], dtype=f)
    F = mm(mm(mm(inv3(Km1.T.copy()), tx), R12), inv3(Km2))
    return ep, F.reshape(9)


def fill_geometry(case: TriangulationCase) -> TriangulationCase:
    P = case.kf1.shape[0]
    ep = np.zeros((P, 2), dtype=np.float32)
    f12 = np.zeros((P, 9), dtype=np.float32)
    for p in range(P):
        ep[p], f12[p] = triangulation_geometry_numpy(case.T1w[p], case.T2w[p], case.K, case.K)
    case.ep, case.f12 = ep, f12
    return case


# ---------------------------------------------------------------- C5
@dataclass
class KnnCase:
    q: np.ndarray
    db: np.ndarray
    th_low: int = 50
    nnratio: float = 0.8


def make_knn_case(seed: int, nq: int, nd: int, planted_frac: float = 0.1) -> KnnCase:
    """C5: nq queries vs nd database descriptors; planted_frac of the queries are bit-flipped copies of
    database rows; a few database rows are exact duplicates (first-index tie-break) and a few queries
    have two near-equal neighbours (ratio-test failures)."""
    rng = np.random.default_rng(seed)
    db = random_descriptors(rng, nd)
    q = random_descriptors(rng, nq)
    n_pl = max(1, int(planted_frac * nq))
    who = rng.permutation(nq)[:n_pl]
    tgt = rng.integers(0, nd, n_pl)
    q[who] = planted_copies(rng, db[tgt])
    n_dup = max(1, nd // 1000)
    a = rng.integers(0, nd, n_dup)
    b = rng.integers(0, nd, n_dup)
    db[b] = db[a]
    # re-plant after duplication so that planted rows still exist; ties on duplicates are intended
    return KnnCase(q, db)


# ---------------------------------------------------------------- 8(f) rank 1: Frame::isInFrustum / Tracking::SearchLocalPoints
@dataclass
class FrustumCase:
    frame: HostFrame
    Tcw34: np.ndarray       # [3][4] row-major pose of the frame (the stand-in Sophus of the oracle build decomposes it)
    Rcw: np.ndarray
    tcw: np.ndarray
    Ow: np.ndarray          # -Rcw^T tcw evaluated in fp32 exactly as the stand-in SE3f::inverse() does
    K: np.ndarray
    mbf: float
    viewing_cos_limit: float
    log_scale_factor: float
    world_pos: np.ndarray
    normal: np.ndarray
    min_distance: np.ndarray
    max_distance: np.ndarray
    desc: np.ndarray
    bad: np.ndarray
    n_obs: np.ndarray
    skip: np.ndarray
    kp_prior_obs: np.ndarray
    kp_mp: np.ndarray
    nnratio: float = 0.8


def _f32_matvec_t(R, t):
    """-(R^T t) in fp32, left to right, as the oracle build's matrix-form SE3f::inverse() evaluates it"""
    Rt = R.T.astype(np.float32)
    out = np.empty(3, dtype=np.float32)
    for r in range(3):
        acc = np.float32(Rt[r, 0] * t[0])
        acc = np.float32(acc + np.float32(Rt[r, 1] * t[1]))
        acc = np.float32(acc + np.float32(Rt[r, 2] * t[2]))
        out[r] = -acc
    return out


def make_frustum_case(seed: int, n_kp: int = 2000, n_mp: int = 5000, planted_frac: float = 0.35) -> FrustumCase:
    """Local map points around a camera: a third project onto keypoints of the frame (planted descriptors, distance chosen so that
    PredictScale lands on the keypoint's octave), the rest are spread in front of, behind and beside the camera so that every gate of
    Frame::isInFrustum (depth sign, image bounds, distance invariance, viewing cosine) rejects some of them.  Points exactly on the
    image border and on the distance limits are included."""
    rng = np.random.default_rng(seed)
    frame = make_frame(rng, n_kp)
    sf = frame.scale_factors
    R = _small_rotation(rng, 1)[0].astype(np.float32)
    t = rng.normal(0, 0.3, 3).astype(np.float32)
    Ow = _f32_matvec_t(R, t)
    K = np.array([FX, FY, CX, CY], dtype=np.float32)
    # random points in camera coordinates -> world
    Pc = np.stack([rng.uniform(-6, 6, n_mp), rng.uniform(-5, 5, n_mp), rng.uniform(-2, 14, n_mp)], axis=1)
    n_pl = int(planted_frac * n_mp)
    kp = rng.integers(0, n_kp, n_pl)
    z = rng.uniform(1.5, 12.0, n_pl)
    u = frame.kp_xy[kp, 0] + rng.normal(0, 1.0, n_pl)
    v = frame.kp_xy[kp, 1] + rng.normal(0, 1.0, n_pl)
    Pc[:n_pl] = np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], axis=1)
    Pw = (Pc - t.astype(np.float64)) @ R.astype(np.float64)  # R^T (Pc - t)
    world = Pw.astype(np.float32)
    dist = np.linalg.norm(world.astype(np.float64) - Ow.astype(np.float64), axis=1)
    # distance limits: the planted points get max_distance = dist * sf[octave] (so PredictScale = octave); others random
    lvl = rng.integers(0, 8, n_mp)
    lvl[:n_pl] = frame.octave[kp]
    max_d = (dist * sf[lvl] * rng.uniform(0.98, 1.02, n_mp)).astype(np.float32)
    min_d = (max_d / (1.2 ** 7) * rng.uniform(0.5, 1.4, n_mp)).astype(np.float32)
    far = rng.random(n_mp) < 0.1
    max_d[far] = (dist[far] * rng.uniform(0.5, 0.9, far.sum())).astype(np.float32)  # beyond 1.2 x max for some of these
    normal = (world.astype(np.float64) - Ow.astype(np.float64)) / np.maximum(dist, 1e-6)[:, None]
    normal += rng.normal(0, 0.5, (n_mp, 3))  # some viewing angles beyond the limit
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    desc = random_descriptors(rng, n_mp)
    desc[:n_pl] = planted_copies(rng, frame.desc[kp])
    bad = (rng.random(n_mp) < 0.03).astype(np.uint8)
    n_obs = rng.integers(1, 20, n_mp).astype(np.int32)
    skip = ((rng.random(n_mp) < 0.08) | (bad > 0)).astype(np.uint8)
    prior = np.zeros(n_kp, dtype=np.int32)
    kp_mp = np.full(n_kp, -1, dtype=np.int32)
    held = rng.permutation(n_kp)[: n_kp // 10]
    prior[held] = rng.integers(0, 4, held.size)
    kp_mp[held] = rng.integers(0, n_mp, held.size)
    T = np.concatenate([R, t[:, None]], axis=1).astype(np.float32)
    return FrustumCase(frame, T, R, t, Ow, K, 40.0, 0.5, float(np.log(np.float32(1.2))), world, normal.astype(np.float32), min_d, max_d, desc,
                       bad, n_obs, skip, prior, kp_mp)
