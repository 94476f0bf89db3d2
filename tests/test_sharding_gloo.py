"""CPU, world_size 2 over gloo: the N>1 path (contiguous query shards + one all-gather of the match
indices) assembles exactly the single-rank result.  Per-rank compute is the oracle here (the GPU
kernels are covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT
from orb_slam3_comments_ghr_b200 import synth
from orb_slam3_comments_ghr_b200.sharding import all_gather_rows, compact_pairs_from_rows, max_shard, pairs_from_compact, shard_bounds


def test_shard_bounds_cover():
    for n in (0, 1, 7, 8, 4096, 262144, 1000003):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) == max_shard(n, w) or n == 0


def _worker(rank, world, port, nq, nd, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    o = Oracle()
    kc = synth.make_knn_case(61, nq, nd)
    lo, hi = shard_bounds(nq, rank, world)
    bi, bd, sd, mt = o.knn2_ratio(kc.q[lo:hi], kc.db, 50, 0.8, 1)
    local = torch.from_numpy(np.stack([bi, bd, sd, mt], axis=1))
    full = all_gather_rows(local, nq)
    # C4-style 2-D payload (pairs x features)
    tc = synth.fill_geometry(synth.make_triangulation_case(62, n_pairs=5, n_feat=256))
    plo, phi = shard_bounds(5, rank, world)
    nm, m = o.search_for_triangulation_batch(tc.kfs, tc.kf1[plo:phi], tc.kf2[plo:phi], tc.ep[plo:phi], tc.f12[plo:phi])
    fullm = all_gather_rows(torch.from_numpy(m), 5)
    # C4 result form of the fused GPU all-gather (vMatchedPairs: counts + (idx1 << 16 | idx2) entries): shard, compact, gather
    tc6 = synth.fill_geometry(synth.make_triangulation_case(63, n_pairs=6, n_feat=256))
    clo, chi = shard_bounds(6, rank, world)
    _, m6 = o.search_for_triangulation_batch(tc6.kfs, tc6.kf1[clo:chi], tc6.kf2[clo:chi], tc6.ep[clo:chi], tc6.f12[clo:chi])
    c6, e6 = compact_pairs_from_rows(m6)
    gc = all_gather_rows(torch.from_numpy(c6), 6)
    ge = all_gather_rows(torch.from_numpy(e6.view(np.int32)), 6)
    # equal shards: the direct path into a caller-owned buffer
    elo, ehi = shard_bounds(300, rank, world)
    buf = torch.empty((300, 4), dtype=local.dtype)
    eq = all_gather_rows(torch.from_numpy(np.stack(o.knn2_ratio(kc.q[elo:ehi], kc.db, 50, 0.8, 1), axis=1)), 300, out=buf)
    if rank == 0:
        ret["knn"] = full.numpy()
        ret["tri"] = fullm.numpy()
        ret["knn_eq"] = eq.numpy()
        ret["tri_counts"], ret["tri_entries"] = gc.numpy(), ge.numpy().view(np.uint32)
        ret["eq_in_place"] = eq.data_ptr() == buf.data_ptr()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_allgather_matches_single_rank(oracle):
    nq, nd = 301, 4000
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, nq, nd, ret), nprocs=2, join=True)
    kc = synth.make_knn_case(61, nq, nd)
    exp = np.stack(oracle.knn2_ratio(kc.q, kc.db, 50, 0.8, 1), axis=1)
    assert np.array_equal(ret["knn"], exp)
    assert np.array_equal(ret["knn_eq"], exp[:300]) and ret["eq_in_place"]
    tc = synth.fill_geometry(synth.make_triangulation_case(62, n_pairs=5, n_feat=256))
    nm, m = oracle.search_for_triangulation_batch(tc.kfs, tc.kf1, tc.kf2, tc.ep, tc.f12)
    assert np.array_equal(ret["tri"], m)
    tc6 = synth.fill_geometry(synth.make_triangulation_case(63, n_pairs=6, n_feat=256))
    nm6, m6 = oracle.search_for_triangulation_batch(tc6.kfs, tc6.kf1, tc6.kf2, tc6.ep, tc6.f12)
    assert np.array_equal(ret["tri_counts"], nm6) and int(nm6.sum()) > 0
    for p in range(6):  # vMatchedPairs of every pair, ascending idx1 (ORBmatcher.cc:1317-1325)
        i1 = np.flatnonzero(m6[p] >= 0)
        assert np.array_equal(pairs_from_compact(ret["tri_counts"], ret["tri_entries"], p), np.stack([i1, m6[p, i1]], axis=1))
