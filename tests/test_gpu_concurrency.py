"""Reentrancy of the boundary (run with -m gpu): the reference calls ORBmatcher from three threads at once -- Tracking,
LocalMapping and LoopClosing (System.cc:234,254) -- so three host threads, each with its own context (stream + workspace),
hammer DIFFERENT searches concurrently and every result must equal the single-threaded one.  Kernel attributes are per-device
state set once at orbgpu_create; frame slabs are recycled through events, not device-wide synchronisations."""
import threading

import numpy as np
import pytest

from helpers import attach_featvec, golden_voc
from orb_slam3_comments_ghr_b200 import synth

pytestmark = pytest.mark.gpu


def _workloads(M, oracle, device):
    """three per-thread jobs: (name, callable(ctx) -> tuple of arrays)"""
    ci = synth.make_init_case(901, n=3000)
    cp = synth.make_projection_case(902, n_kp=14000, n_mp=6000, th=3.0)  # lock table beyond the 48 KB default shared memory
    voc = golden_voc()
    cb = synth.make_bow_case(903, voc)
    kf_h, f_h = attach_featvec(oracle, voc, cb.kf, 2), attach_featvec(oracle, voc, cb.f, 2)
    kc = synth.make_knn_case(904, 1500, 40000)
    tc = synth.fill_geometry(synth.make_triangulation_case(905, n_pairs=48, n_feat=1500))

    def tracking(ctx):  # SearchForInitialization + SearchByProjection: frames uploaded and destroyed on every call
        m = M.ORBmatcher(ci.nnratio, True, ctx)
        a = m.SearchForInitialization(ctx.upload_frame(ci.f1), ctx.upload_frame(ci.f2), ci.prev_matched, ci.window_size)
        b = M.ORBmatcher(cp.nnratio, True, ctx).SearchByProjection(ctx.upload_frame(cp.frame), cp.mps, 3.0, False, 50.0, cp.kp_prior_obs, cp.kp_mp)
        return (np.int64(a[0]), a[1], a[2], np.int64(b[0]), b[1])

    def local_mapping(ctx):  # batched SearchForTriangulation (persistent pipeline kernel, ~200 KB dynamic shared memory)
        ks = ctx.upload_kfset(tc.kfs)
        nm, mt = M.ORBmatcher(0.6, False, ctx).SearchForTriangulation(ks, tc.kf1, tc.kf2, tc.ep, tc.f12)
        return (nm, mt)

    def loop_closing(ctx):  # transform + SearchByBoW + brute-force 2-NN on the tcgen05 engine (197 KB dynamic shared memory)
        dkf, df = ctx.upload_frame(kf_h), ctx.upload_frame(f_h)
        n, mt = M.ORBmatcher(cb.nnratio, True, ctx).SearchByBoW(dkf, df, cb.kf_mp_valid)
        w = ctx.upload_frame(cb.f).transform(ctx.upload_vocabulary(voc), 2)
        k = M.ORBmatcher(kc.nnratio, True, ctx).SearchByNN(ctx.upload_database(kc.db), kc.q, kc.th_low)
        return (np.int64(n), mt) + tuple(w) + tuple(k)

    return [("tracking", tracking), ("local_mapping", local_mapping), ("loop_closing", loop_closing)]


def _hammer(M, jobs, device, rounds):
    expected = []
    for _, fn in jobs:  # single-threaded results first
        ctx = M.Context(device)
        expected.append(fn(ctx))
        ctx.close()
    errors = []
    start = threading.Barrier(len(jobs))

    def run(i):
        try:
            ctx = M.Context(device)
            start.wait()
            for r in range(rounds):
                got = jobs[i][1](ctx)
                for a, b in zip(got, expected[i]):
                    if not np.array_equal(a, b):
                        errors.append(f"{jobs[i][0]} round {r}: result differs from the single-threaded run")
                        return
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append(f"{jobs[i][0]}: {e!r}")

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_three_threads_concurrent(oracle):
    from orb_slam3_comments_ghr_b200 import matcher as M
    _hammer(M, _workloads(M, oracle, 0), 0, rounds=12)


def test_second_device_in_one_process(oracle):
    """per-device kernel attributes: a context on a second GPU of the same process must get its own shared-memory opt-in"""
    from orb_slam3_comments_ghr_b200 import matcher as M
    n = M.load_library().orbgpu_device_count()
    if n < 2:
        pytest.skip("needs two GPUs in one process")
    jobs = _workloads(M, oracle, 1)
    _hammer(M, jobs, 1, rounds=2)
    _hammer(M, jobs, 0, rounds=1)
