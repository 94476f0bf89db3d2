"""Multi-GPU check of the fused search + all-gather in the vMatchedPairs form (run under torchrun on N GPUs of one box): every rank
holds the same key-frame set, owns a shard of the pair list, and after a step must hold -- for ALL pairs -- exactly the compact
result it computes itself for the whole batch on its own GPU (n_ranks = 1 gather) and the oracle's rows for a sample.  Several steps,
both buffers, with a deliberately slow rank so that the epoch protocol (a rank may be one step ahead) is exercised."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200.sharding import TriangulationGather, compact_pairs_from_rows, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
P, NF = 64 * world, 1200
case = synth.fill_geometry(synth.make_triangulation_case(77, n_pairs=P, n_feat=NF))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
lo, hi = shard_bounds(P, rank, world)
# reference on this GPU: the whole batch through a one-rank gather
solo = TriangulationGather(matcher, case.kfs, P, NF, 0, 1, dev, 0.6, False, use_graph=False)
solo.set_inputs(t(case.kf1), t(case.kf2), t(case.ep), t(case.f12))
c0, e0 = solo.step()
torch.cuda.synchronize()
c0, e0 = c0.cpu().numpy(), e0.cpu().numpy().view(np.uint32)
ok = True
for graph in (False, True):
    tg = TriangulationGather(matcher, case.kfs, P, NF, rank, world, dev, 0.6, False, use_graph=graph)
    tg.set_inputs(t(case.kf1[lo:hi]), t(case.kf2[lo:hi]), t(case.ep[lo:hi]), t(case.f12[lo:hi]))
    for step in range(6):
        if rank == (step % world):
            time.sleep(0.02)  # one rank lags: the others wait in the kernel for its epoch
        c, e = tg.step()
        torch.cuda.synchronize()
        c, e = c.cpu().numpy(), e.cpu().numpy().view(np.uint32)
        good = np.array_equal(c, c0) and all(np.array_equal(e[p, :c[p]], e0[p, :c0[p]]) for p in range(P)) and tg.status() == 0
        ok = ok and good
        if not good:
            print(f"rank {rank} graph={graph} step {step}: MISMATCH (counts equal: {np.array_equal(c, c0)})", flush=True)
    offs, pairs = tg.download(*tg.step())
    ok = ok and np.array_equal(offs, np.concatenate([[0], np.cumsum(c0)]).astype(np.int32)) and pairs.shape[0] == int(c0.sum())
    offs, pairs = np.array(offs), np.array(pairs)
    o_own, p_own = tg.download(*tg.step(), own_only=True)  # this rank's pairs only == its slice of the full download
    ok = ok and np.array_equal(np.asarray(o_own), offs[lo:hi + 1] - offs[lo]) and np.array_equal(np.asarray(p_own), pairs[offs[lo]:offs[hi]])
    del tg
if rank == 0:  # and the oracle, for a sample of pairs
    from oracle.pyoracle import Oracle
    nm, m = Oracle().search_for_triangulation_batch(case.kfs, case.kf1[:16], case.kf2[:16], case.ep[:16], case.f12[:16], 0, 0, 0, n_threads=8)
    cc, ee = compact_pairs_from_rows(m)
    ok = ok and np.array_equal(cc, c0[:16]) and all(np.array_equal(ee[p, :cc[p]], e0[p, :cc[p]]) for p in range(16))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"GATHER CHECK {'OK' if int(flag.item()) else 'FAILED'}: {world} ranks, {P} pairs x {NF} features, {int(c0.sum())} matches, 14 steps + 4 downloads per rank")
dist.barrier()
dist.destroy_process_group()
