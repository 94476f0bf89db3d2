"""multi-GPU check (torchrun): the fused search + all-gather over peer memory (torch symmetric memory, NVLink stores from inside
the kernel) assembles on EVERY rank exactly the rows a single rank computes for all pairs; also times it against the NCCL
all-gather of the same rows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from orb_slam3_comments_ghr_b200 import matcher, synth
from orb_slam3_comments_ghr_b200.sharding import all_gather_rows, shard_bounds
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
P_total, NF = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 2000
ctx = matcher.Context(lr, stream=torch.cuda.current_stream().cuda_stream)
tc = synth.fill_geometry(synth.make_triangulation_case(20261018, n_pairs=P_total, n_feat=NF))   # same case on every rank
ks = ctx.upload_kfset(tc.kfs)
m = matcher.ORBmatcher(0.6, False, ctx)
lo, hi = shard_bounds(P_total, rank, world); P = hi - lo
kf1, kf2 = torch.from_numpy(tc.kf1).to(dev), torch.from_numpy(tc.kf2).to(dev)
ep, f12 = torch.from_numpy(tc.ep).to(dev), torch.from_numpy(tc.f12).to(dev)
# reference: every rank computes ALL pairs locally
full = torch.empty((P_total, NF), dtype=torch.int32, device=dev); fnm = torch.empty(P_total, dtype=torch.int32, device=dev)
m.SearchForTriangulation_dev(ks, P_total, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(), full.data_ptr(), fnm.data_ptr())
# symmetric result buffers, double-buffered: [2][P_total*NF] rows + [2][P_total] counts
rows = symm_mem.empty((2, P_total * NF), dtype=torch.int32, device=dev)
cnts = symm_mem.empty((2, P_total), dtype=torch.int32, device=dev)
hr = symm_mem.rendezvous(rows, dist.group.WORLD.group_name)
hc = symm_mem.rendezvous(cnts, dist.group.WORLD.group_name)
rows.fill_(-1); cnts.fill_(0)
hr.barrier(channel=0)
MODE = int(os.environ.get("PEER_MODE", "2"))   # 2: whole rows copied to the peers (default); 1: preset + individual match stores
order = [rank] + [r for r in range(world) if r != rank]  # target 0 = this rank's own buffer
def targets(b):
    return [hr.buffer_ptrs[r] + b * P_total * NF * 4 for r in order], [hc.buffer_ptrs[r] + b * P_total * 4 for r in order]
def step(k, mm=None):
    b = k & 1
    tm, tn = targets(b)
    (mm or m).SearchForTriangulation_peers_dev(ks, P, kf1[lo:hi].data_ptr(), kf2[lo:hi].data_ptr(), ep[lo:hi].data_ptr(), f12[lo:hi].data_ptr(), tm, tn, lo, MODE)
    if MODE == 1:
        rows[b ^ 1].fill_(-1)      # my copy of the NEXT step's buffer; every rank has done this when the barrier releases
    hr.barrier(channel=0)
    return rows[b].view(P_total, NF), cnts[b]
ok = True
for k in range(4):
    r, c = step(k)
    torch.cuda.synchronize()
    same = bool(torch.equal(r, full)) and bool(torch.equal(c, fnm))
    ok = ok and same
    if not same:
        print(f"rank {rank} step {k}: MISMATCH rows {int((r != full).sum())} counts {int((c != fnm).sum())}", flush=True)
# timing: fused vs NCCL all-gather of the dense rows
def timeit(fn, n=20):
    for i in range(4): fn(i)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
out = torch.empty((max(P, 1), NF), dtype=torch.int32, device=dev); nmt = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
gathered = torch.empty((P_total, NF), dtype=torch.int32, device=dev)
def nccl_step(i):
    m.SearchForTriangulation_dev(ks, P, kf1[lo:hi].data_ptr(), kf2[lo:hi].data_ptr(), ep[lo:hi].data_ptr(), f12[lo:hi].data_ptr(), out.data_ptr(), nmt.data_ptr())
    return all_gather_rows(out, P_total, out=gathered)
t_fused, t_nccl = timeit(step), timeit(nccl_step)
# the same two steps captured in CUDA graphs (two steps per graph: both halves of the double buffer), so that the host-side
# launch path (three Python / ctypes calls per step) does not bound a step that is only tens of microseconds of GPU work
t_fused_g = t_nccl_g = float("nan")
try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gctx = matcher.Context(lr, stream=side.cuda_stream)
        gm = matcher.ORBmatcher(0.6, False, gctx)
        def gstep(k):
            step(k, gm)
        gstep(0); gstep(1)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            gstep(0); gstep(1)
    torch.cuda.current_stream().wait_stream(side)
    t_fused_g = timeit(lambda i: g.replay(), n=20) / 2
    torch.cuda.synchronize()
    okg = bool(torch.equal(rows[1].view(P_total, NF), full)) and bool(torch.equal(cnts[1], fnm))
    ok = ok and okg
except Exception as e:
    print(f"rank {rank}: graph capture failed: {e!r}", flush=True)
if rank == 0:
    print(f"world {world} mode {MODE}: fused peer-store all-gather {'OK' if ok else 'FAILED'}; step {t_fused*1e3:.1f} us fused vs {t_nccl*1e3:.1f} us kernel + NCCL all-gather "
          f"({P_total / t_fused / 1e3:.2f} vs {P_total / t_nccl / 1e3:.2f} M pairs/s); fused in a CUDA graph {t_fused_g*1e3:.1f} us "
          f"({P_total / t_fused_g / 1e3:.2f} M pairs/s)", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
