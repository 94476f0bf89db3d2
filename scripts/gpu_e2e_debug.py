"""development aid: where does the C5 end-to-end step spend its time under torchrun (per rank)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from orb_slam3_comments_ghr_b200 import matcher
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = matcher.Context(lr, stream=torch.cuda.current_stream().cuda_stream)
nq, nd = 262144 // world, 4194304
db_h = torch.randint(0, 256, (nd, 32), dtype=torch.uint8).pin_memory().numpy()
q_h = torch.randint(0, 256, (nq, 32), dtype=torch.uint8).pin_memory().numpy()
m = matcher.ORBmatcher(0.8, True, ctx)
for it in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); hdb = ctx.upload_database(db_h); t1 = time.perf_counter()
    r = m.SearchByNN(hdb, q_h, 50); t2 = time.perf_counter()
    del hdb; torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"rank {rank} it {it}: upload {1e3*(t1-t0):.1f} ms  search {1e3*(t2-t1):.1f} ms  free {1e3*(t3-t2):.1f} ms", flush=True)
if world > 1: dist.destroy_process_group()
