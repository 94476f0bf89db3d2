"""probe: torch symmetric memory (peer pointers over NVLink) under torchrun"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1 << 20,), dtype=torch.int32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    t.fill_(rank + 1)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.int32)
    print(rank, "peer value", int(peer[5].item()), flush=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100): hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print(rank, "barrier us", (time.perf_counter() - t0) * 1e4, flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
