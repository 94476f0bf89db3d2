"""Query sharding for the batched workloads: frame-pair batches (C4) and kNN
queries (C5) are split contiguously by row over the ranks of one box; each rank works on its rows
with no data-path collective, and ONE all-gather of the match indices assembles the result.
The database / keyframe set is replicated.  Works with NCCL (GPU tensors) and gloo (CPU tests)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """rows [lo, hi) of rank: contiguous, sizes differ by at most one, earlier ranks get the extra row"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


def all_gather_rows(local: torch.Tensor, n_total: int, out: torch.Tensor = None) -> torch.Tensor:
    """local: [rows_of_this_rank, ...] -> [n_total, ...] on every rank with a single all-gather.
    Equal shards (n_total divisible by the world size) gather straight into `out` (a caller-owned [n_total, ...]
    buffer, allocated when omitted) with no staging copy; ragged shards are padded to the largest shard and trimmed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max_shard(n_total, world)
    if n_total % world == 0:
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * cap: r * cap + (hi - lo)])
    return torch.cat(parts, dim=0)


# ------------------------------------------------------------------------------------------------------------------
# C4: the batched SearchForTriangulation result in the reference's own form, vMatchedPairs (ORBmatcher.cc:1317-1325)
def compact_pairs_from_rows(matches12: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """dense rows vMatches12[P, n_feat] (-1 none) -> (counts[P] int32, entries[P, n_feat] uint32): the matches of pair p in
    ascending idx1 as (idx1 << 16 | idx2), entries beyond counts[p] are 0xFFFFFFFF -- the layout the fused all-gather
    (orbgpu_search_for_triangulation_batch_gather_dev) leaves in every rank's buffers."""
    m = np.asarray(matches12)
    P, n = m.shape
    counts = (m >= 0).sum(axis=1).astype(np.int32)
    ent = np.full((P, n), 0xFFFFFFFF, dtype=np.uint32)
    for p in range(P):
        i1 = np.flatnonzero(m[p] >= 0)
        ent[p, : i1.size] = (i1.astype(np.uint32) << 16) | m[p, i1].astype(np.uint32)
    return counts, ent


def pairs_from_compact(counts: np.ndarray, entries: np.ndarray, p: int) -> np.ndarray:
    """vMatchedPairs of pair p as [(idx1, idx2), ...] from the compact form"""
    e = np.asarray(entries[p, : int(counts[p])], dtype=np.uint32)
    return np.stack([(e >> 16).astype(np.int32), (e & 0xFFFF).astype(np.int32)], axis=1)


class TriangulationGather:
    """Pairs sharded by index over the ranks of one box; every rank ends up with the vMatchedPairs of ALL pairs.  The search
    kernel stores the compact entries of each finished pair into the buffers of all ranks over NVLink peer memory (torch
    symmetric memory) and the ranks meet through epoch flags -- no NCCL collective and no barrier in the step.  Steps alternate
    between two buffers (a rank may run one step ahead of its peers) and are captured in CUDA graphs."""

    def __init__(self, matcher_mod, kfset_host, p_total: int, n_feat: int, rank: int, world: int, device: torch.device,
                 nnratio: float = 0.6, check_ori: bool = False, use_graph: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.M, self.rank, self.world, self.dev = matcher_mod, rank, world, device
        self.p_total, self.n_feat = p_total, n_feat
        assert p_total % world == 0 and n_feat % 4 == 0
        self.lo, self.hi = shard_bounds(p_total, rank, world)
        # graph replays need a capturable (non-default) stream; the direct launch goes on the caller's current stream, with no
        # cross-stream event hops around the one kernel of the step
        self.side = torch.cuda.Stream(device=device) if use_graph else None
        self.ctx = matcher_mod.Context(device.index, stream=(self.side or torch.cuda.current_stream(device)).cuda_stream)
        self.m = matcher_mod.ORBmatcher(nnratio, check_ori, self.ctx)
        self.ks = self.ctx.upload_kfset(kfset_host)
        if world > 1:
            alloc = lambda shape, dt: symm_mem.empty(shape, dtype=dt, device=device)  # noqa: E731
        else:
            alloc = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)  # noqa: E731
        self.pairs = alloc((2, p_total * n_feat), torch.int32)
        self.counts = alloc((2, p_total), torch.int32)
        self.flags = alloc((64,), torch.int32)
        self.state = torch.zeros(8, dtype=torch.int32, device=device)  # [0] epoch, [1] shipped pairs, [4] status
        self.state[0] = 1
        self.flags.zero_()
        self.counts.zero_()
        if world > 1:
            name = dist.group.WORLD.group_name
            hp, hc, hf = symm_mem.rendezvous(self.pairs, name), symm_mem.rendezvous(self.counts, name), symm_mem.rendezvous(self.flags, name)
            self._h = (hp, hc, hf)
            pp, cp, fp = hp.buffer_ptrs, hc.buffer_ptrs, hf.buffer_ptrs
            hp.barrier(channel=0)
        else:
            pp, cp, fp = [self.pairs.data_ptr()], [self.counts.data_ptr()], [self.flags.data_ptr()]
        torch.cuda.synchronize(device)
        self.g = [matcher_mod.tri_gather_struct(rank, [p + b * p_total * n_feat * 4 for p in pp], [c + b * p_total * 4 for c in cp], fp,
                                                self.state.data_ptr(), self.state.data_ptr() + 16) for b in (0, 1)]
        self.k = 0
        self.graphs = None
        self.inputs = None
        self.use_graph = use_graph

    def set_inputs(self, kf1: torch.Tensor, kf2: torch.Tensor, ep: torch.Tensor, f12: torch.Tensor):
        """device tensors of THIS rank's pairs (fixed addresses: the graphs replay on them)"""
        self.inputs = (kf1, kf2, ep, f12)
        self.graphs = None

    def _launch(self, b: int):
        kf1, kf2, ep, f12 = self.inputs
        self.m.SearchForTriangulation_gather_dev(self.ks, self.hi - self.lo, kf1.data_ptr(), kf2.data_ptr(), ep.data_ptr(), f12.data_ptr(),
                                                 self.g[b], self.lo)

    def _capture(self):
        cur = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            self._launch(0)
            self._launch(1)  # warm-up outside capture: both buffers, an even number of steps on every rank
            self.side.synchronize()
            self.k += 2
            graphs = []
            for b in (0, 1):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=self.side):
                    self._launch(b)
                graphs.append(g_)
        cur.wait_stream(self.side)
        self.graphs = graphs

    def step(self):
        """one sharded search + all-gather on the current stream; returns (counts[P_total], entries[P_total, n_feat]) views of
        the buffer this step filled (valid after the stream reaches this point)"""
        b = self.k & 1
        if self.use_graph and self.graphs is None:
            self._capture()
            b = self.k & 1
        self.k += 1
        if self.graphs is not None:
            self.graphs[b].replay()
        elif self.side is None:
            self._launch(b)
        else:
            cur = torch.cuda.current_stream(self.dev)
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                self._launch(b)
            cur.wait_stream(self.side)
        return self.counts[b], self.pairs[b].view(self.p_total, self.n_feat)

    def download(self, counts: torch.Tensor, entries: torch.Tensor, out=None, own_only: bool = False):
        """the gathered result of a step on the host: (pair_offsets[P_total + 1], pairs[total, 2]) -- vMatchedPairs of ALL pairs;
        own_only: only this rank's pairs [lo, hi) (the host work that follows -- triangulating the matches -- shards the same way)"""
        if own_only:
            if self.side is not None:
                self.side.wait_stream(torch.cuda.current_stream(self.dev))
            return self.m.TriangulationGatherDownload(self.hi - self.lo, self.n_feat, counts.data_ptr() + 4 * self.lo,
                                                      entries.data_ptr() + 4 * self.lo * self.n_feat, out=out)
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(self.dev))  # the step was replayed on the current stream; the download runs on the context's
        offs, pairs = self.m.TriangulationGatherDownload(self.p_total, self.n_feat, counts.data_ptr(), entries.data_ptr(), out=out)
        return offs, pairs

    def status(self) -> int:
        """non-zero when a peer did not arrive within the time-out of the wait kernel"""
        return int(self.state[4].item())
