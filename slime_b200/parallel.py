"""Data-parallel plumbing for the SliME prefill path: samples are independent, so the batch is
split into contiguous per-rank shards, weights are replicated, and the ONLY exchange step is the gather of the
last-token logits.  torch.distributed does the plumbing (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the batch owned by `rank`; block sizes differ by at most one."""
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_order(costs: Sequence[int], world: int) -> List[int]:
    """Sample order that evens out per-rank work when crop counts differ (the reference's training-side
    analogue is group_by_modality_length, llava/train/train.py:131): sort by cost descending and deal the
    samples round-robin, then lay the ranks' lists out contiguously so shard_bounds() picks them up."""
    idx = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    per_rank: List[List[int]] = [[] for _ in range(world)]
    sizes = [shard_bounds(len(costs), r, world) for r in range(world)]
    cap = [hi - lo for lo, hi in sizes]
    r = 0
    for i in idx:
        while len(per_rank[r]) >= cap[r]:
            r = (r + 1) % world
        per_rank[r].append(i)
        r = (r + 1) % world
    return [i for lst in per_rank for i in sorted(lst)]


def gather_logits(local_logits: torch.Tensor, n_samples: int, group=None) -> torch.Tensor:
    """All-gather [B_local, V] logits into [n_samples, V] in global sample order.  Ranks may own blocks that
    differ by one row: shorter blocks are zero-padded for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_logits
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_bounds(n_samples, r, world) for r in range(world))]
    width = max(sizes)
    assert local_logits.shape[0] == sizes[rank], (local_logits.shape, sizes, rank)
    buf = local_logits
    if sizes[rank] < width:
        buf = torch.cat([local_logits, local_logits.new_zeros(width - sizes[rank], local_logits.shape[1])])
    out = local_logits.new_empty(world * width, local_logits.shape[1])
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(s == width for s in sizes):
        return out
    return torch.cat([out[r * width: r * width + sizes[r]] for r in range(world)])


class SlimeComm:
    """NCCL communicator owned by libslime_b200 (include/slime_b200.h: slime_comm_*; NCCL bound with dlopen).  The 128-byte
    unique id is created on rank 0 and reaches the other ranks through the torch.distributed store (bootstrap plumbing:
    `dist.broadcast_object_list`, which works on any backend)."""

    def __init__(self, device, dtype=None, group=None):
        import ctypes as C

        from . import _lib as L

        self.lib = L.load(dtype)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_char * 128)()
        if self.rank == 0:
            L.check(self.lib.slime_comm_unique_id(buf), "comm_unique_id", self.lib)
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            L.check(self.lib.slime_comm_init(C.byref(self._h), ident, self.rank, self.world), "comm_init", self.lib)

    def allgather_logits(self, local: torch.Tensor, out: torch.Tensor, stream: "torch.cuda.Stream") -> None:
        import ctypes as C

        from . import _lib as L

        assert local.dtype == torch.float32 and out.dtype == torch.float32 and local.is_contiguous() and out.is_contiguous()
        assert out.shape[0] == self.world * local.shape[0] and out.shape[1] == local.shape[1]
        L.check(self.lib.slime_allgather_logits(self._h, L.ptr(local), L.ptr(out), local.shape[0], local.shape[1],
                                                C.c_void_p(stream.cuda_stream)), "allgather_logits", self.lib)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.slime_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LogitsGather:
    """The path's one exchange step, taken OFF the compute stream: the all-gather of the [B_local, V] last-token logits
    is issued on a side stream behind an event, so a rank's next prefill never waits for the slowest rank's previous
    one (on the compute stream the collective acts as a per-step barrier: every step then runs at the pace of the most
    power-throttled GPU).  `depth` result buffers rotate; result(slot) makes the calling stream wait for that gather.
    Equal block sizes on every rank (the benchmark's weak-scaling layout); gather_logits() handles ragged blocks.
    On CUDA the collective is the library's own entry point (slime_allgather_logits over its NCCL communicator, SlimeComm);
    `use_lib=False` falls back to torch.distributed's all_gather_into_tensor (same NCCL underneath).
    On CPU tensors (gloo tests) the gather simply runs synchronously."""

    def __init__(self, rows: int, vocab: int, device, dtype=torch.float32, group=None, depth: int = 2, use_lib: bool = True):
        self.group = group
        self.comm = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.empty(self.world * rows, vocab, dtype=dtype, device=device) for _ in range(depth)]
        self.done = [None] * depth
        self.count = 0
        self.stream = torch.cuda.Stream(device) if self.cuda else None
        if self.cuda and self.world > 1 and use_lib and dtype == torch.float32:
            self.comm = SlimeComm(device, group=group)

    def submit(self, local_logits: torch.Tensor) -> int:
        slot = self.count % len(self.bufs)
        self.count += 1
        if self.world == 1:
            self.bufs[slot].copy_(local_logits)
            return slot
        if not self.cuda:
            dist.all_gather_into_tensor(self.bufs[slot], local_logits.contiguous(), group=self.group)
            return slot
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        local_logits.record_stream(self.stream)  # allocated on the compute stream, consumed on the side stream
        with torch.cuda.stream(self.stream):
            if self.comm is not None:
                self.comm.allgather_logits(local_logits.contiguous(), self.bufs[slot], self.stream)
            else:
                dist.all_gather_into_tensor(self.bufs[slot], local_logits, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[slot] = ev
        return slot

    def result(self, slot: int) -> torch.Tensor:
        """The gathered [world * rows, V] logits of submission `slot`; the current stream waits for that gather."""
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.done[slot])
        return self.bufs[slot]

    def drain(self) -> None:
        """Make the current stream wait for every gather issued so far (end of a timed region)."""
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.stream)
