"""Row-range sharding of one batch across the GPUs of a box.

Rows are independent, so a Projector needs no exchange at all and a Filter needs exactly one:
reassembling the SelectionVector.  Each rank filters its contiguous row range and emits GLOBAL
row numbers (`index_base` = first row of the range); because ranges are contiguous and each
run is ascending, concatenating the runs in rank order gives the global ascending vector.
`torch.distributed` is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_rows: int, world: int, rank: int, align: int = 64) -> Tuple[int, int]:
    """[first, last) rows of `rank`.  Range starts are multiples of `align` rows so that no two
    shards share a validity-bitmap word and value loads stay 128-byte aligned."""
    per = (num_rows + world - 1) // world
    per = (per + align - 1) // align * align
    first = min(rank * per, num_rows)
    return first, min(first + per, num_rows)


def gather_selection(local_indices: torch.Tensor, count: int, dst: int = 0,
                     group: Optional[dist.ProcessGroup] = None,
                     out: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], int]:
    """Gather every rank's first `count` entries of `local_indices` onto rank `dst`, in rank
    order.  One all-gather of the counts, then variable-length point-to-point transfers of the
    index runs straight into their final position (no padding to the maximum count).
    Returns (global vector on dst | None elsewhere, total count)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_indices.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([count], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    c = [int(x) for x in counts.tolist()]
    total = sum(c)
    if rank == dst:
        if out is None or out.numel() < total:
            out = torch.empty(max(total, 1), dtype=local_indices.dtype, device=dev)
        ops, off = [], 0
        for r in range(world):
            if r == dst:
                out[off: off + c[r]].copy_(local_indices[: c[r]])
            elif c[r] > 0:
                ops.append(dist.P2POp(dist.irecv, out[off: off + c[r]], r, group))
            off += c[r]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out[:total], total
    if count > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_indices[:count].contiguous(), dst, group)]):
            w.wait()
    return None, total


class PeerGather:
    """SelectionVector gather without SMs: every rank writes its index run straight into rank
    `dst`'s buffer over NVLink with the copy engines (peer memory mapped through CUDA IPC), so
    the transfer runs underneath the persistent filter kernel of the next batch -- an NCCL
    send/recv kernel cannot (it finds no free CTA slot until the filter kernel ends).
    Counts travel host-side over a gloo group; the data never touches the host.

    Protocol per batch: `start(local, count)` -> all-gather of counts (gloo), then one async
    device-to-peer copy at this rank's prefix offset on `copy_stream`; `finish()` -> every rank
    waits for its own copy, then a gloo barrier tells `dst` that the vector is complete."""

    def __init__(self, capacity: int, dtype: torch.dtype, device: torch.device, dst: int = 0):
        from torch.multiprocessing.reductions import reduce_tensor
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.device = device
        self.cpu_group = dist.new_group(backend="gloo")
        self.copy_stream = torch.cuda.Stream(device)
        self.done = torch.cuda.Event()
        self.buf = None
        payload = [None]
        if self.rank == dst:
            self.buf = torch.empty(capacity, dtype=dtype, device=device)
            fn, args = reduce_tensor(self.buf)
            payload = [(fn, args)]
        dist.broadcast_object_list(payload, src=dst, group=self.cpu_group)
        if self.rank != dst:
            fn, args = payload[0]
            self.buf = fn(*args)          # rank dst's memory, mapped into this process
        self.total = 0
        self._pending = False

    def start(self, local: torch.Tensor, count: int, after: Optional[torch.cuda.Event] = None) -> int:
        # plain CPU tensors over gloo (the *_object collectives pickle and take milliseconds)
        mine = torch.tensor([int(count)], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.cpu_group)
        counts = [int(t[0]) for t in gathered]
        off = sum(counts[: self.rank])
        self.total = sum(counts)
        if self.total > self.buf.numel():
            raise RuntimeError("PeerGather capacity %d < %d selected rows" % (self.buf.numel(), self.total))
        with torch.cuda.stream(self.copy_stream):
            if after is not None:
                self.copy_stream.wait_event(after)
            if count > 0:
                self.buf[off: off + count].copy_(local[:count], non_blocking=True)
            self.done.record(self.copy_stream)
        self._pending = True
        return self.total

    def finish(self) -> Optional[torch.Tensor]:
        """Blocks the host until every rank's run has landed; returns the vector on `dst`."""
        if self._pending:
            self.done.synchronize()
            flag = torch.ones(1, dtype=torch.int32)
            dist.all_reduce(flag, group=self.cpu_group)   # acts as the barrier
            self._pending = False
        return self.buf[: self.total] if self.rank == self.dst else None
