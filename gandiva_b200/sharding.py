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
