"""N>1 host logic on CPU: world_size-2 gloo run of the row-range sharding + SelectionVector
gather used by bench.py --gpus N (the filter itself is replaced by the oracle here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import gandiva_b200 as gandiva
    import oracle
    from gandiva_b200.sharding import gather_selection, shard_range
    first, last = shard_range(n, world, rank)
    batch = cases.q6_batch(n, seed=42).slice(first, last - first)
    b = gandiva.TreeExprBuilder()
    local = oracle.filter_indices(cases.q6_condition(b), batch) + np.uint64(first)   # global row numbers
    t = torch.from_numpy(local.astype(np.int64))
    out, total = gather_selection(t, len(local), dst=0)
    if rank == 0:
        q.put((out.numpy().copy(), total))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_alignment():
    from gandiva_b200.sharding import shard_range
    n, world = 1_000_003, 8
    edges = [shard_range(n, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and b0 % 64 == 0
    assert shard_range(10, 4, 3) == (10, 10)  # empty tail shard


@pytest.mark.timeout(120)
def test_gather_selection_world2():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cases
    import gandiva_b200 as gandiva
    import oracle
    n, world = 300_001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, total = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = gandiva.TreeExprBuilder()
    want = oracle.filter_indices(cases.q6_condition(b), cases.q6_batch(n, seed=42))
    assert total == len(want)
    assert np.array_equal(got.astype(np.uint64), want)


def test_shard_rows_with_root():
    """Shards for a root-gathered Filter: contiguous, aligned, exact total, the root's range shorter by what
    absorbing (world-1) runs costs; equal ranges when nothing is selected."""
    from gandiva_b200.sharding import shard_rows_with_root
    for world in (1, 2, 4, 8):
        for total in (10_000_000_000, 1_000_003, 64 * world):
            for sel in (0.0, 0.0181, 0.3):
                r = shard_rows_with_root(total, world, sel, 20.0)
                assert len(r) == world and sum(r) == total and min(r) > 0
                assert all(x % 64 == 0 for x in r[:-1]) or world == 1
                if world > 1 and sel > 0 and total > 1_000_000:
                    assert r[0] < r[1]
                if world > 1 and sel == 0.0 and total % (64 * world) == 0:
                    assert len(set(r)) == 1
    r = shard_rows_with_root(10_000_000_000, 8, 0.0181, 20.0)
    assert abs(r[0] / 1.25e9 - 0.911) < 0.01 and abs(r[1] / 1.25e9 - 1.0127) < 0.005


def test_wave_layout_is_a_wave_major_partition():
    """Slices of a batch filtered in waves: every rank's rows are all used, slices are aligned, and walking the
    slices wave by wave, rank by rank, walks the global row space from 0 to the total without gap or overlap —
    which is why the concatenated runs are the ascending SelectionVector of the whole table."""
    from gandiva_b200.sharding import shard_rows_with_root, wave_layout
    for shard in ([1000], [5_000_003, 4_999_936], shard_rows_with_root(10_000_000_000, 8, 0.0127, 20.0)):
        for waves in (1, 2, 4, 7):
            rows, first = wave_layout(shard, waves)
            assert [sum(r) for r in rows] == list(shard)
            pos = 0
            for j in range(waves):
                for r in range(len(shard)):
                    assert first[r][j] == pos
                    pos += rows[r][j]
                    if j < waves - 1 and rows[r][j] and rows[r][j] != shard[r]:   # (a tiny shard is one slice)
                        assert rows[r][j] % 1024 == 0
            assert pos == sum(shard)
