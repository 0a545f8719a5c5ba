"""torchrun worker for test_peer_selection_push (needs >= 2 GPUs): every rank filters its row
range of the synthetic lineitem table, PeerSelection reassembles the SelectionVector on rank 0
with gdv_selection_push, and rank 0 compares it with the CPU oracle over the whole table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import gandiva_b200 as gandiva  # noqa: E402
from gandiva_b200.sharding import PeerSelection, shard_range, wave_layout  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    total_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_003
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    waves = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ranges = [shard_range(total_rows, world, r) for r in range(world)]
    first, last = ranges[rank]
    n = last - first
    # waves > 1: every rank's shard is filtered in slices; the global row order is wave-major, i.e. slice
    # (rank r, wave j) IS the global row range [wfirst[j], wfirst[j] + wrows[j]) of the same table
    all_rows, all_first = wave_layout([b_ - a_ for a_, b_ in ranges], waves)
    wrows, wfirst = all_rows[rank], all_first[rank]
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    ship = torch.empty(n, dtype=torch.int32, device=dev)
    disc = torch.empty(n, dtype=torch.float64, device=dev)
    qty = torch.empty(n, dtype=torch.float64, device=dev)
    b = gandiva.TreeExprBuilder()
    filt = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cases.q6_condition(b)),
                               gandiva.Configuration(device=local, sm_reserve=2))
    ps = PeerSelection(capacity=int(total_rows * 0.05) + 1024, local_rows=n if waves == 1 else max(wrows),
                       mode="UINT64", device=dev, slots=2, ctas=8, waves=waves)
    cols = [(0, ship.data_ptr(), 0, 0), (0, disc.data_ptr(), 0, 0), (0, qty.data_ptr(), 0, 0)]
    ok = True
    for step in range(steps):
        seed = 42 + step // 2          # two consecutive steps share a table, then it changes
        if waves == 1:
            ps.before_filter(step, stream)
            for kind, t in ((0, ship), (1, disc), (2, qty)):
                gandiva.generate_lineitem(local, kind, seed, first, n, t.data_ptr(), 0, 0, st)
            ptr, cap, mode, cnt = ps.filter_target(step)
            filt.evaluate_device(n, cols, ptr, cap, mode, st, cnt, sync=False, index_base=first)
            ps.after_filter(step, stream)
        else:
            at = 0
            for j in range(waves):
                ps.before_filter(step, stream, wave=j)
                for kind, t in ((0, ship), (1, disc), (2, qty)):
                    gandiva.generate_lineitem(local, kind, seed, wfirst[j], wrows[j],
                                              t.data_ptr() + at * t.element_size(), 0, 0, st)
                wcols = [(0, ship.data_ptr() + 4 * at, 0, 0), (0, disc.data_ptr() + 8 * at, 0, 0),
                         (0, qty.data_ptr() + 8 * at, 0, 0)]
                ptr, cap, mode, cnt = ps.filter_target(step, wave=j)
                filt.evaluate_device(wrows[j], wcols, ptr, cap, mode, st, cnt, sync=False, index_base=wfirst[j])
                ps.after_filter(step, stream, wave=j)
                at += wrows[j]
        if step == steps - 1 or step == 1:
            ps.finish(stream)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                import oracle
                vec, total = ps.result(step)
                want = oracle.filter_indices(cases.q6_condition(b), cases.q6_batch(total_rows, seed=seed), threads=8)
                got = vec.cpu().numpy().astype(np.uint64)
                ok = ok and total == len(want) and np.array_equal(got, want) and not ps.overflowed()
            dist.barrier()
    filt.sync(st)
    ps.close()
    if rank == 0:
        print("PEER_PUSH_OK" if ok else "PEER_PUSH_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
