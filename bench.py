#!/usr/bin/env python
"""bench.py — BASELINE.json metric: rows/sec (and HBM GB/s) of the TPC-H Q6 Filter
(shipdate range AND discount range AND quantity) over synthetic lineitem rows.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (CPU arm: oracle port on the host cores)

A "step" is one pass of the hot path (Filter::Evaluate -> one fused predicate + ordered
compaction kernel) over one batch of `--rows` lineitem rows per GPU (default 1e9 = configs[1]
of BASELINE.json; at N>1 each rank owns a contiguous row range of the same size: weak scaling,
configs[4] shape), inputs resident in HBM.  Inputs (20 GB) are far larger than L2 (126 MB),
so every step streams from HBM; no explicit flush is needed.
`value`   = rows/s over all ranks, device-resident (CUDA events, max over ranks).
`e2e`     = same metric through the C-ABI with HOST (pinned) buffers: H2D of every input
            column and D2H of the selection vector inside the timed region.
`roofline`= algorithmic bytes (20 B/row in + 4 B per selected row out) over the
            kernel's CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` = oracle (kind "port": scalar interpreter, NOT Gandiva's LLVM JIT which cannot
            be built in this image) on a bounded sample with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALGO_IN_BYTES_PER_ROW = 20.0   # date32 + 2 x float64, no validity buffers in the base run
IDX_BYTES = 4.0                # uint32 selection vector entries


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("GDV_BENCH_ROWS", 1_000_000_000)),
                    help="lineitem rows per GPU per step")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per e2e step (0 = same as --rows)")
    ap.add_argument("--e2e-chunk", type=int, default=32 * 1024 * 1024, help="rows per host RecordBatch")
    ap.add_argument("--cpu-rows", type=int, default=0, help="cpu_baseline sample rows (0 = auto)")
    ap.add_argument("--rows-per-thread", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--filter-loader", type=int, default=0,
                    help="Configuration.loader of the Q6 Filter: 0 = fused filter kernel, 3 = two-pass "
                         "(projector -> truth bitmap, gdv_bitmap_to_sel -> SelectionVector)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"],
                    help="N>1: how the SelectionVector reaches rank 0.  push (default): "
                         "gdv_selection_push, device-side NVLink stores into rank 0's vector, no host "
                         "sync; nccl: all-gather of counts + send/recv (host reads the count)")
    ap.add_argument("--no-overlap", action="store_true", help="--gather nccl: gather inside each step, no pipelining")
    ap.add_argument("--peer-gather", action="store_true",
                    help="--gather nccl: copy-engine peer writes (CUDA IPC) instead of NCCL send/recv; "
                         "measured slower in round 1 (host-side gloo sync), see DESIGN.md")
    ap.add_argument("--push-ctas", type=int, default=4, help="--gather push: CTAs of the push kernel")
    ap.add_argument("--sm-reserve", type=int, default=-1,
                    help="SMs left free for the push / NCCL kernels (default: --push-ctas when N>1)")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def q6_filter(gandiva, cases, cfg):
    b = gandiva.TreeExprBuilder()
    return gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cases.q6_condition(b)), cfg), b


def cpu_baseline(rows: int, threads: int, steps: int = 1, warmup: int = 0):
    """Oracle (port of the reference's CPU structure) on host cores; returns rows/s + details."""
    import numpy as np
    import pyarrow as pa
    import cases
    import gandiva_b200 as gandiva
    import oracle
    batch = cases.q6_batch(rows, seed=42)
    b = gandiva.TreeExprBuilder()
    cond = cases.q6_condition(b)
    for _ in range(warmup):
        oracle.filter_indices(cond, batch, threads=threads)
    t0 = time.perf_counter()
    total = 0
    for _ in range(max(steps, 1)):
        total += len(oracle.filter_indices(cond, batch, threads=threads))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows / dt, dt, total // max(steps, 1)


def run_reference(args):
    """--impl reference: the reference's CPU path.  dremio/gandiva cannot be built here (no
    source in /root/reference, no LLVM), so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = oracle.hardware_threads()
    rows = args.cpu_rows or 16 * 1024 * 1024
    rps, dt, cnt = cpu_baseline(rows, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "rows/sec (TPC-H Q6 filter over synthetic lineitem)",
        "value": rps, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "TPC-H Q6 filter, CPU sample of %d rows per step" % rows,
                   "rows_per_step": rows},
        "cpu_baseline": {"value": rps, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": "%d rows/step x %d steps, scalar interpreter oracle (not the "
                                   "LLVM-JIT reference: unbuildable here)" % (rows, args.steps)},
        "e2e": {"value": rps, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cases
    import gandiva_b200 as gandiva

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(dev)   # a real (non-default) stream: events, kernels and
    torch.cuda.set_stream(stream)     # NCCL ops are all ordered on it
    st = stream.cuda_stream

    n = args.rows
    first_row = rank * n
    idx_mode = "UINT32" if (world == 1 and n <= (1 << 32)) else "UINT64"
    idx_dtype = torch.int32 if idx_mode == "UINT32" else torch.int64
    use_push = world > 1 and not args.no_gather and args.gather == "push"
    sm_reserve = args.sm_reserve if args.sm_reserve >= 0 else (args.push_ctas if use_push else 0)
    cfg = gandiva.Configuration(device=local_rank, rows_per_thread=args.rows_per_thread,
                                block_threads=args.block_threads, sm_reserve=sm_reserve,
                                loader=args.filter_loader)
    filt, _ = q6_filter(gandiva, cases, cfg)

    # ---- inputs resident in HBM (generated on device; same stream as oracle/lineitem.h) -----
    ship = torch.empty(n, dtype=torch.int32, device=dev)
    disc = torch.empty(n, dtype=torch.float64, device=dev)
    qty = torch.empty(n, dtype=torch.float64, device=dev)
    for kind, t in ((0, ship), (1, disc), (2, qty)):
        gandiva.generate_lineitem(local_rank, kind, 42, first_row, n, t.data_ptr(), 0, 0, st)
    out_idx = None if use_push else torch.empty(n, dtype=idx_dtype, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    cols = [(0, ship.data_ptr(), 0, 0), (0, disc.data_ptr(), 0, 0), (0, qty.data_ptr(), 0, 0)]
    torch.cuda.synchronize()

    from gandiva_b200.sharding import gather_selection
    pipelined = world > 1 and not args.no_gather and not args.no_overlap and not use_push
    ps = None
    if use_push:
        from gandiva_b200.sharding import PeerSelection
        ps = PeerSelection(capacity=int(n * world * 0.03) + 4096, local_rows=n, mode=idx_mode, device=dev,
                           slots=2, ctas=args.push_ctas)
    gstep = {"i": 0}
    # N>1: the gather of batch i runs on a second stream while the filter kernel of batch i+1
    # runs (double-buffered index buffers).  The filter was built with sm_reserve so that NCCL's
    # copy CTAs find free slots next to the persistent filter CTAs.
    comm_stream = torch.cuda.Stream(dev) if world > 1 else None
    if use_push:                       # the kernel-only timing below writes here
        out_idx = ps.local[0] if rank != 0 else torch.empty(n, dtype=idx_dtype, device=dev)
    bufs = [out_idx, torch.empty(n, dtype=idx_dtype, device=dev) if pipelined else out_idx]
    cnts = [d_count, torch.zeros(1, dtype=torch.int64, device=dev)]
    host_cnt = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)]
    ev_k = [torch.cuda.Event() for _ in range(2)]
    ev_g = [torch.cuda.Event() for _ in range(2)]
    gathered = {"buf": None}
    peer = None
    if pipelined and args.peer_gather:
        try:   # copy-engine gather into rank 0's IPC-mapped buffer (no SMs: overlaps the next kernel)
            from gandiva_b200.sharding import PeerGather
            peer = PeerGather(int(n * world * 0.03) + 1024, idx_dtype, dev, dst=0)
        except Exception as e:  # pragma: no cover - fall back to NCCL send/recv
            if rank == 0:
                print("PeerGather unavailable (%s); using NCCL send/recv" % e, file=sys.stderr)
            peer = None

    def launch(i):
        b = i % 2
        filt.evaluate_device(n, cols, bufs[b].data_ptr(), n, idx_mode, st, cnts[b].data_ptr(),
                             sync=False, index_base=first_row)
        if world > 1 and not args.no_gather:
            host_cnt[b].copy_(cnts[b], non_blocking=True)
            ev_k[b].record(stream)

    def gather(i):
        b = i % 2
        ev_k[b].synchronize()              # kernel i and its count copy are done
        cnt = int(host_cnt[b][0])
        if peer is not None:
            peer.finish()                  # previous batch's vector is complete on rank 0
            total = peer.start(bufs[b], cnt, after=ev_k[b])
            ev_g[b] = peer.done
            return total
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ev_k[b])
            out, total = gather_selection(bufs[b], cnt, dst=0, out=gathered["buf"])
            if rank == 0 and (gathered["buf"] is None or out.numel() > gathered["buf"].numel()):
                gathered["buf"] = torch.empty(int(total * 1.1) + 16, dtype=idx_dtype, device=dev)
            ev_g[b].record(comm_stream)
        return total

    def run_steps_push(k, events=None):
        """N>1, --gather push: filter kernel i+1 overlaps the NVLink push of run i; nothing
        touches the host between steps."""
        for i in range(k):
            g = gstep["i"]
            gstep["i"] += 1
            ps.before_filter(g, stream)
            ptr, cap, mode, cnt_ptr = ps.filter_target(g)
            filt.evaluate_device(n, cols, ptr, cap, mode, st, cnt_ptr, sync=False, index_base=first_row)
            # the last run of the job has no filter kernel to hide under: push it with the whole GPU
            ps.after_filter(g, stream, ctas=(2 * 148 if i == k - 1 else 0))
            if events is not None and i + 1 < k:
                events[i + 1].record(stream)
        ps.finish(stream)                          # the last step closes after its push landed
        if events is not None:
            events[k].record(stream)
        return None

    def run_steps(k, events=None):
        """k passes of the hot path; with N>1 each pass ends with the SelectionVector on rank 0."""
        if use_push:
            return run_steps_push(k, events)
        total = None
        for i in range(k):
            if pipelined and i >= 2:
                stream.wait_event(ev_g[i % 2])      # buffer i%2 is free again
            launch(i)
            if world > 1 and not args.no_gather:
                if pipelined:
                    if i >= 1:
                        total = gather(i - 1)
                else:
                    total = gather(i)
                    stream.wait_event(ev_g[i % 2])
            if events is not None:
                events[i + 1].record(stream)
        if pipelined and k >= 1:
            total = gather(k - 1)
            if peer is not None:
                peer.finish()
            stream.wait_event(ev_g[(k - 1) % 2])
            if k >= 2:
                stream.wait_event(ev_g[k % 2])
            if events is not None:
                events[k].record(stream)          # the last event closes after the last gather
        return total

    run_steps(max(args.warmup, 3))
    count = filt.sync(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = gandiva.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    run_steps(args.steps, ev)
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = gandiva.launch_count() - launches0
    count = filt.sync(st)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        tc = torch.tensor([count], dtype=torch.int64, device=dev)
        dist.all_reduce(tc)
        total_selected = int(tc.item())
    else:
        total_selected = count
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)
    gather_check = None
    full_check = None
    if world == 1:
        # outside the timed region, at the full bench size: the SelectionVector of the last step is
        # ascending, its length is the count of an independent torch evaluation of the predicate,
        # and it holds exactly torch.nonzero of that mask
        try:
            torch.cuda.synchronize()
            mask = (ship >= 8766) & (ship < 9131) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24)
            want_n = int(mask.sum().item())
            idx = out_idx[:count].to(torch.int64)
            if idx_mode == "UINT32":
                idx = idx & 0xFFFFFFFF
            ok = (count == want_n)
            if ok and count > 1:
                ok = bool((idx[1:] > idx[:-1]).all().item())
            ok = ok and bool(torch.equal(idx, torch.nonzero(mask).flatten()))
            full_check = ("ok: %d rows == torch mask count, ascending, == nonzero(mask)" % count) if ok else \
                         ("FAILED: count %d, torch %d" % (count, want_n))
            del mask, idx
        except Exception as e:  # noqa: BLE001 - the bench line must still be printed
            full_check = "not run: %r" % (e,)
    if use_push:
        # outside the timed region: the last step's vector on rank 0 is complete, ascending and
        # holds exactly the rows all ranks selected
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            vec, total = ps.result(gstep["i"] - 1)
            ok = (total == total_selected) and not ps.overflowed()
            if ok and total > 1:
                ok = bool((vec[1:] > vec[:-1]).all().item())
            ok = ok and int(vec[-1].item()) < n * world and int(vec[0].item()) >= 0
            gather_check = "ok: %d ascending global indices on rank 0" % total if ok else "FAILED"

    # ---- roofline of the dominant (only) kernel ----------------------------------------------
    peak, peak_src = measured_peak_gbs()
    algo_bytes = ALGO_IN_BYTES_PER_ROW * n + (IDX_BYTES if idx_mode == "UINT32" else 8.0) * count
    # kernel-only duration: time K evaluate calls without the gather
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kreps = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    kev[0].record(stream)
    for _ in range(kreps):
        filt.evaluate_device(n, cols, out_idx.data_ptr(), n, idx_mode, st, d_count.data_ptr(),
                             sync=False, index_base=first_row)
    kev[1].record(stream)
    torch.cuda.synchronize()
    filt.sync(st)
    kernel_ms = kev[0].elapsed_time(kev[1]) / kreps
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    info = filt.kernel_info
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel/config
        tj = json.load(open(os.path.join(ROOT, "profiles", "q6_filter_traffic.json")))
        if int(tj["rows"]) == n and idx_mode == "UINT32":
            traffic = float(tj["traffic_bytes_per_launch"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": info["name"].rsplit("_", 1)[0], "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "regs": info["regs"],
                "rows_per_thread": info["rows_per_thread"], "block_threads": info["block_threads"]}

    # ---- e2e: host (pinned) buffers through the C-ABI, copies inside the timed region --------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, gandiva, cases, torch, np, dev, local_rank, rank, world, n,
                      ship, disc, qty, first_row)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        threads = oracle.hardware_threads()
        rows = args.cpu_rows or 16 * 1024 * 1024
        rps, dt, _ = cpu_baseline(rows, threads)
        if dt < 5.0 and not args.cpu_rows:  # aim for ~10-30 s of CPU work
            rows = int(min(rows * 12.0 / max(dt, 1e-3), 512 * 1024 * 1024)) // 64 * 64
            rps, dt, _ = cpu_baseline(rows, threads)
        cpu = {"value": rps, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": "%d rows of the same synthetic lineitem, oracle scalar interpreter on %d "
                         "threads, %.1f s" % (rows, threads, dt)}

    if rank == 0:
        line = {
            "metric": "rows/sec (TPC-H Q6 filter over synthetic lineitem)", "value": value,
            "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "TPC-H Q6 filter (BASELINE.json configs[1]%s)" %
                                   ("; configs[4] sharding" if world > 1 else ""),
                       "rows_per_gpu": n, "total_rows": n * world, "selectivity": total_selected / (n * world),
                       "selection_vector": idx_mode + ((" reassembled on rank 0 over NVLink" + (", gdv_selection_push: device-side stores into rank 0's vector (CUDA IPC), counts exchanged through a board in rank 0's HBM, no host sync, overlapped with the next batch's kernel (%d SMs reserved)" % sm_reserve if use_push else ((", copy-engine peer writes into rank 0 (CUDA IPC)" if peer is not None else ", NCCL send/recv") + ", overlapped with the next batch's kernel" if pipelined else ", NCCL send/recv"))) if world > 1 and not args.no_gather else ""),
                       "gather_check": gather_check, "full_size_check": full_check,
                       "l2_policy": "inputs (20 B/row x %d rows) larger than L2; no flush" % n,
                       "parallelism": "row-range shards, %d" % world},
            "hbm_gbs": achieved, "per_step_ms": per_step,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        if ps is not None:
            ps.close()
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, gandiva, cases, torch, np, dev, local_rank, rank, world, n, ship, disc, qty,
            first_row):
    """Q6 through gdv_filter_evaluate with HOST buffers: every step copies its input columns
    H2D from pinned host memory, runs the fused kernel, and copies the selection vector D2H."""
    import ctypes as C
    import psutil
    rows = args.e2e_rows or n
    avail = psutil.virtual_memory().available
    need = rows * 24 + (1 << 30)
    if need > avail * 0.5:
        rows = int(avail * 0.5 - (1 << 30)) // 24 // 1024 * 1024
    chunk = min(args.e2e_chunk, rows)
    rows = rows // chunk * chunk
    if rows <= 0:
        return None
    # host-resident copy of the first `rows` rows of this rank's shard, in pinned memory
    def pinned(nbytes):
        p = C.c_void_p()
        gandiva._check(gandiva.lib.gdv_host_alloc(nbytes, C.byref(p)))
        return p
    h_ship, h_disc, h_qty = pinned(rows * 4), pinned(rows * 8), pinned(rows * 8)
    h_idx = pinned(chunk * 4)
    try:
        def host_view(ptr, count, np_dtype):
            ct = np.ctypeslib.as_ctypes_type(np_dtype)
            return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(count,)))
        host_view(h_ship, rows, np.int32).copy_(ship[:rows])
        host_view(h_disc, rows, np.float64).copy_(disc[:rows])
        host_view(h_qty, rows, np.float64).copy_(qty[:rows])
        torch.cuda.synchronize()
        filt, _ = q6_filter(gandiva, cases, gandiva.Configuration(device=local_rank))
        stream = torch.cuda.current_stream()

        def one_pass():
            total = 0
            for c0 in range(0, rows, chunk):
                cols = (gandiva.gdv_column_t * 3)()
                cols[0].values = h_ship.value + c0 * 4
                cols[1].values = h_disc.value + c0 * 8
                cols[2].values = h_qty.value + c0 * 8
                cb = gandiva.gdv_batch_t(chunk, 3, gandiva.GDV_MEM_HOST, cols)
                sel = gandiva.gdv_selection_t(h_idx, chunk, 0, gandiva.GDV_SEL_UINT32,
                                              gandiva.GDV_MEM_HOST, 0)
                gandiva._check(gandiva.lib.gdv_filter_evaluate(filt._h, C.byref(cb), C.byref(sel),
                                                               gandiva._stream_handle(stream.cuda_stream), 0, None))
                total += sel.num_slots
            return total
        one_pass()
        steps = max(1, min(args.steps, 3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        sel_total = 0
        for _ in range(steps):
            sel_total = one_pass()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps
        ms = e0.elapsed_time(e1) / steps
        sec = max(wall, ms * 1e-3)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return {"value": rows * world / sec, "unit": "rows/s",
                "h2d_bytes_per_step": int(rows * 20) * world, "d2h_bytes_per_step": int(sel_total * 4 + 8 * (rows // chunk)) * world,
                "rows_per_step": rows * world, "chunk_rows": chunk, "ms_per_step": sec * 1e3,
                "api": "gdv_filter_evaluate(GDV_MEM_HOST), pinned host buffers, one call per %d-row RecordBatch" % chunk}
    finally:
        for p in (h_ship, h_disc, h_qty, h_idx):
            gandiva.lib.gdv_host_free(p)


if __name__ == "__main__":
    main()
