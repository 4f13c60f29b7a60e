#!/usr/bin/env python
"""Pinned host -> device copy ceiling of this box with 1, 2, 4, 8 GPUs uploading AT THE SAME TIME.

    python tools/h2d_ceiling.py [--mb 805] [--reps 6] [--gpus 1,2,4,8]

What bench.py's `e2e` leg is bounded by: every rank uploads its 0.8 GB level image per step, so the aggregate
host-memory / PCIe path of the VM decides how far the end-to-end number can scale.  One worker process per GPU
(spawned here, no torch.distributed), each with its own pinned buffer (first touched by the worker, bound to the GPU's
NUMA node when NVML reports one), started on a shared barrier; every worker times `reps` back-to-back
cudaMemcpyAsync-sized copies with CUDA events.  Also times the host-side block-max scan of a lesion mask of the bench's
size (the other host cost inside the e2e step) with the thread count a rank gets.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import time


def worker(idx, n, mb, reps, barrier, q):
    import torch
    torch.cuda.set_device(idx)
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if len(near) >= 2:
            os.sched_setaffinity(0, near)
    except Exception:
        pass
    nbytes = mb << 20
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    q.put((idx, nbytes * reps / (ms * 1e-3) / 1e9))


def mask_scan(threads, h=16384 + 1568, w=16384):
    import torch
    torch.set_num_threads(threads)
    m = torch.zeros((h, w), dtype=torch.uint8).pin_memory()
    R = 32
    n = h // R
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        m[:n * R].view(n, R * w).amax(dim=1).ne(0).tolist()
        t.append(time.perf_counter() - t0)
    return min(t) * 1e3, h * w / min(t) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=805)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--gpus", default="1,2,4,8")
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    out = {"tool": "h2d_ceiling", "mb_per_copy": args.mb, "reps": args.reps, "gpus_in_box": have, "host_cpus": os.cpu_count(), "concurrent": {}}
    ctx = mp.get_context("spawn")
    for n in [int(v) for v in args.gpus.split(",") if int(v) <= have]:
        barrier, q = ctx.Barrier(n), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(i, n, args.mb, args.reps, barrier, q)) for i in range(n)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=300) for _ in range(n))
        [p.join() for p in ps]
        per = [round(r[1], 2) for r in res]
        out["concurrent"][str(n)] = {"per_gpu_GBps": per, "min_GBps": min(per), "aggregate_GBps": round(sum(per), 1)}
    cpus = len(os.sched_getaffinity(0))
    out["mask_block_max_scan"] = {}
    for n in (1, 8):
        th = max(1, min(16, cpus // n))
        ms, gbs = mask_scan(th)
        out["mask_block_max_scan"][f"threads_{th}_(rank_share_at_N={n})"] = {"ms": round(ms, 2), "GBps": round(gbs, 1)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
