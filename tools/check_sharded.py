#!/usr/bin/env python
"""N-rank result == 1-rank result, bit for bit, on real GPUs (north_star: "bit-exact patch sets" under scaling).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/check_sharded.py

Every rank tiles its tile-row shard of one synthetic slide level (resident path AND the host-buffer path), the shards meet
in the exchange step (sharding.SurvivorExchange: device-side pack, one NCCL all-gather, index + scatter kernels) and every
rank ends up with the same canonically ordered (coords, labels, features, logits).  Rank 0 additionally runs the WHOLE
level alone (``pipeline.process_level``) and compares array-equal: coordinates, labels, feature bits, logit bits.  Prints
ONE JSON line on rank 0 and exits non-zero on any mismatch.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline, sharding  # noqa: E402
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import patch_and_stride  # noqa: E402
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import make_lesion_mask, make_level, seeded_resnet18  # noqa: E402


def crc_of(d):
    c = 0
    for k in ("coords", "labels", "features", "logits"):
        c = zlib.crc32(np.ascontiguousarray(d[k]).tobytes(), c)
    return f"{c:08x}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--width", type=int, default=5000)
    ap.add_argument("--rows-per-rank", type=int, default=2300)
    ap.add_argument("--groups", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ge.build()
    torch.set_num_threads(max(1, min(16, len(os.sched_getaffinity(0)) // world)))
    L, W, H = args.level, args.width, args.rows_per_rank * world
    P, S = patch_and_stride(L)
    ny = (H + S - 1) // S
    nx = (W + S - 1) // S
    i0, i1 = sharding.shard_rows(ny, world, rank)
    y0, y1 = sharding.slab_rows(i0, i1, S, P, H)
    seed = 77
    img_h = torch.from_numpy(make_level(seed, L, W, H, y0, y1)).pin_memory()
    msk_h = torch.from_numpy(make_lesion_mask(seed, L, W, H, y0, y1)).pin_memory()
    packed = features.pack_resnet18(seeded_resnet18(seed=0, classifier=True).state_dict(), dev)
    rows_max = max(sharding.shard_rows(ny, world, r)[1] - sharding.shard_rows(ny, world, r)[0] for r in range(world))

    # ---- resident path: row groups of at most ~half a shard, so that several segments per rank are exercised ----
    xchg = pipeline.exchange_for_level(dev, W, rows_max, S, 2, max_candidates=max(nx, nx * ((rows_max + 1) // 2)))
    img_d, msk_d = img_h.to(dev), msk_h.to(dev)
    pipeline.process_level_exchanged(img_d, msk_d, L, packed, xchg, row_range=(0, i1 - i0), y_offset=y0)
    got = {k: v.cpu().numpy().copy() for k, v in xchg.result().items()}

    # ---- host-buffer path (pinned H2D in row groups, every group one segment) ----
    pipe = pipeline.HostPipeline(int(img_h.shape[0]), W, dev, with_mask=True, num_classes=2)
    gb = pipeline.upload_group_bounds(0, rows_max, args.groups)
    xh = sharding.SurvivorExchange(dev, nx * (max(b - a for a, b in zip(gb, gb[1:])) + 1), 2, nx, S, segs_per_rank=len(gb) - 1)
    r = pipeline.process_level_host(img_h, msk_h, L, packed, pipe, row_range=(0, i1 - i0), groups=args.groups, exchange=xh, y_offset=y0)
    got_host = {"coords": r.coords.numpy().copy(), "labels": r.labels.numpy().copy(), "features": r.features.numpy().copy(),
                "logits": r.logits.numpy().copy()}

    # every rank must hold the same bytes
    crcs = [None] * world
    mine = (crc_of(got), crc_of(got_host))
    if world > 1:
        dist.all_gather_object(crcs, mine)
    else:
        crcs = [mine]
    ok_same = all(c == crcs[0] for c in crcs) and mine[0] == mine[1]

    ok_single, n_single = True, None
    if rank == 0:
        full_i = torch.from_numpy(make_level(seed, L, W, H)).to(dev)
        full_m = torch.from_numpy(make_lesion_mask(seed, L, W, H)).to(dev)
        ref = pipeline.process_level(full_i, full_m, L, packed)
        want = {"coords": ref.coords.cpu().numpy(), "labels": ref.labels.cpu().numpy(), "features": ref.features.cpu().numpy(),
                "logits": ref.logits.cpu().numpy()}
        n_single = len(ref)
        for name, g in (("resident", got), ("host", got_host)):
            for k in want:
                if not (g[k].shape == want[k].shape and np.array_equal(g[k].view(np.uint8) if g[k].dtype != np.uint8 else g[k],
                                                                       want[k].view(np.uint8) if want[k].dtype != np.uint8 else want[k])):
                    ok_single = False
                    print(f"MISMATCH {name}.{k}: {g[k].shape} vs {want[k].shape}", file=sys.stderr)
        print(json.dumps({"check": "N-rank == 1-rank (coords, labels, feature bits, logit bits)", "world": world, "level": L,
                          "level_image": [W, H], "patch": P, "stride": S, "survivors_single_rank": n_single,
                          "survivors_gathered": int(got["coords"].shape[0]), "tumor_labelled": int(want["labels"].sum()),
                          "segments_resident": xchg.nseg, "segments_host": xh.nseg,
                          "all_ranks_hold_identical_bytes": bool(ok_same), "equals_single_rank": bool(ok_single),
                          "patch_set_crc32": mine[0], "per_rank_crc32": [c[0] for c in crcs]}), flush=True)
    flag = torch.tensor([1 if (ok_same and ok_single) else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
