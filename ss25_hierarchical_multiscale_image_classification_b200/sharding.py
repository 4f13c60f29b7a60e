"""Tile-row sharding of the candidate grid across ranks and the one exchange step of the path.

The reference has no multi-GPU support on this path (``nn.DataParallel`` wraps a model that is never
run, ``src/main.py:839-842``).  Here the unit of work is one candidate patch; ranks own contiguous
candidate *grid rows* (``y // stride``) and need no data-path collective: the only exchange is the
variable-length gather of per-rank survivor counts, coordinates, labels, features and logits
(SURVEY.md section 8e).  Works with any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU
tests); nothing here touches pixels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(ny_total: int, world: int, rank: int):
    """Contiguous candidate grid-row range ``[i0, i1)`` of ``rank``; sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(ny_total, world)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def slab_rows(i0: int, i1: int, stride: int, patch: int, height: int):
    """Level rows ``[y0, y1)`` a rank must hold to tile grid rows ``[i0, i1)``: its own rows plus a
    read-only halo of ``patch - stride`` rows (no inter-GPU halo exchange: the source is the host)."""
    if i1 <= i0:
        return i0 * stride, i0 * stride
    return i0 * stride, min(height, (i1 - 1) * stride + patch)


def canonical_order(coords: torch.Tensor) -> torch.Tensor:
    """Permutation that sorts ``(x, y)`` rows into the reference's emission order (x outer, y inner)."""
    key = coords[:, 0].to(torch.int64) * (1 << 32) + coords[:, 1].to(torch.int64)
    return torch.argsort(key, stable=True)


def gather_survivors(tensors: dict, group=None, sort: bool = True, count_group=None) -> dict:
    """All-gather variable-length per-rank results.

    ``tensors`` maps names to tensors whose first dimension is this rank's survivor count (it must contain
    ``"coords"`` int32 ``[n,2]`` in GLOBAL level coordinates).  Every rank returns the concatenation over
    ranks, in canonical ``(x, y)`` order when ``sort`` -- array-equal to a single-rank run.

    One exchange step: the per-survivor rows of all tensors are packed side by side into one byte matrix, so the
    whole result travels in a single ``all_gather_into_tensor`` (after the tiny all-gather of the counts that
    sizes the padding).

    ``count_group``: an optional CPU-side (gloo) group over the same ranks.  Each rank already knows its own count
    on the host (it sliced its tensors with it), so exchanging the counts over the CPU group keeps the GPU queue
    free of host round trips: the data all-gather, the concatenation and the sort are then all enqueued
    asynchronously and the next step's kernels can be queued behind them.  Without it the counts travel over
    ``group`` and are read back, which drains the stream once per call."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = dict(tensors)
    else:
        ref = tensors["coords"]
        dev, n = ref.device, int(ref.shape[0])
        if count_group is not None:
            counts = torch.empty((world,), dtype=torch.int64)
            dist.all_gather_into_tensor(counts, torch.tensor([n], dtype=torch.int64), group=count_group)
            counts = counts.tolist()
        else:
            cnt = torch.tensor([n], dtype=torch.int64, device=dev)
            counts = torch.empty((world,), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(counts, cnt, group=group)
            counts = counts.tolist()
        mx = max(counts)
        # pack: each tensor contributes (row bytes rounded up to 8) columns of a uint8 matrix
        views, widths = [], []
        for t in tensors.values():
            rowb = t.element_size()
            for d in t.shape[1:]:
                rowb *= int(d)
            flat = t.contiguous().reshape(n, -1).view(torch.uint8) if n else None
            views.append((flat, rowb))
            widths.append((rowb + 7) // 8 * 8)
        packed = torch.zeros((mx, sum(widths)), dtype=torch.uint8, device=dev)
        off = 0
        for (flat, rowb), w in zip(views, widths):
            if n:
                packed[:n, off:off + rowb] = flat.reshape(n, rowb)
            off += w
        buf = torch.empty((world * mx, sum(widths)), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(buf, packed, group=group)
        rows = torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)]) if world > 1 else buf[:n]
        out, off = {}, 0
        total = rows.shape[0]
        for (name, t), (_, rowb), w in zip(tensors.items(), views, widths):
            col = rows[:, off:off + rowb].contiguous()
            out[name] = col.view(t.dtype).reshape((total,) + tuple(t.shape[1:]))
            off += w
    if sort and out["coords"].shape[0]:
        perm = canonical_order(out["coords"])
        out = {k: v[perm] for k, v in out.items()}
    return out
