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

from . import _lib


def shard_rows(ny_total: int, world: int, rank: int):
    """Contiguous candidate grid-row range ``[i0, i1)`` of ``rank``; sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(ny_total, world)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def cyclic_blocks(ny_total: int, world: int, rank: int, block_rows: int):
    """Block-cyclic sharding of the candidate grid rows: blocks of ``block_rows`` rows are dealt round-robin, rank ``r``
    owns blocks ``r, r + world, r + 2 world, ...`` (tissue is spatially clumped; contiguous shards of a whole-slide level
    leave some ranks without a single survivor).  Returns ``(blocks, segs_per_rank)``: ``blocks`` = this rank's
    ``(block_index, i0, i1)`` grid-row ranges in ascending order, ``segs_per_rank`` = the per-rank block count every rank
    must present to ``SurvivorExchange(..., cyclic=True)`` (ranks with one block fewer pack an empty segment)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    block_rows = max(1, int(block_rows))
    nblocks = (ny_total + block_rows - 1) // block_rows
    segs_per_rank = (nblocks + world - 1) // world
    blocks = [(b, b * block_rows, min((b + 1) * block_rows, ny_total)) for b in range(rank, nblocks, world)]
    return blocks, segs_per_rank


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


class SurvivorExchange:
    """The exchange step on the CUDA path: fixed-size, count-free, sort-free (``csrc/exchange.cu``).

    Every rank owns ``segs_per_rank`` SEGMENTS (one per tile-scan + forward pass over a contiguous range of candidate grid
    rows, in ascending y order); ``pack`` writes a segment's capacity-sized outputs and its DEVICE-side survivor count
    into the send buffer, ``merge`` runs ONE ``all_gather_into_tensor`` (NCCL over NVLink; the counts travel in the
    segment headers) and the library's index + scatter kernels, which place every row at its final position in the
    canonical ``(x, y)`` order -- array-equal to a single-rank run.  Nothing waits on the host: all of it is enqueued
    behind the network kernels and the next step's kernels queue up behind it; ``total`` stays on the device until
    ``result()`` is asked for.  With ``world == 1`` the same merge turns the row groups of one rank into one ordered
    result (no collective)."""

    def __init__(self, device, seg_capacity: int, num_classes: int, nx: int, stride: int, segs_per_rank: int = 1, group=None,
                 with_features: bool = True, cyclic: bool = False, cyclic_world: int | None = None):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cap, self.k, self.nx, self.stride, self.spr = int(seg_capacity), int(num_classes), int(nx), int(stride), int(segs_per_rank)
        self.fd = 512 if with_features else 0
        # cyclic: local segment g of rank r is grid-row block g * world + r of the level (block-cyclic sharding: balances
        # clumped tissue); otherwise every rank owns one contiguous row range and its segments follow each other in y
        self.cyclic = bool(cyclic)
        self._cyclic_world = cyclic_world      # tests: emulate the rank-major storage of W ranks inside one process
        l = _lib.lib()
        self.seg_bytes = int(l.hipac_exchange_segment_bytes(self.cap, self.fd, self.k))
        self.nseg = self.world * self.spr
        self.send = torch.empty((self.spr * self.seg_bytes,), dtype=torch.uint8, device=self.device)
        self.recv = torch.empty((self.nseg * self.seg_bytes,), dtype=torch.uint8, device=self.device) if self.world > 1 else self.send
        self.out_cap = self.nseg * self.cap
        self.coords = torch.empty((max(self.out_cap, 1), 2), dtype=torch.int32, device=self.device)
        self.labels = torch.empty((max(self.out_cap, 1),), dtype=torch.uint8, device=self.device)
        self.features = torch.empty((max(self.out_cap, 1), 512), dtype=torch.float32, device=self.device) if self.fd else None
        self.logits = torch.empty((max(self.out_cap, 1), max(self.k, 1)), dtype=torch.float32, device=self.device)
        self.total = torch.zeros((2,), dtype=torch.int32, device=self.device)
        self.ws = torch.empty((int(l.hipac_exchange_workspace_bytes(self.nseg, self.nx)),), dtype=torch.uint8, device=self.device)
        self._total_host = torch.zeros((2,), dtype=torch.int32).pin_memory() if self.device.type == "cuda" else None

    def pack(self, seg: int, coords, labels, features, logits, count: torch.Tensor, y_offset: int = 0, stream=None):
        """Enqueue the packing of local segment ``seg``; ``count`` is the int32 device counter of the scan (element 0) and
        the tensors hold at least ``min(count, their own first dimension)`` valid rows (at most the segment capacity)."""
        if not (0 <= seg < self.spr):
            raise ValueError(f"segment {seg} outside [0, {self.spr})")
        rows = int(coords.shape[0])
        if rows > self.cap:
            raise ValueError(f"{rows} candidate rows exceed the segment capacity {self.cap}")
        if (self.fd and (features is None or int(features.shape[0]) < rows)) or (self.k and (logits is None or int(logits.shape[0]) < rows)):
            raise ValueError("feature / logit tensors are missing or shorter than the coordinate tensor")
        st = stream or torch.cuda.current_stream(self.device)
        dst = self.send[seg * self.seg_bytes:]
        _lib.check(_lib.lib().hipac_exchange_pack(coords.data_ptr(), labels.data_ptr(), features.data_ptr() if self.fd else None,
                                                  logits.data_ptr() if self.k else None, self.fd, self.k,
                                                  count.data_ptr(), rows, int(y_offset), dst.data_ptr(), st.cuda_stream),
                   "hipac_exchange_pack")

    def pack_empty(self, seg: int, stream=None):
        """Mark local segment ``seg`` as holding no rows (a rank with fewer row groups than ``segs_per_rank``)."""
        if getattr(self, "_zero", None) is None:
            self._zero = torch.zeros((2,), dtype=torch.int32, device=self.device)
        st = stream or torch.cuda.current_stream(self.device)
        dst = self.send[seg * self.seg_bytes:]
        _lib.check(_lib.lib().hipac_exchange_pack(self.coords.data_ptr(), self.labels.data_ptr(), self.features.data_ptr() if self.fd else None,
                                                  self.logits.data_ptr() if self.k else None, self.fd, self.k, self._zero.data_ptr(), 0, 0,
                                                  dst.data_ptr(), st.cuda_stream), "hipac_exchange_pack")

    def merge(self, stream=None):
        """Enqueue the all-gather (world > 1) and the index + scatter kernels; returns nothing (see ``result``)."""
        st = stream or torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            if self.world > 1:
                dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
            _lib.check(_lib.lib().hipac_exchange_merge(self.recv.data_ptr(), self.nseg, self.cap, self.fd, self.k, self.stride, self.nx,
                                                       (self._cyclic_world or self.world) if self.cyclic else 0,
                                                       self.coords.data_ptr(), self.labels.data_ptr(),
                                                       self.features.data_ptr() if self.fd else None,
                                                       self.logits.data_ptr() if self.k else None, self.total.data_ptr(), self.out_cap,
                                                       self.ws.data_ptr(), int(self.ws.numel()), st.cuda_stream),
                       "hipac_exchange_merge")

    def result(self) -> dict:
        """Wait for the merge and slice the outputs to the true row count (the one host read of the step)."""
        self._total_host.copy_(self.total, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        n = min(int(self._total_host[0]), self.out_cap)
        out = {"coords": self.coords[:n], "labels": self.labels[:n]}
        if self.fd:
            out["features"] = self.features[:n]
        if self.k:
            out["logits"] = self.logits[:n]
        return out
