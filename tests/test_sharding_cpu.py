"""Host-side multi-rank logic on CPU: row sharding, halo slabs and the variable-length gather (gloo, world 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hipac_oracle as orc
from ss25_hierarchical_multiscale_image_classification_b200 import sharding
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide


def test_shard_rows_partition_the_grid():
    for ny in (0, 1, 7, 74, 447):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_rows(ny, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == ny
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_rows(10, 2, 2)


def test_cyclic_blocks_partition_the_grid_and_agree_on_segment_count():
    for ny, world, B in [(447, 8, 8), (74, 8, 8), (19, 2, 3), (5, 8, 8), (64, 4, 16), (1, 3, 1), (100, 7, 1)]:
        seen, spr = [], None
        for r in range(world):
            blocks, s = sharding.cyclic_blocks(ny, world, r, B)
            spr = s if spr is None else spr
            assert s == spr and spr - 1 <= len(blocks) <= spr                            # every rank presents spr segments (at most one empty)
            assert [b for b, _, _ in blocks] == list(range(r, (ny + B - 1) // B, world))   # round-robin, ascending
            for b, i0, i1 in blocks:
                assert i0 == b * B and 0 < i1 - i0 <= B and i1 <= ny
                seen += list(range(i0, i1))
        assert sorted(seen) == list(range(ny))                                           # disjoint cover of the grid rows
    with pytest.raises(ValueError):
        sharding.cyclic_blocks(10, 2, 2, 4)


def test_slab_rows_cover_every_patch_of_the_shard():
    H, S, P = 16384, 224, 1792
    ny = (H + S - 1) // S
    for world in (2, 4, 8):
        for r in range(world):
            i0, i1 = sharding.shard_rows(ny, world, r)
            y0, y1 = sharding.slab_rows(i0, i1, S, P, H)
            assert y0 == i0 * S and y1 == min(H, (i1 - 1) * S + P)
            assert y1 - y0 <= (i1 - i0) * S + (P - S)


def test_canonical_order_matches_reference_emission_order():
    _, _, grid = orc.candidate_grid(1000, 900, 3)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(grid))
    shuffled = torch.from_numpy(grid[perm].astype(np.int32))
    back = shuffled[sharding.canonical_order(shuffled)]
    assert np.array_equal(back.numpy(), grid.astype(np.int32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, level, w0, h0, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slide = SyntheticSlide(w0, h0, seed=seed)
    img, mask = slide.level_array(level), slide.lesion_mask(level)
    p, s = orc.patch_and_stride(level)
    ny = (img.shape[0] + s - 1) // s
    i0, i1 = sharding.shard_rows(ny, world, rank)
    y0, y1 = sharding.slab_rows(i0, i1, s, p, img.shape[0])
    # each rank tiles ONLY its slab (own rows + halo); the oracle stands in for the CUDA kernels on CPU
    slab = np.full((max(y1 - y0, 1), img.shape[1], 3), 255, np.uint8)
    slab[: y1 - y0] = img[y0:y1]
    mine = orc.extract_patches_oracle(img[:y1], mask[:y1], level, row_range=(i0, i1), want_images=False)
    bottom = y1 >= img.shape[0]
    if not bottom:
        # a slab that stops above the image bottom must not change any patch of the shard
        full = orc.extract_patches_oracle(img, mask, level, row_range=(i0, i1), want_images=False)
        assert np.array_equal(full["coords"], mine["coords"]) and np.array_equal(full["labels"], mine["labels"])
    coords = torch.from_numpy(mine["coords"])
    feats = coords.float().sum(1, keepdim=True).repeat(1, 4)             # stand-in payload tied to the coords
    # odd ranks' worlds exchange the counts over a separate CPU group (the no-host-round-trip mode of the GPU bench)
    cg = dist.new_group(backend="gloo") if world == 3 else None
    out = sharding.gather_survivors({"coords": coords, "labels": torch.from_numpy(mine["labels"]), "features": feats},
                                    count_group=cg)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), coords=out["coords"].numpy(), labels=out["labels"].numpy(),
             features=out["features"].numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_gather_equals_single_rank(tmp_path, world):
    level, w0, h0, seed = 2, 12000, 9000, 1234
    mp.spawn(_worker, args=(world, _free_port(), level, w0, h0, seed, str(tmp_path)), nprocs=world, join=True)
    slide = SyntheticSlide(w0, h0, seed=seed)
    want = orc.extract_patches_oracle(slide.level_array(level), slide.lesion_mask(level), level, want_images=False)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["coords"], want["coords"])              # every rank holds the full, ordered result
        assert np.array_equal(got["labels"], want["labels"])
        assert np.array_equal(got["features"][:, 0], want["coords"].sum(1).astype(np.float32))
