"""The exchange step's kernels (csrc/exchange.cu) through the C ABI, on one GPU: segments in, canonical (x, y) order out.

The merge must reproduce ``np.lexsort`` of the concatenated segments -- which is the reference's emission order for the
whole level (x outer, y inner, src/main.py:682-683) -- for any counts, including empty segments and overflowing counts,
and the pipelines built on it must equal the plain single-pass result bit for bit."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _segments(rng, nseg, nx, stride, rows_per_seg, cap, keep_p, k, empty=()):
    """Emission-ordered random segments with ascending disjoint y ranges."""
    segs = []
    for s in range(nseg):
        xs, ys = np.meshgrid(np.arange(nx) * stride, (s * rows_per_seg + np.arange(rows_per_seg)) * stride, indexing="ij")
        cand = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
        keep = rng.random(len(cand)) < (0.0 if s in empty else keep_p)
        c = cand[keep]
        n = len(c)
        segs.append(dict(coords=c, labels=rng.integers(0, 2, n).astype(np.uint8), features=rng.standard_normal((n, 512)).astype(np.float32),
                         logits=rng.standard_normal((n, max(k, 1))).astype(np.float32)))
    return segs


@pytest.mark.parametrize("nseg,nx,rows,keep_p,k,fd", [(1, 7, 5, 0.5, 2, True), (4, 74, 9, 0.4, 2, True), (8, 447, 7, 0.3, 2, False),
                                                      (3, 5, 40, 1.0, 0, True), (5, 23, 11, 0.05, 3, True), (16, 301, 3, 0.6, 2, True)])
def test_merge_equals_lexsort(nseg, nx, rows, keep_p, k, fd):
    from ss25_hierarchical_multiscale_image_classification_b200 import sharding
    rng = np.random.default_rng(nseg * 100 + nx)
    stride = 224
    cap = nx * rows
    segs = _segments(rng, nseg, nx, stride, rows, cap, keep_p, k, empty=(1,) if nseg > 2 else ())
    x = sharding.SurvivorExchange("cuda", cap, k, nx, stride, segs_per_rank=nseg, with_features=fd)
    y_off = 1000 * stride
    for s, g in enumerate(segs):
        n = len(g["coords"])
        pad = rng.integers(0, 5)                                          # input tensors may be larger than the count
        c = torch.zeros((n + pad + 1, 2), dtype=torch.int32, device="cuda")
        c[:n] = torch.from_numpy(g["coords"]).cuda()
        lab = torch.zeros((n + pad + 1,), dtype=torch.uint8, device="cuda")
        lab[:n] = torch.from_numpy(g["labels"]).cuda()
        f = torch.zeros((n + pad + 1, 512), dtype=torch.float32, device="cuda")
        f[:n] = torch.from_numpy(g["features"]).cuda()
        lg = torch.zeros((n + pad + 1, max(k, 1)), dtype=torch.float32, device="cuda")
        lg[:n] = torch.from_numpy(g["logits"]).cuda()
        if c.shape[0] > cap:
            c, lab, f, lg = c[:cap], lab[:cap], f[:cap], lg[:cap]
        count = torch.tensor([n, 0], dtype=torch.int32, device="cuda")
        x.pack(s, c, lab, f, lg, count, y_offset=y_off)
    x.merge()
    out = x.result()
    allc = np.concatenate([g["coords"] for g in segs]) + np.array([0, y_off], np.int32)
    order = np.lexsort((allc[:, 1], allc[:, 0]))
    assert np.array_equal(out["coords"].cpu().numpy(), allc[order])
    assert np.array_equal(out["labels"].cpu().numpy(), np.concatenate([g["labels"] for g in segs])[order])
    if fd:
        assert np.array_equal(out["features"].cpu().numpy(), np.concatenate([g["features"] for g in segs])[order])
    else:
        assert "features" not in out
    if k:
        assert np.array_equal(out["logits"].cpu().numpy(), np.concatenate([g["logits"] for g in segs])[order])
    assert int(x.total[0]) == len(allc)


def test_pack_clamps_an_overflowing_count():
    from ss25_hierarchical_multiscale_image_classification_b200 import sharding
    x = sharding.SurvivorExchange("cuda", 6, 2, 3, 224)
    c = torch.tensor([[0, 0], [0, 224], [224, 0], [448, 224]], dtype=torch.int32, device="cuda")
    count = torch.tensor([9, 0], dtype=torch.int32, device="cuda")          # claims more rows than the tensors hold
    x.pack(0, c, torch.ones(4, dtype=torch.uint8, device="cuda"), torch.ones((4, 512), device="cuda"), torch.ones((4, 2), device="cuda"), count)
    x.merge()
    out = x.result()
    assert torch.equal(out["coords"], c) and int(x.total[0]) == 4
    with pytest.raises(ValueError):
        x.pack(1, c, c[:, 0].to(torch.uint8), torch.ones((4, 512), device="cuda"), torch.ones((4, 2), device="cuda"), count)
    with pytest.raises(ValueError):
        big = torch.zeros((7, 2), dtype=torch.int32, device="cuda")
        x.pack(0, big, torch.zeros(7, dtype=torch.uint8, device="cuda"), torch.ones((7, 512), device="cuda"), torch.ones((7, 2), device="cuda"), count)


def test_exchanged_pipelines_equal_single_pass_bitwise():
    """Row groups through the exchange (resident and host-buffer paths, world = 1) == one plain pass over the level."""
    from oracle import hipac_oracle as orc
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline, sharding
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(10000, 9000, seed=31)
    level = 1
    img, mask = slide.level_array(level), slide.lesion_mask(level)          # 5000 x 4500, P = 896, S = 224
    net = orc.make_resnet18(seed=0, classifier=True)
    packed = features.pack_resnet18(net.state_dict(), "cuda")
    d_img, d_mask = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
    ref = pipeline.process_level(d_img, d_mask, level, packed)
    assert len(ref) > 50
    nx, ny = 23, 21
    x = pipeline.exchange_for_level("cuda", 5000, ny, 224, 2, max_candidates=nx * 6)      # 4 row groups
    assert x.spr == 4
    n_cand = pipeline.process_level_exchanged(d_img, d_mask, level, packed, x)
    out = x.result()
    assert n_cand == ref.candidates
    for k, want in (("coords", ref.coords), ("labels", ref.labels), ("features", ref.features), ("logits", ref.logits)):
        assert torch.equal(out[k], want), k
    # host-buffer path with the exchange
    pipe = pipeline.HostPipeline(4500, 5000, "cuda", with_mask=True, num_classes=2)
    gb = pipeline.upload_group_bounds(0, ny, 3)
    xh = sharding.SurvivorExchange("cuda", nx * (max(b - a for a, b in zip(gb, gb[1:])) + 1), 2, nx, 224, segs_per_rank=3)
    ih, mh = torch.from_numpy(img).pin_memory(), torch.from_numpy(mask).pin_memory()
    for _ in range(2):                                                      # twice: stale mask blocks / buffer reuse
        r = pipeline.process_level_host(ih, mh, level, packed, pipe, groups=3, exchange=xh)
        for got, want in ((r.coords, ref.coords), (r.labels, ref.labels), (r.features, ref.features), (r.logits, ref.logits)):
            assert torch.equal(got, want.cpu())


@pytest.mark.parametrize("world,spr,nx,rows", [(2, 3, 9, 4), (8, 7, 447, 2), (4, 1, 30, 5), (3, 5, 17, 3)])
def test_block_cyclic_merge_equals_lexsort(world, spr, nx, rows):
    """Block-cyclic sharding: stored rank-major (as an all-gather delivers the segments), local segment g of rank r is
    grid-row block g * world + r; the merge must still deliver the whole level in (x, y) order."""
    from ss25_hierarchical_multiscale_image_classification_b200 import sharding
    rng = np.random.default_rng(world * 10 + spr)
    stride, k = 224, 2
    cap = nx * rows
    x = sharding.SurvivorExchange("cuda", cap, k, nx, stride, segs_per_rank=world * spr, cyclic=True, cyclic_world=world)
    allc, alll, allf, allg = [], [], [], []
    for r in range(world):
        for g in range(spr):
            b = g * world + r                                              # global block index = y order
            xs, ys = np.meshgrid(np.arange(nx) * stride, (b * rows + np.arange(rows)) * stride, indexing="ij")
            cand = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
            keep = rng.random(len(cand)) < (0.0 if (r + g) % 4 == 3 else 0.5)
            c = cand[keep]
            n = len(c)
            lab = rng.integers(0, 2, n).astype(np.uint8)
            f = rng.standard_normal((n, 512)).astype(np.float32)
            lg = rng.standard_normal((n, k)).astype(np.float32)
            # the segment's tensors hold LOCAL coordinates; y_offset moves the block to its place in the level
            local = c - np.array([0, b * rows * stride], np.int32) + np.array([0, g * rows * stride], np.int32)
            pad = 2
            tc = torch.zeros((n + pad, 2), dtype=torch.int32, device="cuda")
            tc[:n] = torch.from_numpy(local).cuda()
            tl = torch.zeros((n + pad,), dtype=torch.uint8, device="cuda")
            tl[:n] = torch.from_numpy(lab).cuda()
            tf = torch.zeros((n + pad, 512), device="cuda")
            tf[:n] = torch.from_numpy(f).cuda()
            tg = torch.zeros((n + pad, k), device="cuda")
            tg[:n] = torch.from_numpy(lg).cuda()
            if tc.shape[0] > cap:
                tc, tl, tf, tg = tc[:cap], tl[:cap], tf[:cap], tg[:cap]
            x.pack(r * spr + g, tc, tl, tf, tg, torch.tensor([n, 0], dtype=torch.int32, device="cuda"), y_offset=(b - g) * rows * stride)
            allc.append(c), alll.append(lab), allf.append(f), allg.append(lg)
    x.merge()
    out = x.result()
    allc = np.concatenate(allc)
    order = np.lexsort((allc[:, 1], allc[:, 0]))
    assert np.array_equal(out["coords"].cpu().numpy(), allc[order])
    assert np.array_equal(out["labels"].cpu().numpy(), np.concatenate(alll)[order])
    assert np.array_equal(out["features"].cpu().numpy(), np.concatenate(allf)[order])
    assert np.array_equal(out["logits"].cpu().numpy(), np.concatenate(allg)[order])
