"""BASELINE.json-size checks through size-independent properties (the CPU oracle would take minutes here):
two independent GPU implementations agree bit for bit, shards concatenate to the whole, the pipelined host path
equals the resident path, and every level of a pyramid matches the oracle on a size it finishes in seconds."""
import numpy as np
import pytest

from oracle import hipac_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bench_image():
    """The bench workload's level-0 image (configs[1]): 16384 x 16384 RGB + lesion mask."""
    from concurrent.futures import ThreadPoolExecutor
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import make_lesion_mask, make_level
    H = W = 16384
    img = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    msk = torch.empty((H, W), dtype=torch.uint8).pin_memory()

    def fill(r):
        img.numpy()[r:r + 256] = make_level(1234, 0, W, H, r, r + 256)
        msk.numpy()[r:r + 256] = make_lesion_mask(1234, 0, W, H, r, r + 256)

    with ThreadPoolExecutor(16) as ex:
        list(ex.map(fill, range(0, H, 256)))
    return img, msk


def test_bench_workload_fused_equals_direct_and_shards(bench_image):
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
    img, msk = (t.cuda() for t in bench_image)
    a = extract_patches_tensor(img, msk, 0, mode="fused", layout="s2d16")
    b = extract_patches_tensor(img, msk, 0, mode="direct", layout="s2d16")
    assert a.candidates == b.candidates == 74 * 74 and len(a) == len(b) > 1000
    assert torch.equal(a.coords, b.coords) and torch.equal(a.labels, b.labels)
    assert torch.equal(a.batch.view(torch.int16), b.batch.view(torch.int16))
    assert 0 < int(a.labels.sum()) < len(a)
    del b
    parts = [extract_patches_tensor(img, msk, 0, row_range=r, layout="s2d16") for r in [(0, 19), (19, 37), (37, 56), (56, 74)]]
    coords = torch.cat([p.coords for p in parts])
    key = coords[:, 0].long() * (1 << 32) + coords[:, 1].long()
    order = torch.argsort(key)
    assert torch.equal(coords[order], a.coords)
    assert torch.equal(torch.cat([p.batch for p in parts])[order].view(torch.int16), a.batch.view(torch.int16))


def test_pipelined_host_path_equals_resident_path(bench_image):
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
    img_h, msk_h = bench_image
    packed = features.pack_resnet18(orc.make_resnet18(seed=0, classifier=True).state_dict(), "cuda")
    res = pipeline.process_level(img_h.cuda(), msk_h.cuda(), 0, packed)
    pipe = pipeline.HostPipeline(16384, 16384, "cuda", with_mask=True)
    host = pipeline.process_level_host(img_h, msk_h, 0, packed, pipe, groups=4)
    key = host.coords[:, 0].long() * (1 << 32) + host.coords[:, 1].long()
    order = torch.argsort(key)
    assert torch.equal(host.coords[order], res.coords.cpu()) and torch.equal(host.labels[order], res.labels.cpu())
    assert torch.equal(host.features[order], res.features.cpu())          # identical kernels on identical inputs
    assert torch.equal(host.logits[order], res.logits.cpu())
    dense = int(img_h.numel() + msk_h.numel())
    assert int(img_h.numel()) < pipe.last_h2d_bytes < dense               # sparse mask upload: only non-zero row blocks


def test_sparse_mask_upload_tracks_a_changing_mask():
    """The device mask of a reused HostPipeline must equal the host mask of the CURRENT step (stale rows re-zeroed),
    for ragged heights, dense masks and the dense-upload mode."""
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
    rng = np.random.default_rng(3)
    H, W = 1500, 2000                                                    # 1500 = 46 * 32 + 28: ragged last block
    img = torch.from_numpy(rng.integers(0, 230, size=(H, W, 3), dtype=np.uint8)).pin_memory()
    packed = features.pack_resnet18(orc.make_resnet18(seed=0, classifier=True).state_dict(), "cuda")
    masks = []
    for k, (r0, r1) in enumerate([(100, 140), (1480, 1500), (0, 1500), (700, 701)]):
        m = np.zeros((H, W), np.uint8)
        m[r0:r1, 50 + 10 * k: 900] = 1 + k
        masks.append(torch.from_numpy(m).pin_memory())
    for sparse in (True, False):
        pipe = pipeline.HostPipeline(H, W, "cuda", with_mask=True, sparse_mask=sparse)
        for m in masks + masks[:1]:
            host = pipeline.process_level_host(img, m, 3, packed, pipe, groups=3)
            assert torch.equal(pipe.mask.cpu(), m)
            res = pipeline.process_level(img.cuda(), m.cuda(), 3, packed)
            key = host.coords[:, 0].long() * (1 << 32) + host.coords[:, 1].long()
            order = torch.argsort(key)
            assert torch.equal(host.coords[order], res.coords.cpu()) and torch.equal(host.labels[order], res.labels.cpu())
            assert 0 < int(res.labels.sum())


def test_all_levels_pyramid_and_heatmap():
    """configs[2]/[3] shape at oracle-friendly size: every level of one pyramid, features + per-patch heatmap."""
    from ss25_hierarchical_multiscale_image_classification_b200 import features, heatmap, pipeline
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(9000, 7000, seed=42)
    net = orc.make_resnet18(seed=0, classifier=True)
    packed = features.pack_resnet18(net.state_dict(), "cuda")
    for level in (3, 2, 1, 0):
        img, mask = slide.level_array(level), slide.lesion_mask(level)
        want = orc.extract_patches_oracle(img, mask, level, want_images=(level >= 2))
        r = pipeline.process_level(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), level, packed)
        assert r.candidates == want["candidates"]
        assert np.array_equal(r.coords.cpu().numpy(), want["coords"]) and np.array_equal(r.labels.cpu().numpy(), want["labels"])
        if level >= 2 and len(r):
            ref_f, ref_l = orc.resnet18_features_fp32(net, want["images"])
            got = r.features.cpu().numpy()
            cos = (got * ref_f).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref_f, axis=1))
            assert cos.min() >= 0.9995 and (np.abs(got - ref_f).max(1) / np.abs(ref_f).max(1)).max() <= 1e-2
            w, h = slide.level_dimensions[level]
            grid = heatmap.heatmap(r.coords, r.logits, w, h, want["stride"], fill=-1.0).cpu().numpy()
            ref_p = torch.softmax(torch.from_numpy(ref_l), 1)[:, 1].numpy()
            for (x, y), pr in zip(want["coords"], ref_p):
                assert abs(grid[y // 224, x // 224] - pr) < 2e-2
            assert (grid >= 0).sum() == len(r)


def test_process_level_in_row_groups_equals_single_shot():
    """Levels too large for one worst-case batch are processed in candidate-row groups: same result, same order."""
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(20000, 16000, seed=9)
    img = torch.from_numpy(slide.level_array(3)).cuda()
    msk = torch.from_numpy(slide.lesion_mask(3)).cuda()
    packed = features.pack_resnet18(orc.make_resnet18(seed=0, classifier=True).state_dict(), "cuda")
    one = pipeline.process_level(img, msk, 3, packed)
    many = pipeline.process_level(img, msk, 3, packed, max_candidates=40)      # 12 candidate columns -> 3 rows per group
    assert many.candidates == one.candidates and len(one) > 10
    assert torch.equal(many.coords, one.coords) and torch.equal(many.labels, one.labels)
    assert torch.equal(many.features, one.features) and torch.equal(many.logits, one.logits)
