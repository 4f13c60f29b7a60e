"""Stage-1 parity on the GPU: hipac_tile_scan (through the C ABI) vs the CPU oracle and the
golden vectors frozen from the real reference.  Everything here is bit exact."""
import zlib

import numpy as np
import pytest

from conftest import golden_stage1_cases, load_golden, slide_for
from oracle import hipac_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _pitched(a):
    """Device copy of a [H,W,(3)] array whose row pitch is a multiple of 16 bytes (a view into a wider buffer): the
    layout the TMA streaming pass needs, with an arbitrary image width."""
    h, w = a.shape[:2]
    wp = (w + 15) // 16 * 16 + 16
    buf = torch.full((h, wp) + tuple(a.shape[2:]), 77, dtype=torch.uint8, device="cuda")   # padding bytes are garbage
    buf[:, :w] = torch.from_numpy(a).cuda()
    return buf[:, :w]


def _run(level_img, mask, level, stride, mode="auto", layout="nhwc3", row_range=None, pitched=False):
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
    if pitched:
        img = _pitched(level_img)
        m = _pitched(mask) if mask is not None else None
    else:
        img = torch.from_numpy(level_img).cuda()
        m = torch.from_numpy(mask).cuda() if mask is not None else None
    return extract_patches_tensor(img, m, level, stride=stride, layout=layout, want_u8=True, mode=mode,
                                  row_range=row_range)


def _bf16_bits(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


MODES = ["direct", "auto"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", golden_stage1_cases())
def test_golden_reference_vectors(case, mode):
    g = load_golden(f"stage1_{case}.npz")
    slide = slide_for(g)
    level = int(g["level"])
    stride = None if int(g["stride"]) < 0 else int(g["stride"])
    out = _run(slide.level_array(level), slide.lesion_mask(level), level, stride, mode=mode)
    assert np.array_equal(out.coords.cpu().numpy(), g["coords"])
    assert np.array_equal(out.labels.cpu().numpy(), g["labels"])
    u8 = out.images_u8.cpu().numpy()
    crc = np.array([zlib.crc32(im.tobytes()) for im in u8], dtype=np.uint32)
    assert np.array_equal(crc, g["crc32"])
    # normalised bf16 batch == bf16(ToTensor+Normalize) of the exact uint8 image
    want = orc.to_bf16_bits(orc.normalize_u8(u8)) if len(u8) else np.zeros((0, 224, 224, 3), np.uint16)
    assert np.array_equal(_bf16_bits(out.batch), want)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("level,stride,w,h", [(3, None, 1117, 903), (2, None, 1500, 1100), (2, 448, 1500, 1100),
                                               (1, None, 2100, 1900), (1, 896, 2100, 1900), (0, None, 2500, 2300),
                                               (0, 1792, 4000, 3700), (1, 333, 2100, 1900), (3, 100, 700, 500),
                                               (3, None, 100, 90), (2, None, 449, 225)])
def test_random_images_vs_oracle(level, stride, w, h, mode):
    rng = np.random.default_rng(level * 1000 + w)
    img = rng.integers(200, 256, size=(h, w, 3), dtype=np.uint8)          # hovers around the 240 threshold
    img[h // 5: h // 2, w // 6: w // 2] = rng.integers(0, 256, size=(h // 2 - h // 5, w // 2 - w // 6, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[h // 4: h // 4 + 37, w // 3: w // 3 + 55] = 255
    mask[h - 1, w - 1] = 1                                                 # last pixel, value 1 (>0 counts)
    want = orc.extract_patches_oracle(img, mask, level, stride=stride)
    out = _run(img, mask, level, stride, mode=mode)
    assert out.candidates == want["candidates"]
    assert np.array_equal(out.coords.cpu().numpy(), want["coords"])
    assert np.array_equal(out.labels.cpu().numpy(), want["labels"])
    assert np.array_equal(out.images_u8.cpu().numpy(), want["images"])


@pytest.mark.parametrize("mode", MODES)
def test_s2d16_layout_is_a_permutation_of_nhwc3(mode):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(700, 900, 3), dtype=np.uint8)
    a = _run(img, None, 2, None, mode=mode, layout="nhwc3")
    b = _run(img, None, 2, None, mode=mode, layout="s2d16")
    assert len(a) == len(b) > 0
    x = a.batch.view(torch.int16).reshape(len(a), 112, 2, 112, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(len(a), 112, 112, 12)
    y = b.batch.view(torch.int16)
    assert tuple(y.shape[1:]) == (112, 115, 16)                             # explicit W padding: 2 left, 1 right
    assert torch.equal(y[:, :, 2:114, :12], x)
    assert int(y[..., 12:].abs().max()) == 0
    assert int(y[:, :, :2].abs().max()) == 0 and int(y[:, :, 114:].abs().max()) == 0
    assert int(a.labels.sum()) == 0                                        # no mask -> all "normal"


@pytest.mark.parametrize("mode", MODES)
def test_row_range_shards_concatenate_to_full(mode):
    rng = np.random.default_rng(3)
    img = rng.integers(100, 256, size=(1300, 1000, 3), dtype=np.uint8)
    full = _run(img, None, 2, None, mode=mode)
    ny = (1300 + 223) // 224
    parts = [_run(img, None, 2, None, mode=mode, row_range=r) for r in [(0, 2), (2, 2), (2, 5), (5, ny)]]
    coords = torch.cat([p.coords for p in parts]).cpu().numpy()
    imgs = torch.cat([p.images_u8 for p in parts]).cpu().numpy()
    order = np.lexsort((coords[:, 1], coords[:, 0]))                       # canonical (x, y) order
    assert np.array_equal(coords[order], full.coords.cpu().numpy())
    assert np.array_equal(imgs[order], full.images_u8.cpu().numpy())
    assert sum(p.candidates for p in parts) == full.candidates


def test_pitched_input_and_errors():
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(500, 611, 3), dtype=np.uint8)
    big = torch.zeros((500, 640, 3), dtype=torch.uint8, device="cuda")
    big[:, :611] = torch.from_numpy(img).cuda()
    view = big[:, :611]                                                    # row pitch 1920 bytes, W = 611
    out = extract_patches_tensor(view, None, 3, want_u8=True, layout=None)
    want = orc.extract_patches_oracle(img, None, 3)
    assert np.array_equal(out.coords.cpu().numpy(), want["coords"])
    assert np.array_equal(out.images_u8.cpu().numpy(), want["images"])
    with pytest.raises(ValueError):
        extract_patches_tensor(view.float(), None, 3)
    with pytest.raises(ValueError):
        extract_patches_tensor(view, None, 3, row_range=(0, 99))
    with pytest.raises(RuntimeError):
        extract_patches_tensor(view, None, 3, capacity=1)


@pytest.mark.parametrize("level,stride", [(0, None), (1, None), (2, None), (3, None), (0, 1792), (1, 896), (2, 448), (1, 448), (0, 32 * 8)])
def test_fused_equals_direct_bitwise(level, stride):
    """The read-once path and the per-patch path are two implementations of the same function."""
    rng = np.random.default_rng(level + 17)
    h, w = 2600 + 37 * level, 3100 + 11 * level
    img = rng.integers(150, 256, size=(h, w, 3), dtype=np.uint8)
    img[300:1900, 500:2500] = rng.integers(0, 256, size=(1600, 2000, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[700:720, 900:2000] = 7
    a = _run(img, mask, level, stride, mode="direct", layout="s2d16")
    b = _run(img, mask, level, stride, mode="fused", layout="s2d16")
    assert a.candidates == b.candidates and len(a) == len(b) > 0
    assert torch.equal(a.coords, b.coords) and torch.equal(a.labels, b.labels)
    assert torch.equal(a.images_u8, b.images_u8)
    assert torch.equal(a.batch.view(torch.int16), b.batch.view(torch.int16))


def test_fused_mode_rejects_unaligned_stride():
    img = np.full((1000, 1000, 3), 100, np.uint8)
    with pytest.raises(RuntimeError, match="fused scan needs"):
        _run(img, None, 1, 300, mode="fused")


@pytest.mark.parametrize("level,stride,w,h", [(0, None, 3100, 2600), (1, None, 3111, 2637), (2, None, 3122, 2674), (0, 1792, 4000, 3700),
                                               (1, 896, 2100, 1900), (2, 448, 1500, 1100), (1, 448, 2100, 1900), (0, 256, 2500, 2300),
                                               (2, None, 449, 225), (1, None, 100, 90), (0, None, 8, 3000), (2, 32, 700, 600),
                                               (0, None, 5953, 1800), (0, None, 5955, 1796), (1, None, 2977, 1000)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_streaming_pass_equals_direct_and_legacy_bitwise(level, stride, w, h, with_mask):
    """The TMA streaming pass (one read: cell sums + planes) vs the per-patch path and the cp.async kernels, on
    16-byte-pitched images of arbitrary width (strip edges, white padding, partial cells, tiny images)."""
    rng = np.random.default_rng(level * 100 + w)
    img = rng.integers(150, 256, size=(h, w, 3), dtype=np.uint8)
    img[h // 8: 3 * h // 4, w // 6: 5 * w // 6] = rng.integers(0, 256, size=(3 * h // 4 - h // 8, 5 * w // 6 - w // 6, 3), dtype=np.uint8)
    mask = None
    if with_mask:
        mask = np.zeros((h, w), np.uint8)
        mask[h // 3: h // 3 + 20, w // 4: w // 2] = 7
        mask[h - 1, w - 1] = 1
    a = _run(img, mask, level, stride, mode="direct", layout="s2d16")
    b = _run(img, mask, level, stride, mode="fused", layout="s2d16", pitched=True)
    c = _run(img, mask, level, stride, mode="fused_legacy", layout="s2d16", pitched=True)
    for o in (b, c):
        assert a.candidates == o.candidates and len(a) == len(o)
        assert torch.equal(a.coords, o.coords) and torch.equal(a.labels, o.labels)
        assert torch.equal(a.images_u8, o.images_u8)
        assert torch.equal(a.batch.view(torch.int16), o.batch.view(torch.int16))


def test_streaming_pass_row_range_shards():
    """Row-range shards of the streaming pass concatenate (in canonical order) to the full scan."""
    rng = np.random.default_rng(5)
    h, w = 5000, 2400
    img = rng.integers(120, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[2000:2100, 300:900] = 255
    full = _run(img, mask, 1, None, mode="fused", pitched=True)
    ny = (h + 223) // 224
    parts = [_run(img, mask, 1, None, mode="fused", pitched=True, row_range=r) for r in [(0, 5), (5, 6), (6, ny)]]
    coords = torch.cat([q.coords for q in parts]).cpu().numpy()
    order = np.lexsort((coords[:, 1], coords[:, 0]))
    assert np.array_equal(coords[order], full.coords.cpu().numpy())
    assert np.array_equal(torch.cat([q.labels for q in parts]).cpu().numpy()[order], full.labels.cpu().numpy())
    assert np.array_equal(torch.cat([q.images_u8 for q in parts]).cpu().numpy()[order], full.images_u8.cpu().numpy())


def test_pending_scans_resolve_in_any_order():
    """Two scans enqueued back to back share the thread's pinned count buffer: resolving the OLDER one afterwards must
    still return its own count (it falls back to its device counter)."""
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_enqueue, extract_patches_tensor
    rng = np.random.default_rng(11)
    a_img = torch.from_numpy(rng.integers(0, 200, size=(700, 900, 3), dtype=np.uint8)).cuda()       # all tissue
    b_img = torch.from_numpy(np.full((500, 500, 3), 250, np.uint8)).cuda()                           # all background
    pa = extract_patches_enqueue(a_img, None, 3, layout="nhwc3")
    pb = extract_patches_enqueue(b_img, None, 3, layout="nhwc3")
    rb, ra = pb.resolve(), pa.resolve()
    assert len(rb) == 0 and rb.candidates == 9
    want = extract_patches_tensor(a_img, None, 3, layout="nhwc3")
    assert 0 < len(ra) == len(want) <= 20 and ra.candidates == 20 and torch.equal(ra.coords, want.coords)


def test_many_candidates_multi_block_compaction():
    """More than 1024 candidates: the compaction runs as several CTAs (per-block totals + per-block scatter) and must
    still emit survivors in the reference's x-outer / y-inner order."""
    rng = np.random.default_rng(21)
    h, w = 1900, 2300
    img = rng.integers(215, 256, size=(h, w, 3), dtype=np.uint8)          # hovers around the threshold: ragged keep pattern
    img[200:1500, 300:1900] = rng.integers(0, 256, size=(1300, 1600, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[900:950, 1000:1100] = 255
    for stride, mode in ((32, "auto"), (32, "direct"), (40, "auto")):     # 72 x 60 = 4320 / 58 x 48 = 2784 candidates at P = 224
        want = orc.extract_patches_oracle(img, mask, 3, stride=stride, want_images=False)
        assert want["candidates"] > 2048
        out = _run(img, mask, 3, stride, mode=mode, layout=None)
        assert out.candidates == want["candidates"]
        assert np.array_equal(out.coords.cpu().numpy(), want["coords"])
        assert np.array_equal(out.labels.cpu().numpy(), want["labels"])
    # capacity smaller than the survivor count: the count is still the true one, only `capacity` rows are written
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_enqueue
    pend = extract_patches_enqueue(torch.from_numpy(img).cuda(), None, 3, stride=32, layout=None, capacity=1500)
    n, n_c = (int(v) for v in pend.count.cpu().tolist())
    want = orc.extract_patches_oracle(img, None, 3, stride=32, want_images=False)
    assert n == len(want["coords"]) > 1500 and n_c == want["candidates"]
    assert np.array_equal(pend.coords.cpu().numpy()[:1500], want["coords"][:1500])


def test_unaligned_pitch_is_repitched_onto_the_streaming_pass_and_raw_abi_fallback():
    """A contiguous [H, W, 3] tensor with 3*W % 16 != 0: the Python API re-pitches it so that the read-once streaming
    pass runs (no silent two-pass fallback); the raw C ABI with the unaligned pitch still works through the cp.async
    kernels.  Both equal the direct path bit for bit."""
    from ss25_hierarchical_multiscale_image_classification_b200 import _lib
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import alloc_level_image, upload_level_rows
    rng = np.random.default_rng(33)
    h, w, level = 1300, 1501, 2
    img = rng.integers(100, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[400:420, 100:900] = 3
    ref = _run(img, mask, level, None, mode="direct", layout="s2d16")
    _lib.profile(True)
    out = _run(img, mask, level, None, mode="auto", layout="s2d16")
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile(False)
    assert "scan_planes" in prof and "downsample_planes" not in prof, prof.keys()
    assert torch.equal(ref.coords, out.coords) and torch.equal(ref.labels, out.labels)
    assert torch.equal(ref.images_u8, out.images_u8) and torch.equal(ref.batch.view(torch.int16), out.batch.view(torch.int16))
    # pitched staging filled by the 2-D upload: same result
    dimg, dmask = alloc_level_image(h, w, "cuda"), alloc_level_image(h, w, "cuda", channels=1)
    assert dimg.stride(0) % 16 == 0 and dmask.stride(0) % 16 == 0
    upload_level_rows(dimg, img[:700]), upload_level_rows(dimg, img[700:], 700)
    upload_level_rows(dmask, mask)
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
    up = extract_patches_tensor(dimg, dmask, level, layout="s2d16", want_u8=True)
    assert torch.equal(ref.coords, up.coords) and torch.equal(ref.images_u8, up.images_u8) and torch.equal(ref.labels, up.labels)
    # raw ABI, contiguous unaligned pitch, fused mode: cp.async kernels
    l = _lib.lib()
    t, m = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
    P, S = 448, 224
    ny = (h + S - 1) // S
    cap = ((w + S - 1) // S) * ny
    coords = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
    labels = torch.empty((cap,), dtype=torch.uint8, device="cuda")
    u8 = torch.empty((cap, 224, 224, 3), dtype=torch.uint8, device="cuda")
    count = torch.zeros((2,), dtype=torch.int32, device="cuda")
    wsb = l.hipac_tile_scan_workspace_bytes(h, w, P, S, 0, ny, _lib.SCAN_FUSED)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")
    rc = l.hipac_tile_scan(t.data_ptr(), h, w, 3 * w, m.data_ptr(), w, P, S, 0, ny, coords.data_ptr(), labels.data_ptr(), u8.data_ptr(),
                           None, 0, count.data_ptr(), cap, ws.data_ptr(), wsb, _lib.SCAN_FUSED, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, l.hipac_last_error()
    n = int(count[0])
    assert n == len(ref) and torch.equal(coords[:n], ref.coords) and torch.equal(u8[:n], ref.images_u8)
