"""Hardware-behaviour test: UMMA smem descriptors shifted by whole rows inside a TMA-written swizzled tile."""
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("swz", [128, 32])
@pytest.mark.parametrize("shift", [0, 1, 2, 3, 7, 8, 9, 13, 57, 58, 59, 115, 116, 128])
def test_shifted_descriptor(swz, shift):
    from ss25_hierarchical_multiscale_image_classification_b200 import _lib
    l = _lib.lib()
    k = swz // 2
    g = torch.Generator(device="cuda").manual_seed(shift * 7 + swz)
    A = torch.randn((256, k), generator=g, device="cuda").bfloat16()
    B = torch.randn((64, k), generator=g, device="cuda").bfloat16()
    D = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    rc = l.hipac_debug_umma_shift(A.data_ptr(), B.data_ptr(), D.data_ptr(), shift, swz, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "hipac_debug_umma_shift")
    torch.cuda.synchronize()
    ref = A[shift:shift + 128].float() @ B.float().t()
    assert torch.allclose(D, ref, rtol=1e-3, atol=1e-3), float((D - ref).abs().max())
