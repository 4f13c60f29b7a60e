"""Host-side logic of the pipelined upload (no GPU): the sparse lesion-mask upload must leave the staging mask equal to
the host mask of the CURRENT step for any sequence of masks, and the tapered row groups must partition the grid."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from ss25_hierarchical_multiscale_image_classification_b200 import pipeline


class _CpuPipe(pipeline.HostPipeline):
    """HostPipeline whose staging mask lives on the CPU: exercises begin_step / upload_mask_rows without CUDA."""

    def __init__(self, h, w, sparse=True):
        self.mask = torch.zeros((h, w), dtype=torch.uint8)
        self.sparse_mask = sparse
        self._dirty = []
        self.last_h2d_bytes = 0


@pytest.mark.parametrize("h,w", [(1500, 2000), (4096, 4096), (100, 4096), (33, 17), (3333, 6144)])
@pytest.mark.parametrize("sparse", [True, False])
def test_sparse_mask_upload_reproduces_every_mask(h, w, sparse):
    rng = np.random.default_rng(h * 7 + w)
    pipe = _CpuPipe(h, w, sparse)
    for step in range(6):
        m = np.zeros((h, w), np.uint8)
        for _ in range(int(rng.integers(0, 4))):
            r0, c0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            m[r0:r0 + int(rng.integers(1, 400)), c0:c0 + int(rng.integers(1, 3000))] = int(rng.integers(1, 256))
        if step == 4:
            m[:] = 0                                     # an all-zero mask after a non-zero one: stale rows must vanish
        mh = torch.from_numpy(m)
        pipe.begin_step()
        bounds = [0, h // 3, min(h, 2 * h // 3 + 5), h]
        for g in range(3):
            pipe.upload_mask_rows(mh, bounds[g], bounds[g + 1])
        assert torch.equal(pipe.mask, mh)
        nz_rows = int((mh.amax(dim=1) != 0).sum())
        if sparse:
            assert pipe.last_h2d_bytes <= (nz_rows + 2 * pipeline.HostPipeline.MASK_BLOCK_ROWS * 4) * w
            assert pipe.last_h2d_bytes >= nz_rows * w
        else:
            assert pipe.last_h2d_bytes == h * w


def test_upload_group_bounds_partition_the_rows():
    for i0, i1, groups in [(0, 74, 4), (0, 74, 6), (5, 9, 8), (0, 1, 4), (3, 3, 2), (0, 447, 5)]:
        b = pipeline.upload_group_bounds(i0, i1, groups)
        assert b[0] == i0 and b[-1] == i1 and all(x <= y for x, y in zip(b, b[1:]))
        if i1 - i0 >= 8 * max(1, min(groups, i1 - i0)):
            sizes = [y - x for x, y in zip(b, b[1:])]
            assert sizes[0] >= sizes[-1] > 0            # tapered: the exposed last group is the smallest
