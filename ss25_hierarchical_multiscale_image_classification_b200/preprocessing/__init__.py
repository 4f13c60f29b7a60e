"""Patch extraction (the reference's ``src/preprocessing`` / ``extract_patches`` home)."""
from .tensor_api import PatchBatch, extract_patches_tensor, grid_shape, patch_and_stride  # noqa: F401
