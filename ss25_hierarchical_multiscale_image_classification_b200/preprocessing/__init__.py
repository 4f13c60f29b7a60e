"""Patch extraction (the reference's ``src/preprocessing`` / ``extract_patches`` home)."""
from .extract import extract_patches, extract_patches_per_slide, parse_xml_mask, scan_slide  # noqa: F401
from .tensor_api import (PatchBatch, PendingPatchBatch, extract_patches_enqueue, extract_patches_tensor,  # noqa: F401
                         grid_shape, patch_and_stride)
