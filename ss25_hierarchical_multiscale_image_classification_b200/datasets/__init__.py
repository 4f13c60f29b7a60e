"""Dataset classes (the reference's ``src/datasets``)."""
from .patch_dataset import PatchDataset  # noqa: F401
