"""Run the UNMODIFIED reference functions on in-memory synthetic slides.

TEST INFRASTRUCTURE (dev container only).  ``/root/reference`` is not shipped
to the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py``
imports this module; it exists to (a) validate ``oracle/hipac_oracle.py``
against the real reference and (b) generate the frozen vectors under
``tests/golden/`` (``tests/golden/make_golden.py``).

The reference cannot be imported as-is here (no openslide / lxml /
matplotlib / skimage, and HuggingFace ``datasets`` shadows ``src/datasets``),
so the recipe from SURVEY.md §8c is applied: ``sys.modules`` stubs for the
missing third-party packages, a fake ``openslide.OpenSlide`` backed by a
``SyntheticSlide``, ``parse_xml_mask`` monkey-patched to return the synthetic
lesion mask, and ``PIL.Image.Image.save`` captured in memory instead of
writing PNGs.  The reference's own code does all grid / padding / label /
tissue / naming decisions (src/main.py:609-732) and the resize + normalise
(src/main.py:812-818, src/datasets/patch_dataset.py:70-81).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """``HIPAC_REFERENCE_ROOT``, else the read-only checkout of the dev container, else the git-ignored copy of the few
    needed files that ``oracle/fetch_reference.py`` leaves under ``baseline/_ref`` (that one travels to the GPU box)."""
    env = os.environ.get("HIPAC_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "src", "main.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "main.py"))


_main_mod = None
_slides: dict[str, object] = {}


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _FakeOpenSlide:
        def __new__(cls, path):
            key = os.path.basename(path).replace(".tif", "")
            if key not in _slides:
                raise IOError(f"no synthetic slide registered for {path}")
            return _slides[key]

    mod("openslide", OpenSlide=_FakeOpenSlide, open_slide=_FakeOpenSlide)

    class XMLSyntaxError(Exception):
        pass

    etree = mod("lxml.etree", XMLSyntaxError=XMLSyntaxError, parse=lambda p: None)
    mod("lxml", etree=etree)
    plt = mod("matplotlib.pyplot")
    mod("matplotlib", pyplot=plt)
    measure = mod("skimage.measure")
    mod("skimage", measure=measure)
    # HF `datasets` (a regular package in the venv) shadows the reference's
    # namespace package src/datasets -> pre-register the reference's directory.
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF_ROOT, "src", "datasets")]
    sys.modules["datasets"] = ds


def load_reference_main():
    """Import ``/root/reference/src/main.py`` once, under stubs."""
    global _main_mod
    if _main_mod is not None:
        return _main_mod
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    _install_stubs()
    for p in (REF_ROOT, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        import main as ref_main  # noqa: E402  (the reference's src/main.py)
    _main_mod = ref_main
    return ref_main


def run_reference_extract_patches(slide, level: int, stride=None, with_mask: bool = True,
                                  keep_images: bool = True, mask_arr=None):
    """Call the reference's ``extract_patches(level=..., stride=...)`` on one synthetic slide.

    Returns a list of ``(name, x, y, label, rgb uint8[P,P,3] | None)`` in the order the
    reference saved them."""
    from PIL import Image

    ref = load_reference_main()
    name = slide.name
    _slides[name] = slide
    saved = []

    def fake_save(self, fp, *a, **k):
        base = os.path.basename(str(fp))
        stem = base[:-4]
        parts = stem.split("_")
        label = 1 if parts[-1] == "tumor" else 0
        y = int(parts[-2][1:])
        x = int(parts[-3][1:])
        saved.append((base, x, y, label, np.array(self) if keep_images else None))

    if mask_arr is None:
        mask_arr = slide.lesion_mask(level) if with_mask else None

    def fake_parse_xml_mask(xml_path, level_dims, sl):
        assert tuple(level_dims) == tuple(sl.level_dimensions[level])
        return Image.fromarray(mask_arr, "L")

    cwd = os.getcwd()
    orig_save = Image.Image.save
    orig_parse = ref.parse_xml_mask
    with tempfile.TemporaryDirectory() as tmp:
        img_dir = os.path.join(tmp, "data", "camelyon16", "train", "img")
        ann_dir = os.path.join(tmp, "data", "camelyon16", "train", "mask", "annotations")
        os.makedirs(img_dir)
        os.makedirs(ann_dir)
        open(os.path.join(img_dir, name + ".tif"), "wb").close()
        if mask_arr is not None:
            open(os.path.join(ann_dir, name + ".xml"), "wb").close()
        try:
            os.chdir(tmp)
            Image.Image.save = fake_save
            ref.parse_xml_mask = fake_parse_xml_mask
            with contextlib.redirect_stdout(io.StringIO()):
                ref.extract_patches(level=level, stride=stride)
        finally:
            Image.Image.save = orig_save
            ref.parse_xml_mask = orig_parse
            os.chdir(cwd)
            _slides.pop(name, None)
    return saved


class _OfflineResNet18Classifier:
    """Stand-in for the reference's ``ResNet18Classifier`` INSIDE ``extract_features`` only: the real constructor calls
    ``models.resnet18(pretrained=True)`` (src/models/resnet.py:63-65), i.e. downloads ImageNet weights, which cannot
    work offline.  The object is never run by ``extract_features`` and none of its weights reach the feature model (the
    key filter at src/main.py:852-859 matches nothing), so an identically shaped random-init module
    is an exact substitute for this call site."""

    def __new__(cls):
        import torch
        import torchvision

        class M(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.model = torchvision.models.resnet18(weights=None)
                self.model.fc = torch.nn.Linear(512, 2)

            def forward(self, x):
                return self.model(x)

        return M()


def run_reference_as_written(slide, level: int, stride=None, with_mask: bool = True, seed: int = 0, mask_arr=None,
                             capture_weights: bool = False):
    """The reference's two stages EXACTLY as its CLI runs them, coupled through PNG files on disk:
    ``extract_patches(level=, stride=)`` (src/main.py:609-732, real ``Image.save``) and then
    ``extract_features(level=)`` (src/main.py:805-894: ``PatchDataset`` + ``DataLoader(batch_size=512, num_workers=8)``
    + ``ResNet18FeatureExtractor`` + ``np.save``), in a temporary working directory.

    Only substitutions: the openslide / lxml / matplotlib / skimage import stubs, ``parse_xml_mask`` returning the
    synthetic mask, and ``ResNet18Classifier`` -> ``_OfflineResNet18Classifier`` (see there).  ``torch.manual_seed(seed)``
    is set before ``extract_features`` so its random-init feature model is reproducible.

    ``capture_weights``: also return the state dict of the ``ResNet18FeatureExtractor`` that ``extract_features`` built (the
    class is wrapped by a subclass that records ``state_dict()`` after construction; behaviour unchanged), so that another
    implementation can be run with the very same random-init weights.

    Returns dict(stage1_s, stage2_s, n_png, features f32[N,512], labels, paths[, weights])."""
    import time

    import torch
    from PIL import Image

    ref = load_reference_main()
    name = slide.name
    _slides[name] = slide
    if mask_arr is None and with_mask:
        mask_arr = slide.lesion_mask(level)

    def fake_parse_xml_mask(xml_path, level_dims, sl):
        return Image.fromarray(mask_arr, "L")

    cwd = os.getcwd()
    orig_parse, orig_cls, orig_fe = ref.parse_xml_mask, ref.ResNet18Classifier, ref.ResNet18FeatureExtractor
    captured = {}

    class _CapturingFeatureExtractor(orig_fe):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            captured["weights"] = {n: v.detach().clone() for n, v in self.state_dict().items()}

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        img_dir = os.path.join(tmp, "data", "camelyon16", "train", "img")
        ann_dir = os.path.join(tmp, "data", "camelyon16", "train", "mask", "annotations")
        os.makedirs(img_dir)
        os.makedirs(ann_dir)
        open(os.path.join(img_dir, name + ".tif"), "wb").close()
        if mask_arr is not None:
            open(os.path.join(ann_dir, name + ".xml"), "wb").close()
        try:
            os.chdir(tmp)
            ref.parse_xml_mask = fake_parse_xml_mask
            ref.ResNet18Classifier = _OfflineResNet18Classifier
            if capture_weights:
                ref.ResNet18FeatureExtractor = _CapturingFeatureExtractor
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                t0 = time.perf_counter()
                ref.extract_patches(level=level, stride=stride)
                t1 = time.perf_counter()
                pdir = os.path.join(tmp, "data", "camelyon16", "patches", f"level_{level}", name)
                n_png = len(os.listdir(pdir)) if os.path.isdir(pdir) else 0
                torch.manual_seed(seed)
                t2 = time.perf_counter()
                if n_png:
                    ref.extract_features(level=level)
                t3 = time.perf_counter()
            out = dict(stage1_s=t1 - t0, stage2_s=t3 - t2, n_png=n_png)
            fpath = os.path.join(tmp, f"patch_features_{level}.npy")
            if os.path.exists(fpath):
                out["features"] = np.load(fpath)
                out["labels"] = np.load(os.path.join(tmp, f"patch_labels_{level}.npy"))
                out["paths"] = [os.path.basename(l.strip()) for l in open(os.path.join(tmp, f"patch_paths_{level}.txt"))]
            else:
                out["features"], out["labels"], out["paths"] = np.zeros((0, 512), np.float32), np.zeros((0,), np.int64), []
        finally:
            ref.parse_xml_mask, ref.ResNet18Classifier, ref.ResNet18FeatureExtractor = orig_parse, orig_cls, orig_fe
            os.chdir(cwd)
            _slides.pop(name, None)
    if capture_weights:
        out["weights"] = captured.get("weights")
    return out


def reference_transform():
    """The transform chain built inline at src/main.py:812-818."""
    from torchvision import transforms
    return transforms.Compose([
        transforms.Resize((224, 224)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])


def run_reference_resize_u8(rgb: np.ndarray) -> np.ndarray:
    """uint8 224x224x3 output of the reference's ``transforms.Resize((224,224))`` on a PIL image."""
    from PIL import Image
    from torchvision import transforms
    return np.array(transforms.Resize((224, 224))(Image.fromarray(rgb, "RGB")))


def run_reference_features(patches_rgb: list[np.ndarray], state_dict, batch: int = 16):
    """Replay src/main.py:861-873 with the reference's own ``PatchDataset`` item transform and
    ``ResNet18FeatureExtractor`` (random init replaced by ``state_dict`` of a seeded
    torchvision resnet18 mapped onto its ``features.N`` key names)."""
    import torch
    from PIL import Image

    ref = load_reference_main()
    tf = reference_transform()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        os.chdir(tmp)
        try:
            model = ref.ResNet18FeatureExtractor()
        finally:
            os.chdir(cwd)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model.eval()
    outs = []
    with torch.no_grad():
        for i in range(0, len(patches_rgb), batch):
            x = torch.stack([tf(Image.fromarray(p, "RGB").convert("RGB")) for p in patches_rgb[i:i + batch]])
            outs.append(model(x))
    return torch.cat(outs).numpy() if outs else np.zeros((0, 512), np.float32)


def torchvision_to_feature_extractor_keys(tv_state: dict) -> dict:
    """Map torchvision resnet18 keys onto ``ResNet18FeatureExtractor``'s ``features.N.*`` names
    (Sequential over ``children()[:-1]``: conv1,bn1,relu,maxpool,layer1..4,avgpool;
    src/models/resnet.py:36)."""
    order = {"conv1": 0, "bn1": 1, "layer1": 4, "layer2": 5, "layer3": 6, "layer4": 7}
    out = {}
    for k, v in tv_state.items():
        head, rest = k.split(".", 1)
        if head in order:
            out[f"features.{order[head]}.{rest}"] = v
    return out
