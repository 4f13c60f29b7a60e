"""ABMIL head (csrc/mil.cu through the C ABI) vs the reference module's fp32 arithmetic
(src/models/mil_classifier.py:5-45: tanh(V x) -> U -> softmax over the instances -> weighted sum -> MLP)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _reference_forward(model, bag):
    """The reference's forward, fp64 on the CPU (ground truth for the fp32 kernels)."""
    m = model.double().cpu()
    with torch.no_grad():
        out = m(bag.double().cpu())
    model.float()
    return out


@pytest.mark.parametrize("n", [1, 7, 32, 33, 1000, 20011])
@pytest.mark.parametrize("pooling", ["attention", "mean", "max"])
def test_mil_classifier_matches_reference_arithmetic(n, pooling):
    from ss25_hierarchical_multiscale_image_classification_b200.models import MILClassifier
    torch.manual_seed(n)
    model = MILClassifier(512, num_classes=2, pooling=pooling)
    with torch.no_grad():                                   # sharpen the attention so that the softmax is not flat
        if pooling == "attention":
            model.aggregator.attn_U.weight.mul_(8.0)
    bag = torch.relu(torch.randn(n, 512)) * torch.rand(n, 1) * 3.0        # feature-like: non-negative, varied norms
    want_logits, want_attn = _reference_forward(model, bag)
    model = model.cuda()
    logits, attn = model(bag.cuda())
    torch.cuda.synchronize()
    assert logits.shape == (2,) and logits.dtype == torch.float32
    assert torch.allclose(logits.cpu().double(), want_logits, rtol=2e-4, atol=2e-5), (logits.cpu(), want_logits)
    if pooling == "attention":
        assert attn.shape == (n, 1)
        assert torch.allclose(attn.cpu().double(), want_attn, rtol=2e-4, atol=1e-7)
        assert abs(float(attn.sum()) - 1.0) < 1e-4
    else:
        assert attn is None and want_attn is None


def test_device_count_and_state_dict_contract():
    """The instance count may come from device memory (chained behind the tile scan); state-dict keys are the reference's."""
    from ss25_hierarchical_multiscale_image_classification_b200.models import MILClassifier
    from ss25_hierarchical_multiscale_image_classification_b200.models.mil_classifier import mil_forward
    torch.manual_seed(0)
    model = MILClassifier(512, num_classes=3)
    assert sorted(model.state_dict()) == ["aggregator.attn_U.bias", "aggregator.attn_U.weight", "aggregator.attn_V.bias",
                                          "aggregator.attn_V.weight", "classifier.0.bias", "classifier.0.weight",
                                          "classifier.2.bias", "classifier.2.weight"]
    bag = torch.rand(500, 512)
    want_logits, want_attn = _reference_forward(model, bag[:123])
    model = model.cuda()
    count = torch.tensor([123, 500], dtype=torch.int32, device="cuda")
    logits, attn = model(bag.cuda(), count=count)
    torch.cuda.synchronize()
    assert torch.allclose(logits.cpu().double(), want_logits, rtol=2e-4, atol=2e-5)
    assert torch.allclose(attn[:123].cpu().double(), want_attn, rtol=2e-4, atol=1e-7)
    # empty bag: pooled vector 0 -> logits = W2 relu(b1) + b2
    logits0, _, pooled0 = mil_forward(bag[:0].cuda(), model._packed_for(torch.device("cuda", 0)), 3)
    torch.cuda.synchronize()
    with torch.no_grad():
        want0 = model.classifier(torch.zeros(512, device="cuda"))
    assert float(pooled0.abs().max()) == 0.0 and torch.allclose(logits0, want0, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        mil_forward(bag.cuda().half(), model._packed_for(torch.device("cuda", 0)), 3)


def test_features_to_slide_logits_end_to_end():
    """Tile scan -> ResNet18 features (device count) -> ABMIL head, enqueued without a host round trip."""
    from oracle import hipac_oracle as orc
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
    from ss25_hierarchical_multiscale_image_classification_b200.models import MILClassifier
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(6000, 5200, seed=21)
    img, mask = slide.level_array(2), slide.lesion_mask(2)
    packed = features.pack_resnet18(orc.make_resnet18(seed=0, classifier=True).state_dict(), "cuda")
    torch.manual_seed(1)
    head = MILClassifier(512).cuda()
    seg = pipeline.process_level_enqueue(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), 2, packed)
    logits, attn = head(seg.features, count=seg.count)
    torch.cuda.synchronize()
    n = int(seg.count[0])
    want_logits, want_attn = _reference_forward(head, seg.features[:n].cpu())
    assert n > 5 and torch.allclose(logits.cpu().double(), want_logits, rtol=5e-4, atol=5e-5)
    assert torch.allclose(attn[:n].cpu().double(), want_attn, rtol=5e-4, atol=1e-7)
