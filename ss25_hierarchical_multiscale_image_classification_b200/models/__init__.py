"""Model wrappers (the reference's ``src/models``)."""
from .mil_classifier import MILAttentionPooling, MILClassifier  # noqa: F401
from .resnet import ResNet18Classifier, ResNet18ClassifierSIMCLR, ResNet18FeatureExtractor, UnifiedResNet  # noqa: F401
