"""Model wrappers (the reference's ``src/models``)."""
from .resnet import ResNet18Classifier, ResNet18ClassifierSIMCLR, ResNet18FeatureExtractor, UnifiedResNet  # noqa: F401
