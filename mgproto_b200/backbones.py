"""Stock-PyTorch feature extractors for ``construct_MGProto`` (ref models/*_features.py).

Out of the hot path ("stays PyTorch/cuDNN").  Built on torchvision's
architectures with the reference's one structural change: the stem max-pool is skipped
(ResNet / DenseNet) or the last max-pool dropped (VGG) so that a 224x224 image yields a
14x14 map (ref models/resnet_features.py:199, densenet_features.py:116, vgg_features.py:66-68).
Sub-module names follow torchvision (= the reference's), so ``features.*`` checkpoint keys
line up.  ``pretrained=True`` needs weights on disk; there is no network here.
"""
from __future__ import annotations

import torch.nn as nn
import torchvision.models as tvm


class ResNetFeatures(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.conv1, self.bn1, self.relu = net.conv1, net.bn1, net.relu
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))          # stem max-pool deliberately skipped
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class DenseNetFeatures(nn.Module):
    def __init__(self, net):
        super().__init__()
        feats = net.features
        self.features = nn.Sequential()
        for name, mod in feats.named_children():
            if name == "pool0":                          # stem max-pool skipped
                continue
            self.features.add_module(name, mod)
        self.features.add_module("final_relu", nn.ReLU(inplace=True))

    def forward(self, x):
        return self.features(x)


class VGGFeatures(nn.Module):
    def __init__(self, net):
        super().__init__()
        mods = list(net.features.children())
        while isinstance(mods[-1], nn.MaxPool2d):        # final max-pool removed
            mods.pop()
        self.features = nn.Sequential(*mods)

    def forward(self, x):
        return self.features(x)


def _make(kind, ctor):
    def factory(pretrained=False, **kw):
        if pretrained:
            raise RuntimeError("pretrained weights need a local file; load them with load_state_dict "
                               "(no network in this environment)")
        return kind(ctor(weights=None, **kw))
    return factory


base_architecture_to_features = {
    "resnet18": _make(ResNetFeatures, tvm.resnet18), "resnet34": _make(ResNetFeatures, tvm.resnet34),
    "resnet50": _make(ResNetFeatures, tvm.resnet50), "resnet101": _make(ResNetFeatures, tvm.resnet101),
    "resnet152": _make(ResNetFeatures, tvm.resnet152),
    "densenet121": _make(DenseNetFeatures, tvm.densenet121), "densenet161": _make(DenseNetFeatures, tvm.densenet161),
    "densenet169": _make(DenseNetFeatures, tvm.densenet169), "densenet201": _make(DenseNetFeatures, tvm.densenet201),
    "vgg11": _make(VGGFeatures, tvm.vgg11), "vgg11_bn": _make(VGGFeatures, tvm.vgg11_bn),
    "vgg13": _make(VGGFeatures, tvm.vgg13), "vgg13_bn": _make(VGGFeatures, tvm.vgg13_bn),
    "vgg16": _make(VGGFeatures, tvm.vgg16), "vgg16_bn": _make(VGGFeatures, tvm.vgg16_bn),
    "vgg19": _make(VGGFeatures, tvm.vgg19), "vgg19_bn": _make(VGGFeatures, tvm.vgg19_bn),
}


def out_channels(features: nn.Module) -> int:
    """Channels of the map fed to add_on_layers (ref model.py:107-115): the last BatchNorm2d if
    it comes after the last Conv2d (DenseNet's norm5), else the last Conv2d."""
    last = None
    for m in features.modules():
        if isinstance(m, nn.Conv2d):
            last = m.out_channels
        elif isinstance(m, nn.BatchNorm2d):
            last = m.num_features
    if last is None:
        raise ValueError("backbone has no Conv2d/BatchNorm2d to infer channels from")
    return last
