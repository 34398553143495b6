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
        """``pretrained=True`` (what the reference's main.py passes): load ``$MGPROTO_PRETRAINED_DIR/<arch>.pth`` (a
        torchvision state dict) if it exists; otherwise warn and keep the random initialisation -- there is no network
        on the target boxes, and raising here would make ``construct_MGProto(..., pretrained=True)`` unusable."""
        net = ctor(weights=None, **kw)
        if pretrained:
            import os
            import warnings
            import torch
            path = os.path.join(os.environ.get("MGPROTO_PRETRAINED_DIR", ""), ctor.__name__ + ".pth")
            if os.environ.get("MGPROTO_PRETRAINED_DIR") and os.path.exists(path):
                net.load_state_dict(torch.load(path, map_location="cpu"))
            else:
                warnings.warn("mgproto_b200: pretrained=True but no local weights (%s; set MGPROTO_PRETRAINED_DIR): "
                              "backbone keeps its random initialisation" % (path or ctor.__name__ + ".pth"))
        return kind(net)
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


def conv_info(features: nn.Module):
    """(kernel sizes, strides, paddings) of the backbone's main path, the input of the receptive-field walk (ref
    models/*_features.py ``conv_info``): every Conv2d / pooling layer in order, residual ``downsample`` branches
    excluded.  For ResNet / DenseNet the reference's metadata also counts the stem max-pool (3, 2, 1) although its
    forward skips it (models/resnet_features.py:140-142, densenet_features.py:119-121); that quirk is kept so the
    numbers equal the reference's."""
    ks, st, pd = [], [], []

    def one(v):
        return int(v[0] if isinstance(v, (tuple, list)) else v)
    stem_pool_pending = isinstance(features, (ResNetFeatures, DenseNetFeatures))
    for name, m in features.named_modules():
        if "downsample" in name:
            continue
        if isinstance(m, (nn.Conv2d, nn.MaxPool2d, nn.AvgPool2d)):
            ks.append(one(m.kernel_size)); st.append(one(m.stride)); pd.append(one(m.padding))
            if stem_pool_pending and isinstance(m, nn.Conv2d):
                ks.append(3); st.append(2); pd.append(1)
                stem_pool_pending = False
    return ks, st, pd


def proto_layer_rf_info(img_size, kernel_sizes, strides, paddings, prototype_kernel_size=1):
    """[n, jump, receptive-field size, centre of the first field] of the prototype layer (ref
    utils/receptive_field.py:111-141): the standard recurrence n' = floor((n + 2p - k)/s) + 1, j' = j s,
    r' = r + (k - 1) j, start' = start + ((k - 1)/2 - p) j, followed by the prototype kernel as a VALID, stride-1 layer."""
    import math
    n, j, r, start = img_size, 1, 1, 0.5
    for k, s, p in zip(kernel_sizes, strides, paddings):
        n = math.floor((n + 2 * p - k) / s) + 1
        r = r + (k - 1) * j
        start = start + ((k - 1) / 2 - p) * j
        j = j * s
    k = prototype_kernel_size
    n = math.ceil(float(n - k + 1))
    r = r + (k - 1) * j
    start = start + ((k - 1) / 2) * j
    return [n, j, r, start]


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
