"""Audio backbone: ResNet-18 over the 1-channel log-mel map (models/audio.py:22-39).  Plain torch /
cuDNN — outside the hot path by design.  Written against the torchvision state-dict layout
(``resnet.conv1``, ``resnet.layerN.M.*``, ``resnet.fc``) without depending on torchvision."""
from __future__ import annotations

import torch
from torch import nn

from .video import BasicBlock, Dummy, _stage


class _ResNet18(nn.Module):
    def __init__(self, in_channels=3, num_classes=1000):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = _stage(BasicBlock, 64, 64, 2, 1, nn.BatchNorm2d)
        self.layer2 = _stage(BasicBlock, 64, 128, 2, 2, nn.BatchNorm2d)
        self.layer3 = _stage(BasicBlock, 128, 256, 2, 2, nn.BatchNorm2d)
        self.layer4 = _stage(BasicBlock, 256, 512, 2, 2, nn.BatchNorm2d)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


class AudioModel(nn.Module):
    """AudioModel(pretrained=False): [B,1,64,1001] log-mel -> resnet18 features / 22 logits."""

    def __init__(self, pretrained=False):
        super().__init__()
        if pretrained:
            raise RuntimeError("AudioModel(pretrained=True) needs the torchvision ImageNet checkpoint, which is not available offline; "
                               "load a state dict instead")
        self.resnet = _ResNet18(in_channels=1)
        self.resnet.fc = nn.Sequential(nn.Dropout(0.0), nn.Linear(in_features=512, out_features=22))
        self.modes = ["audio_features"]

    def forward(self, x):
        return self.resnet(x)


__all__ = ["AudioModel", "Dummy"]
