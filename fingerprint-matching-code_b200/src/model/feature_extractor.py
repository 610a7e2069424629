"""Backbone holder, mirroring ``/root/reference/src/model/feature_extractor.py:6-66``.

The image backbone is stock torchvision / cuDNN and out of scope for the B200 work (BASELINE.json
north_star); this file only keeps the three logical chunks ``Net`` inherits.  The reference downloads
ImageNet weights at construction (``feature_extractor.py:40``); without network access the download is
skipped and the backbone keeps its random initialisation (checkpoints overwrite it anyway).
"""
import torch.nn as nn
from torchvision import models


class ResNet18_base(nn.Module):
    def __init__(self, final_layers: bool = False):
        super().__init__()
        self.node_layers, self.edge_layers, self.final_layers = self.get_backbone()
        if not final_layers:
            self.final_layers = None
        self.backbone_params = list(self.parameters())

    def forward(self, *inputs):
        raise NotImplementedError

    @property
    def device(self):
        return next(self.parameters()).device

    @staticmethod
    def get_backbone():
        try:
            backbone = models.resnet18(weights=models.ResNet18_Weights.IMAGENET1K_V1)
        except Exception:      # no network / no cached weights
            backbone = models.resnet18(weights=None)
        node_layers = nn.Sequential(
            backbone.conv1, backbone.bn1, backbone.relu, backbone.maxpool,
            backbone.layer1, backbone.layer2, backbone.layer3)          # 256 x H/16 x W/16
        edge_layers = nn.Sequential(backbone.layer4)                    # 512 x H/32 x W/32
        final_layers = nn.Sequential(nn.AdaptiveMaxPool2d((1, 1)))
        return node_layers, edge_layers, final_layers


class ResNet18_final(ResNet18_base):
    """ResNet-18 with the final global-pool layer."""
    def __init__(self):
        super().__init__(final_layers=True)


class ResNet18(ResNet18_base):
    """ResNet-18 without the final global-pool layer."""
    def __init__(self):
        super().__init__(final_layers=False)
