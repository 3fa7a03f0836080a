"""Template encoder: image [B,3,256,256] -> latent [B,8,32,32].

Host-framework (PyTorch/cuDNN) module in this round: SURVEY.md section 8 keeps the encoder
as a library call (row a3) and ranks a custom kernel path "next" (row f1) -- it runs once
per query and once per reference, against N UNet forwards per query.

Mirrors reference `FeatureExtractor` (src/model/encoder/template.py:24-53): ResNet-50
without max-pool and with layer4 at stride 1 (src/model/encoder/resnet.py:93-152, so the
total stride is 8), eval-mode BatchNorm, then ReLU -> 1x1(2048->256) -> ReLU -> 1x1(256->D).
Parameter names equal the reference's `backbone.*` / `projector.*` keys so its
state_dict loads unchanged; the duplicate `encoder.{0,1}.*` aliases the reference
registers (template.py:40) are accepted and ignored.
"""
import torch
from torch import nn
import torch.nn.functional as F


class _Bottleneck(nn.Module):
    # resnet.py:55-91
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(out + x)


class _Backbone(nn.Module):
    # resnet.py:93-152 with use_avg_pooling_and_fc=False (no max-pool, no avgpool/fc in forward)
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.inplanes = 64
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 1)
        self.fc = nn.Linear(2048, 1)   # present in the reference state_dict, unused

    def _make_layer(self, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes * 4))
        layers = [_Bottleneck(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        layers += [_Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class FeatureExtractor(nn.Module):
    """Same constructor/attributes as the reference (template.py:25-45)."""

    def __init__(self, descriptor_size=8, threshold=0.2, normalize=False, **kwargs):
        super().__init__()
        self.latent_dim = descriptor_size
        self.normalize = normalize
        self.threshold = threshold
        self.name = "template"
        self.backbone = _Backbone()
        self.projector = nn.Sequential(
            nn.ReLU(inplace=False), nn.Conv2d(2048, 256, 1, bias=False),
            nn.ReLU(inplace=False), nn.Conv2d(256, descriptor_size, 1, bias=False))
        self.eval()

    def load_state_dict(self, state_dict, strict=True):
        own = {k: v for k, v in state_dict.items() if not k.startswith("encoder.")}
        return super().load_state_dict(own, strict=strict)

    @torch.no_grad()
    def encode_image(self, image, mode=None):
        """template.py:47-53.  fp32 with TF32 off by default, so the latent matches the
        reference's fp32 path; `mode` is accepted and ignored as in the reference."""
        p = next(self.parameters())
        image = image.to(device=p.device, dtype=p.dtype)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            feat = self.projector(self.backbone(image))
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        if self.normalize:
            feat = F.normalize(feat, dim=1)
        return feat.float()
