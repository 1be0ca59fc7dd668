"""ORACLE (test infrastructure, not product): functional torch-CPU fp32 restatement of the
reference's ResNet-50 + 3 deconv backbone, evaluated from a reference-format ``state_dict``.

Follows ``/root/reference/lib/models/backbones/msra_resnet.py``:
  * ``:80-102``  Bottleneck.forward   -> :func:`_bottleneck`
  * ``:137-150`` _make_layer (stride / downsample placement)
  * ``:168-193`` _make_deconv_layer (ConvTranspose2d k4 s2 p1, no bias, + BN + ReLU)
  * ``:195-208`` PoseResNet.forward   -> :func:`resnet50_backbone`
Pinned by ``oracle/make_golden.py`` against the reference module (``tests/golden/res50_*.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5
LAYERS = [3, 4, 6, 3]


def _bn(sd, x, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, EPS)


def _bottleneck(sd, x, p, stride):
    out = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"]), p + ".bn1"))
    out = F.relu(_bn(sd, F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), p + ".bn2"))
    out = _bn(sd, F.conv2d(out, sd[p + ".conv3.weight"]), p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), p + ".downsample.1")
    return F.relu(out + x)


def resnet50_backbone(sd, x, p="backbone_model"):
    x = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"], stride=2, padding=3), p + ".bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, blocks in enumerate(LAYERS, start=1):
        for bi in range(blocks):
            x = _bottleneck(sd, x, f"{p}.layer{li}.{bi}", 2 if (bi == 0 and li > 1) else 1)
    for i in range(3):
        x = F.conv_transpose2d(x, sd[f"{p}.deconv_layers.{3 * i}.weight"], None, stride=2, padding=1)
        x = F.relu(_bn(sd, x, f"{p}.deconv_layers.{3 * i + 1}"))
    return x
