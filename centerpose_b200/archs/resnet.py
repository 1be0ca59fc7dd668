"""ResNet-50 + three dense deconvolutions (the reference's ``'res_50'`` arch,
``lib/models/backbones/msra_resnet.py:64-208``): parameter tree + lowering to fused ops.

Each ``ConvTranspose2d(k4, s2, p1)`` (``msra_resnet.py:168-193``) is lowered to FOUR stride-1 2x2
convolutions, one per output parity (a, b): output row 2i+a only receives kernel rows
ky = 1,3 (a = 0; input rows i, i-1) or ky = 0,2 (a = 1; input rows i+1, i) — so no zero-stuffed
input and no wasted MACs; each sub-conv writes its own sub-lattice of the output through the op's
strided-output fields, with BatchNorm + ReLU folded in.
"""
from __future__ import annotations

import torch
from torch import nn

from .common import StateView, attach, bn, conv
from ..plan import PlanBuilder, Sym, fold_bn

LAYERS = [3, 4, 6, 3]                 # resnet_spec[50] (msra_resnet.py:237-241)
PLANES = [64, 128, 256, 512]
EXP = 4                               # Bottleneck.expansion
DECONV = [256, 256, 256]              # msra_resnet.py:131-135


def build_params(cfg=None) -> nn.Module:
    root = nn.Module()
    attach(root, "conv1", conv(3, 64, 7, 2, 3)); attach(root, "bn1", bn(64))
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip(PLANES, LAYERS), start=1):
        for bi in range(blocks):
            p = f"layer{li}.{bi}"
            stride = 2 if (bi == 0 and li > 1) else 1
            attach(root, p + ".conv1", conv(inplanes, planes, 1)); attach(root, p + ".bn1", bn(planes))
            attach(root, p + ".conv2", conv(planes, planes, 3, stride, 1)); attach(root, p + ".bn2", bn(planes))
            attach(root, p + ".conv3", conv(planes, planes * EXP, 1)); attach(root, p + ".bn3", bn(planes * EXP))
            if bi == 0 and (stride != 1 or inplanes != planes * EXP):
                attach(root, p + ".downsample.0", conv(inplanes, planes * EXP, 1, stride))
                attach(root, p + ".downsample.1", bn(planes * EXP))
            inplanes = planes * EXP
    for i, planes in enumerate(DECONV):
        attach(root, f"deconv_layers.{3 * i}", nn.ConvTranspose2d(inplanes, planes, 4, stride=2, padding=1,
                                                                  output_padding=0, bias=False))
        attach(root, f"deconv_layers.{3 * i + 1}", bn(planes))
        inplanes = planes
    with torch.no_grad():                              # init_weights (msra_resnet.py:210-224)
        for i in range(len(DECONV)):
            nn.init.normal_(root.deconv_layers._modules[str(3 * i)].weight, std=0.001)
    return root


def _deconv_bn_relu(pb: PlanBuilder, P: StateView, x: Sym, wkey: str, bnkey: str) -> Sym:
    """ConvTranspose2d(k4,s2,p1) + BN + ReLU (msra_resnet.py:168-193) -> four parity 2x2 convs (PlanBuilder.deconv_k4s2)."""
    w_t = P(wkey + ".weight").float()                  # (Cin, Cout, 4, 4)
    w_full, b = fold_bn(w_t.permute(1, 0, 2, 3).contiguous(), None, P.bn(bnkey))   # (Cout, Cin, 4, 4)
    return pb.deconv_k4s2(x, w_full, b, relu=True)


def lower(pb: PlanBuilder, P: StateView, x: Sym) -> Sym:
    """PoseResNet.forward (msra_resnet.py:195-208) -> (256 ch, stride 4)."""
    w, b = fold_bn(P("conv1.weight"), None, P.bn("bn1"))
    t = pb.stem(x, w, b, 7, 2, 3, relu=True)
    t = pb.maxpool(t, 3, 2, 1)
    for li, (planes, blocks) in enumerate(zip(PLANES, LAYERS), start=1):
        for bi in range(blocks):
            p = f"layer{li}.{bi}"
            stride = 2 if (bi == 0 and li > 1) else 1
            if P.has(p + ".downsample.0.weight"):
                wd, bd = fold_bn(P(p + ".downsample.0.weight"), None, P.bn(p + ".downsample.1"))
                residual = pb.conv([t], wd, bd, stride=stride, pad=0, relu=False)
            else:
                residual = t
            w1, b1 = fold_bn(P(p + ".conv1.weight"), None, P.bn(p + ".bn1"))
            w2, b2 = fold_bn(P(p + ".conv2.weight"), None, P.bn(p + ".bn2"))
            w3, b3 = fold_bn(P(p + ".conv3.weight"), None, P.bn(p + ".bn3"))
            u = pb.conv([t], w1, b1, stride=1, pad=0, relu=True)
            u = pb.conv([u], w2, b2, stride=stride, pad=1, relu=True)
            t = pb.conv([u], w3, b3, stride=1, pad=0, relu=True, res=residual)
    for i in range(len(DECONV)):
        t = _deconv_bn_relu(pb, P, t, f"deconv_layers.{3 * i}", f"deconv_layers.{3 * i + 1}")
    return t
