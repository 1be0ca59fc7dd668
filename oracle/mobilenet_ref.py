"""ORACLE (test infrastructure, not product): functional torch-CPU fp32 restatement of the
reference's MobileNetV3-Large + DCN-IDAUp backbone (``'mobilenetv3'`` arch), evaluated from a
reference-format ``state_dict``.

Follows ``/root/reference/lib/models/backbones/mobilenet/mobilenetv3.py``:
  * ``:84-93``   hswish / hsigmoid                 -> :func:`hswish`, :func:`hsigmoid`
  * ``:96-111``  SeModule.forward                  -> :func:`_se`
  * ``:114-147`` Block (expand 1x1 -> depthwise k x k -> project 1x1 [-> SE] [+ shortcut])  -> :func:`_block`
  * ``:162-195`` the block table (kernel, in, expand, out, non-linearity, SE, stride)        -> :data:`BLOCKS`
  * ``:62-81``   IDAUp.forward (DCN proj -> depthwise deconv -> + previous -> DCN node)     -> ``dla_ref._ida``
  * ``:213-227`` MobileNetV3.forward               -> :func:`mobilenetv3_backbone`
Pinned by ``oracle/make_golden.py`` against the reference module (``tests/golden/mbv3_*.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dla_ref import _bn, _ida

# (group, kernel, stride, nonlinearity) per Block, in order; channel counts come from the weights.
BLOCKS = (
    [("bneck0", 3, 1, "relu"), ("bneck0", 3, 2, "relu"), ("bneck0", 3, 1, "relu")]
    + [("bneck1", 5, 2, "relu"), ("bneck1", 5, 1, "relu"), ("bneck1", 5, 1, "relu")]
    + [("bneck2", 3, 2, "hswish")] + [("bneck2", 3, 1, "hswish")] * 5 + [("bneck2", 5, 1, "hswish")]
    + [("bneck3", 5, 2, "hswish"), ("bneck3", 5, 1, "hswish")]
)


def hswish(x):
    return x * F.relu6(x + 3.0) / 6.0


def hsigmoid(x):
    return F.relu6(x + 3.0) / 6.0


def _act(x, kind):
    return F.relu(x) if kind == "relu" else hswish(x)


def _se(sd, x, p):
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.relu(_bn(sd, F.conv2d(s, sd[p + ".se.1.weight"]), p + ".se.2"))
    s = hsigmoid(_bn(sd, F.conv2d(s, sd[p + ".se.4.weight"]), p + ".se.5"))
    return x * s


def _block(sd, x, p, k, stride, nl):
    out = _act(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"]), p + ".bn1"), nl)
    w2 = sd[p + ".conv2.weight"]
    out = _act(_bn(sd, F.conv2d(out, w2, stride=stride, padding=k // 2, groups=w2.shape[0]), p + ".bn2"), nl)
    out = _bn(sd, F.conv2d(out, sd[p + ".conv3.weight"]), p + ".bn3")
    if (p + ".se.se.1.weight") in sd:
        out = _se(sd, out, p + ".se")
    if stride == 1:
        sc = x
        if (p + ".shortcut.0.weight") in sd:
            sc = _bn(sd, F.conv2d(x, sd[p + ".shortcut.0.weight"]), p + ".shortcut.1")
        out = out + sc
    return out


def mobilenetv3_backbone(sd, x, p="backbone_model"):
    out = hswish(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"], stride=2, padding=1), p + ".bn1"))
    feats = {}
    idx = {}
    for group, k, stride, nl in BLOCKS:
        i = idx.get(group, 0); idx[group] = i + 1
        out = _block(sd, out, f"{p}.{group}.{i}", k, stride, nl)
        feats[group] = out
    out3 = hswish(_bn(sd, F.conv2d(feats["bneck3"], sd[p + ".conv2.weight"]), p + ".bn2"))
    y = [feats["bneck0"], feats["bneck1"], feats["bneck2"], out3]
    _ida(sd, y, p + ".ida_up", 0, len(y))
    return y[-1]
