"""ORACLE (test infrastructure, not product): functional torch-CPU fp32 restatement of the
reference's HRNet backbone (``'hrnet'`` arch, W32/W48), evaluated from a reference-format ``state_dict``.

Follows ``/root/reference/lib/models/backbones/pose_higher_hrnet.py``:
  * ``:37-53``   BasicBlock.forward, ``:73-95`` Bottleneck.forward        -> :func:`_basic`, :func:`_bottleneck`
  * ``:170-210`` _make_fuse_layers (1x1+BN+nearest-up for j>i, chain of 3x3 s2 for j<i)
  * ``:215-233`` HighResolutionModule.forward                             -> :func:`_module`
  * ``:361-390`` _make_transition_layer, ``:453-489`` PoseHigherResolutionNet.forward -> :func:`hrnet_backbone`
The stage / branch / block structure is read off the state_dict's key names, so W32 and W48
checkpoints both evaluate.  Pinned by ``oracle/make_golden.py`` against the reference module
(``tests/golden/hrnet32_*.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(sd, x, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, EPS)


def _count(sd, fmt):
    n = 0
    while any(k.startswith(fmt.format(n)) for k in sd):
        n += 1
    return n


def _basic(sd, x, p):
    out = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"], padding=1), p + ".bn1"))
    out = _bn(sd, F.conv2d(out, sd[p + ".conv2.weight"], padding=1), p + ".bn2")
    return F.relu(out + x)


def _bottleneck(sd, x, p):
    out = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"]), p + ".bn1"))
    out = F.relu(_bn(sd, F.conv2d(out, sd[p + ".conv2.weight"], padding=1), p + ".bn2"))
    out = _bn(sd, F.conv2d(out, sd[p + ".conv3.weight"]), p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, F.conv2d(x, sd[p + ".downsample.0.weight"]), p + ".downsample.1")
    return F.relu(out + x)


def _fuse_term(sd, x, p, i, j):
    """fuse_layers[i][j] applied to branch j's output."""
    if j > i:
        t = _bn(sd, F.conv2d(x, sd[p + ".0.weight"]), p + ".1")
        return F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
    for k in range(i - j):
        x = _bn(sd, F.conv2d(x, sd[f"{p}.{k}.0.weight"], stride=2, padding=1), f"{p}.{k}.1")
        if k != i - j - 1:
            x = F.relu(x)
    return x


def _module(sd, xs, p):
    nb = len(xs)
    for b in range(nb):
        for k in range(_count(sd, p + ".branches." + str(b) + ".{}.")):
            xs[b] = _basic(sd, xs[b], f"{p}.branches.{b}.{k}")
    n_out = _count(sd, p + ".fuse_layers.{}.")
    ys = []
    for i in range(n_out):
        y = None
        for j in range(nb):
            t = xs[j] if i == j else _fuse_term(sd, xs[j], f"{p}.fuse_layers.{i}.{j}", i, j)
            y = t if y is None else y + t
        ys.append(F.relu(y))
    return ys


def _transition(sd, ys, p, n_next):
    xs = []
    for i in range(n_next):
        if i < len(ys):
            if (f"{p}.{i}.0.weight") in sd:
                xs.append(F.relu(_bn(sd, F.conv2d(ys[i], sd[f"{p}.{i}.0.weight"], padding=1), f"{p}.{i}.1")))
            else:
                xs.append(ys[i])
        else:
            t = ys[-1]
            for k in range(i + 1 - len(ys)):
                t = F.relu(_bn(sd, F.conv2d(t, sd[f"{p}.{i}.{k}.0.weight"], stride=2, padding=1), f"{p}.{i}.{k}.1"))
            xs.append(t)
    return xs


def hrnet_backbone(sd, x, p="backbone_model"):
    x = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"], stride=2, padding=1), p + ".bn1"))
    x = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv2.weight"], stride=2, padding=1), p + ".bn2"))
    for k in range(_count(sd, p + ".layer1.{}.")):
        x = _bottleneck(sd, x, f"{p}.layer1.{k}")
    ys = [x]
    for s in (2, 3, 4):
        nb = _count(sd, f"{p}.stage{s}.0.branches." + "{}.")
        xs = _transition(sd, ys, f"{p}.transition{s - 1}", nb)
        for m in range(_count(sd, f"{p}.stage{s}." + "{}.")):
            xs = _module(sd, xs, f"{p}.stage{s}.{m}")
        ys = xs
    return ys[0]
