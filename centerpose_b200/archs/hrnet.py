"""HRNet-W32 / W48 (the reference's ``'hrnet'`` arch, ``lib/models/backbones/pose_higher_hrnet.py:237-489``;
``experiments/hrnet_w{32,48}_512.yaml``): parameter tree + lowering to fused ops.

Lowering of one ``HighResolutionModule`` output ``y_i = relu(sum_j f_ij(x_j))`` (``:215-233``): the sum is
never materialised as separate tensors — the running sum starts at the identity term ``x_i``, each
down-sampling chain's LAST 3x3/s2 conv adds it as its fused residual, each up-sampling term is one
1x1 conv at the LOW resolution followed by a fused nearest-upsample + add, and the ReLU rides on
whichever op contributes the final term.  BatchNorm is folded everywhere.
"""
from __future__ import annotations

from torch import nn

from .common import StateView, attach, bn, conv
from ..plan import PlanBuilder, Sym, fold_bn

# MODEL.EXTRA of experiments/hrnet_w32_512.yaml:77-116 (W48: channels 48/96/192/384)
W32_EXTRA = {
    "STAGE2": {"NUM_MODULES": 1, "NUM_BRANCHES": 2, "BLOCK": "BASIC", "NUM_BLOCKS": [4, 4], "NUM_CHANNELS": [32, 64]},
    "STAGE3": {"NUM_MODULES": 4, "NUM_BRANCHES": 3, "BLOCK": "BASIC", "NUM_BLOCKS": [4, 4, 4],
               "NUM_CHANNELS": [32, 64, 128]},
    "STAGE4": {"NUM_MODULES": 3, "NUM_BRANCHES": 4, "BLOCK": "BASIC", "NUM_BLOCKS": [4, 4, 4, 4],
               "NUM_CHANNELS": [32, 64, 128, 256]},
}


def _extra(cfg):
    model = getattr(cfg, "MODEL", None) if cfg is not None else None
    extra = None
    if model is not None:
        extra = model.get("EXTRA") if hasattr(model, "get") else getattr(model, "EXTRA", None)
    if not extra or "STAGE2" not in extra:
        return W32_EXTRA
    return extra


def feature_channels(cfg) -> int:
    return int(_extra(cfg)["STAGE2"]["NUM_CHANNELS"][0])


def _add_basic(root, p, c):
    attach(root, p + ".conv1", conv(c, c, 3, 1, 1)); attach(root, p + ".bn1", bn(c))
    attach(root, p + ".conv2", conv(c, c, 3, 1, 1)); attach(root, p + ".bn2", bn(c))


def build_params(cfg=None) -> nn.Module:
    extra = _extra(cfg)
    root = nn.Module()
    attach(root, "conv1", conv(3, 64, 3, 2, 1)); attach(root, "bn1", bn(64))
    attach(root, "conv2", conv(64, 64, 3, 2, 1)); attach(root, "bn2", bn(64))
    inplanes = 64
    for k in range(4):                                  # layer1 = 4 Bottlenecks, planes 64 (:250, :392-407)
        p = f"layer1.{k}"
        attach(root, p + ".conv1", conv(inplanes, 64, 1)); attach(root, p + ".bn1", bn(64))
        attach(root, p + ".conv2", conv(64, 64, 3, 1, 1)); attach(root, p + ".bn2", bn(64))
        attach(root, p + ".conv3", conv(64, 256, 1)); attach(root, p + ".bn3", bn(256))
        if k == 0:
            attach(root, p + ".downsample.0", conv(inplanes, 256, 1)); attach(root, p + ".downsample.1", bn(256))
        inplanes = 256
    pre = [256]
    for s in (2, 3, 4):
        st = extra[f"STAGE{s}"]
        if str(st["BLOCK"]).upper() != "BASIC":
            raise NotImplementedError("centerpose_b200 hrnet: only BASIC stage blocks (as in hrnet_w32/w48 yaml)")
        cur = [int(c) for c in st["NUM_CHANNELS"]]
        nb = int(st["NUM_BRANCHES"])
        for i in range(nb):                             # _make_transition_layer (:361-390)
            p = f"transition{s - 1}.{i}"
            if i < len(pre):
                if cur[i] != pre[i]:
                    attach(root, p + ".0", conv(pre[i], cur[i], 3, 1, 1)); attach(root, p + ".1", bn(cur[i]))
            else:
                for k in range(i + 1 - len(pre)):
                    co = cur[i] if k == i - len(pre) else pre[-1]
                    attach(root, f"{p}.{k}.0", conv(pre[-1], co, 3, 2, 1)); attach(root, f"{p}.{k}.1", bn(co))
        n_mod = int(st["NUM_MODULES"])
        for m in range(n_mod):
            multi = not (s == 4 and m == n_mod - 1)     # stage4's last module keeps branch 0 only (:277-278, :421-424)
            for b in range(nb):
                for k in range(int(st["NUM_BLOCKS"][b])):
                    _add_basic(root, f"stage{s}.{m}.branches.{b}.{k}", cur[b])
            for i in range(nb if multi else 1):         # _make_fuse_layers (:170-210)
                for j in range(nb):
                    p = f"stage{s}.{m}.fuse_layers.{i}.{j}"
                    if j > i:
                        attach(root, p + ".0", conv(cur[j], cur[i], 1)); attach(root, p + ".1", bn(cur[i]))
                    elif j < i:
                        for k in range(i - j):
                            co = cur[i] if k == i - j - 1 else cur[j]
                            attach(root, f"{p}.{k}.0", conv(cur[j], co, 3, 2, 1)); attach(root, f"{p}.{k}.1", bn(co))
        pre = cur
    return root


def _n(P: StateView, fmt: str) -> int:
    """How many consecutive indices n exist with some key under ``fmt.format(n)``."""
    n = 0
    keys = P.sd.keys()
    while True:
        pre = P.prefix + fmt.format(n)
        if not any(k.startswith(pre) for k in keys):
            return n
        n += 1


def _cbr(pb, P, x, ckey, bkey, stride, pad, relu, res=None):
    w, b = fold_bn(P(ckey + ".weight"), None, P.bn(bkey))
    return pb.conv([x], w, b, stride=stride, pad=pad, relu=relu, res=res)


def _basic(pb, P, x, p):
    u = _cbr(pb, P, x, p + ".conv1", p + ".bn1", 1, 1, True)
    return _cbr(pb, P, u, p + ".conv2", p + ".bn2", 1, 1, True, res=x)


def _module(pb: PlanBuilder, P: StateView, xs, p):
    nb = len(xs)
    for b in range(nb):
        for k in range(_n(P, p + ".branches." + str(b) + ".{}.")):
            xs[b] = _basic(pb, P, xs[b], f"{p}.branches.{b}.{k}")
    ys = []
    for i in range(_n(P, p + ".fuse_layers.{}.")):
        others = [j for j in range(nb) if j != i]
        acc = xs[i]
        for j in others:                                 # j < i first (ascending), then j > i
            last = j == others[-1]
            q = f"{p}.fuse_layers.{i}.{j}"
            if j < i:
                t = xs[j]
                for k in range(i - j):
                    end = k == i - j - 1
                    t = _cbr(pb, P, t, f"{q}.{k}.0", f"{q}.{k}.1", 2, 1, relu=(last if end else True),
                             res=acc if end else None)
                acc = t
            else:
                t = _cbr(pb, P, xs[j], q + ".0", q + ".1", 1, 0, relu=False)
                acc = pb.upsample_add(t, acc, 2 ** (j - i), relu=last)
        ys.append(acc)
    return ys


def lower(pb: PlanBuilder, P: StateView, x: Sym) -> Sym:
    """PoseHigherResolutionNet.forward (:453-489) -> branch 0 of the last stage (stride 4)."""
    w, b = fold_bn(P("conv1.weight"), None, P.bn("bn1"))
    t = pb.stem(x, w, b, 3, 2, 1, relu=True)
    t = _cbr(pb, P, t, "conv2", "bn2", 2, 1, True)
    for k in range(_n(P, "layer1.{}.")):
        p = f"layer1.{k}"
        res = _cbr(pb, P, t, p + ".downsample.0", p + ".downsample.1", 1, 0, False) if P.has(p + ".downsample.0.weight") else t
        u = _cbr(pb, P, t, p + ".conv1", p + ".bn1", 1, 0, True)
        u = _cbr(pb, P, u, p + ".conv2", p + ".bn2", 1, 1, True)
        t = _cbr(pb, P, u, p + ".conv3", p + ".bn3", 1, 0, True, res=res)
    ys = [t]
    for s in (2, 3, 4):
        nb = _n(P, f"stage{s}.0.branches." + "{}.")
        xs = []
        for i in range(nb):                              # transitions (:457-476)
            q = f"transition{s - 1}.{i}"
            if i < len(ys):
                xs.append(_cbr(pb, P, ys[i], q + ".0", q + ".1", 1, 1, True) if P.has(q + ".0.weight") else ys[i])
            else:
                u = ys[-1]
                for k in range(i + 1 - len(ys)):
                    u = _cbr(pb, P, u, f"{q}.{k}.0", f"{q}.{k}.1", 2, 1, True)
                xs.append(u)
        for m in range(_n(P, f"stage{s}." + "{}.")):
            xs = _module(pb, P, xs, f"stage{s}.{m}")
        ys = xs
    return ys[0]
