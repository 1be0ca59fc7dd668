"""DLA-34 with deformable-conv IDAUp neck (the reference's ``'dla_34'`` arch,
``lib/models/backbones/pose_dla_dcn.py``): parameter tree + lowering to fused ops.

The network is described once, as data (:data:`LEVELS`, :data:`UP_STAGES`), and walked twice:
``build_params`` registers reference-named parameter holders, ``lower`` emits the op program.
Op numbering in comments follows SURVEY.md Appendix A.
"""
from __future__ import annotations

from torch import nn

from .common import DCNParams, StateView, attach, bilinear_up, bn, conv
from ..plan import PlanBuilder, Sym, fold_bn

CH = [16, 32, 64, 128, 256, 512]          # pose_dla_dcn.py:308-309 dla34 channels
TREE_LEVELS = {2: 1, 3: 2, 4: 2, 5: 1}    # levels of base.level{2..5}  (:307)
FIRST_LEVEL = 2                           # down_ratio 4 (:423)
LAST_LEVEL = 5


# ---------------------------------------------------------------------------- parameters
def _tree_params(root, p, levels, cin, cout, level_root, root_dim=0):
    if root_dim == 0:
        root_dim = 2 * cout
    if level_root:
        root_dim += cin
    if levels == 1:
        for name, ci in (("tree1", cin), ("tree2", cout)):
            attach(root, f"{p}.{name}.conv1", conv(ci, cout, 3, 1, 1))    # stride set at lowering
            attach(root, f"{p}.{name}.bn1", bn(cout))
            attach(root, f"{p}.{name}.conv2", conv(cout, cout, 3, 1, 1))
            attach(root, f"{p}.{name}.bn2", bn(cout))
        attach(root, f"{p}.root.conv", conv(root_dim, cout, 1))
        attach(root, f"{p}.root.bn", bn(cout))
    else:
        _tree_params(root, f"{p}.tree1", levels - 1, cin, cout, False, 0)
        _tree_params(root, f"{p}.tree2", levels - 1, cout, cout, False, root_dim + cout)
    if cin != cout:
        attach(root, f"{p}.project.0", conv(cin, cout, 1))
        attach(root, f"{p}.project.1", bn(cout))


def _ida_params(root, p, o, channels, up_f):
    for i in range(1, len(channels)):
        attach(root, f"{p}.proj_{i}.actf.0", bn(o))
        attach(root, f"{p}.proj_{i}.conv", DCNParams(channels[i], o))
        attach(root, f"{p}.up_{i}", bilinear_up(o, int(up_f[i])))
        attach(root, f"{p}.node_{i}.actf.0", bn(o))
        attach(root, f"{p}.node_{i}.conv", DCNParams(o, o))


def _dla_up_plan():
    """Channel / scale bookkeeping of DLAUp.__init__ (pose_dla_dcn.py:381-396)."""
    channels = CH[FIRST_LEVEL:]
    in_channels = list(channels)
    scales = [2 ** i for i in range(len(channels))]
    stages = []
    for i in range(len(channels) - 1):
        j = -i - 2
        stages.append((f"ida_{i}", channels[j], list(in_channels[j:]), [s // scales[j] for s in scales[j:]]))
        scales[j + 1:] = [scales[j]] * len(scales[j + 1:])
        in_channels[j + 1:] = [channels[j]] * len(in_channels[j + 1:])
    return stages


def build_params(cfg=None) -> nn.Module:
    root = nn.Module()
    attach(root, "base.base_layer.0", conv(3, CH[0], 7, 1, 3))
    attach(root, "base.base_layer.1", bn(CH[0]))
    attach(root, "base.level0.0", conv(CH[0], CH[0], 3, 1, 1)); attach(root, "base.level0.1", bn(CH[0]))
    attach(root, "base.level1.0", conv(CH[0], CH[1], 3, 2, 1)); attach(root, "base.level1.1", bn(CH[1]))
    for lv in range(2, 6):
        _tree_params(root, f"base.level{lv}", TREE_LEVELS[lv], CH[lv - 1], CH[lv], level_root=(lv > 2))
    for name, o, chans, ups in _dla_up_plan():
        _ida_params(root, f"dla_up.{name}", o, chans, ups)
    out_c = CH[FIRST_LEVEL]
    _ida_params(root, "ida_up", out_c, CH[FIRST_LEVEL:LAST_LEVEL], [2 ** i for i in range(LAST_LEVEL - FIRST_LEVEL)])
    return root


# ---------------------------------------------------------------------------- lowering
def _conv_bn(pb: PlanBuilder, P: StateView, srcs, ckey, bkey, stride=1, pad=0, relu=True, res=None):
    w, b = fold_bn(P(ckey + ".weight"), None, P.bn(bkey))
    return pb.conv(srcs, w, b, stride=stride, pad=pad, relu=relu, res=res)


def _block(pb, P, x, p, stride, residual=None):
    """BasicBlock (pose_dla_dcn.py:43-57): two fused conv ops."""
    t = _conv_bn(pb, P, [x], p + ".conv1", p + ".bn1", stride=stride, pad=1, relu=True)
    return _conv_bn(pb, P, [t], p + ".conv2", p + ".bn2", stride=1, pad=1, relu=True,
                    res=residual if residual is not None else x)


def _tree(pb, P, x, p, levels, cin, cout, stride, level_root, children=None):
    """Tree.forward (pose_dla_dcn.py:206-219)."""
    children = [] if children is None else children
    bottom = pb.maxpool(x, stride, stride) if stride > 1 else x
    if level_root:
        children.append(bottom)
    if levels == 1:
        residual = _conv_bn(pb, P, [bottom], p + ".project.0", p + ".project.1", relu=False) \
            if cin != cout else bottom
        x1 = _block(pb, P, x, p + ".tree1", stride, residual)
        x2 = _block(pb, P, x1, p + ".tree2", 1)
        return _conv_bn(pb, P, [x2, x1] + children, p + ".root.conv", p + ".root.bn", relu=True)
    # levels == 2: the outer `project` (ops 12 / 26 in SURVEY Appendix A) is dead compute in the
    # reference — its result is overwritten inside the sub-tree — so it is not lowered.
    x1 = _tree(pb, P, x, p + ".tree1", levels - 1, cin, cout, stride, False)
    children.append(x1)
    return _tree(pb, P, x1, p + ".tree2", levels - 1, cout, cout, 1, False, children)


def _deform(pb, P, x, p):
    """DeformConv (pose_dla_dcn.py:336-348): DCN + BN + ReLU, BN folded into the DCN GEMM."""
    w, b = fold_bn(P(p + ".conv.weight"), P(p + ".conv.bias"), P.bn(p + ".actf.0"))
    return pb.dcn(x, w, b, P(p + ".conv.conv_offset_mask.weight"), P(p + ".conv.conv_offset_mask.bias"))


def _ida(pb, P, layers, p, startp, endp):
    """IDAUp.forward (pose_dla_dcn.py:371-377)."""
    for i in range(startp + 1, endp):
        k = i - startp
        y = _deform(pb, P, layers[i], f"{p}.proj_{k}")
        y = pb.up_add(y, layers[i - 1], P(f"{p}.up_{k}.weight"))
        layers[i] = _deform(pb, P, y, f"{p}.node_{k}")


def lower(pb: PlanBuilder, P: StateView, x: Sym) -> Sym:
    """DLASeg.forward (pose_dla_dcn.py:437-447) -> feature map (64 ch, stride 4)."""
    w, b = fold_bn(P("base.base_layer.0.weight"), None, P.bn("base.base_layer.1"))
    t = pb.stem(x, w, b, 7, 1, 3, relu=True)                                       # op 1
    l0 = _conv_bn(pb, P, [t], "base.level0.0", "base.level0.1", 1, 1)              # op 2
    l1 = _conv_bn(pb, P, [l0], "base.level1.0", "base.level1.1", 2, 1)             # op 3
    layers = [l0, l1]
    cur = l1
    for lv in range(2, 6):
        cur = _tree(pb, P, cur, f"base.level{lv}", TREE_LEVELS[lv], CH[lv - 1], CH[lv], 2, lv > 2)
        layers.append(cur)
    # DLAUp.forward (:398-404)
    out = [layers[-1]]
    work = list(layers)
    for i in range(len(work) - FIRST_LEVEL - 1):
        _ida(pb, P, work, f"dla_up.ida_{i}", len(work) - i - 2, len(work))
        out.insert(0, work[-1])
    y = list(out[:LAST_LEVEL - FIRST_LEVEL])
    _ida(pb, P, y, "ida_up", 0, len(y))
    return y[-1]
