"""Import shims that let the UNMODIFIED reference run on CPU in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``centerpose_b200/`` may import this
file.  It is used by ``oracle/make_golden.py`` (fixture generation + pinning of
the oracle restatement) and only works where ``/root/reference`` is mounted —
it does not exist on the GPU box, so no ``-m gpu`` test, ``smoke()`` or
``bench.py`` path may call it at run time.

What it does:
  1. puts ``/root/reference/lib`` on ``sys.path`` (what ``tools/_init_paths.py:10-13`` does);
  2. pre-registers ``sys.modules['_ext']`` because ``lib/models/model.py:9-22``
     imports every backbone and ``DCNv2/dcn_v2.py:12`` imports ``_ext`` at module
     scope; the stub maps ``dcn_v2_forward`` to ``torchvision.ops.deform_conv2d``
     (the reference's own CPU path is ``AT_ERROR("Not implement on cpu")``,
     ``DCNv2/src/cpu/dcn_v2_cpu.cpp:23``; its CUDA sources need the removed THC API);
  3. neutralises the two network downloads (``pose_dla_dcn.py:292-303``,
     ``msra_resnet.py:226-230``);
  4. builds ``cfg`` from ``yaml.safe_load`` + an attribute dict (``yacs`` is absent).
"""
from __future__ import annotations

import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("CENTERPOSE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


class AttrDict(dict):
    """dict with attribute access, recursively (supports cfg.MODEL.EXTRA and cfg['MODEL']['EXTRA'])."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _install_ext_stub():
    if "_ext" in sys.modules:
        return
    from torchvision.ops import deform_conv2d

    ext = types.ModuleType("_ext")

    def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        return deform_conv2d(input, offset, weight, bias, stride=(sh, sw), padding=(ph, pw),
                             dilation=(dh, dw), mask=mask)

    def _no(*a, **k):  # pragma: no cover
        raise RuntimeError("reference DCNv2 backward/psroi is out of scope")

    ext.dcn_v2_forward = dcn_v2_forward
    ext.dcn_v2_backward = _no
    ext.dcn_v2_psroi_pooling_forward = _no
    ext.dcn_v2_psroi_pooling_backward = _no
    sys.modules["_ext"] = ext


_loaded = {}


def _load_patched_image_module(lib):
    """``lib/utils/image.py:139-140`` as shipped is a SyntaxError (an ``if`` whose body
    is not indented, inside the training-only ``draw_msra_gaussian``).  The hot-path
    functions (``get_affine_transform``/``affine_transform``/``transform_preds``,
    ``image.py:19-66``) are untouched; we indent that one body line IN MEMORY (the
    reference tree is read-only and stays unmodified) so the module can be imported."""
    import importlib.util
    path = os.path.join(lib, "utils", "image.py")
    src = open(path).read()
    bad = "\n    np.maximum(masked_heatmap, masked_gaussian * k, out=masked_heatmap)\n    return heatmap\n\ndef draw_dense_reg"
    good = "\n        np.maximum(masked_heatmap, masked_gaussian * k, out=masked_heatmap)\n    return heatmap\n\ndef draw_dense_reg"
    if bad in src:
        src = src.replace(bad, good)
    import utils  # the reference's package (lib/utils/__init__.py)
    spec = importlib.util.spec_from_loader("utils.image", loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    sys.modules["utils.image"] = mod
    utils.image = mod
    return mod


def load_reference():
    """Returns a namespace with the reference's hot-path callables."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    lib = os.path.join(REF_ROOT, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    _install_ext_stub()
    import models.backbones.pose_dla_dcn as ref_dla
    import models.backbones.msra_resnet as ref_res
    ref_dla.DLA.load_pretrained_model = lambda self, *a, **k: None
    ref_res.model_zoo.load_url = lambda *a, **k: {}
    import models.model as ref_model
    import models.decode as ref_decode
    import models.utils as ref_utils
    ref_image = _load_patched_image_module(lib)
    import utils.post_process as ref_post

    ns = types.SimpleNamespace(
        create_model=ref_model.create_model, load_model=ref_model.load_model,
        save_model=ref_model.save_model,
        multi_pose_decode=ref_decode.multi_pose_decode, _nms=ref_decode._nms,
        _topk=ref_decode._topk, _topk_channel=ref_decode._topk_channel,
        flip_tensor=ref_utils.flip_tensor, flip_lr=ref_utils.flip_lr,
        flip_lr_off=ref_utils.flip_lr_off,
        multi_pose_post_process=ref_post.multi_pose_post_process,
        transform_preds=ref_image.transform_preds,
        get_affine_transform=ref_image.get_affine_transform,
        dla=ref_dla, res=ref_res,
    )
    _loaded["ns"] = ns
    return ns


def load_ref_cfg(name: str) -> AttrDict:
    """name like 'dla_34_512x512' -> cfg from the reference's experiments/*.yaml."""
    with open(os.path.join(REF_ROOT, "experiments", name + ".yaml")) as f:
        return AttrDict(yaml.safe_load(f))


def ref_create_model(yaml_name: str):
    ns = load_reference()
    cfg = load_ref_cfg(yaml_name)
    torch.manual_seed(int(cfg.SEED))
    model = ns.create_model(cfg.MODEL.NAME, cfg.MODEL.HEAD_CONV, cfg)
    return model.eval(), cfg
