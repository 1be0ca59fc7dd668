"""GPU tests of the detector mirror (run / process / post_process / merge_outputs, flip test, batch API)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _detector(precision="fp32", flip=False, nms=False):
    from centerpose_b200.config import default_cfg
    from centerpose_b200.detector import detector_factory
    from oracle.init_recipe import conditioned_state_dict
    cfg = default_cfg("dla_34")
    cfg.TEST.FLIP_TEST = flip
    cfg.TEST.NMS = nms
    det = detector_factory[cfg.TEST.TASK](cfg)
    sd = conditioned_state_dict(det.model.state_dict(), 317)
    det.model.load_state_dict(sd)
    det.model.set_precision(precision)
    return det, sd, cfg


def test_flip_helpers_match_reference_golden():
    """Device-side flip_lr / flip_lr_off == the reference's numpy round-trip versions (lib/models/utils.py:27-47)."""
    g = np.load(os.path.join(GOLD, "flip.npz"))
    det, _, _ = _detector()
    hm_hp = torch.from_numpy(g["hm_hp"]).cuda(); hps = torch.from_numpy(g["hps"]).cuda()
    assert np.array_equal(torch.flip(hm_hp, [3]).cpu().numpy(), g["flip_tensor"])
    assert np.array_equal(det._flip_lr(hm_hp).cpu().numpy(), g["flip_lr"])
    assert np.array_equal(det._flip_lr_off(hps).cpu().numpy(), g["flip_lr_off"])


def _oracle_run(det, sd, cfg, image, flip):
    from oracle import decode_ref, dla_ref, post_process_ref
    images, meta = det.pre_process(image, 1.0)
    hm, wh, hps, reg, hm_hp, hp_off = dla_ref.forward(sd, images.cpu())     # pre_process may run on the device
    hm = hm.sigmoid(); hm_hp = hm_hp.sigmoid()
    if flip:
        idx = list(range(17))
        for a, b in det.flip_idx:
            idx[a], idx[b] = idx[b], idx[a]
        fl = lambda t: torch.flip(t, [3])
        hm = (hm[0:1] + fl(hm[1:2])) / 2
        wh = (wh[0:1] + fl(wh[1:2])) / 2
        h2 = fl(hps[1:2]).view(1, 17, 2, *hps.shape[2:]).clone(); h2[:, :, 0] *= -1
        hps = (hps[0:1] + h2[:, idx].reshape(1, 34, *hps.shape[2:])) / 2
        hm_hp = (hm_hp[0:1] + fl(hm_hp[1:2])[:, idx]) / 2
        reg = reg[0:1]; hp_off = hp_off[0:1]
    dets = decode_ref.multi_pose_decode(hm.numpy(), wh.numpy(), hps.numpy(), reg.numpy(), hm_hp.numpy(), hp_off.numpy(), K=100)
    return post_process_ref.detector_post_process(dets, meta, 1.0)[1]


@pytest.mark.parametrize("flip", [False, True])
def test_detector_run_matches_oracle_pipeline(flip):
    from tests.util import match_rows
    det, sd, cfg = _detector("fp32", flip=flip)
    rng = np.random.RandomState(7)
    image = rng.randint(0, 256, size=(360, 480, 3)).astype(np.uint8)
    ret = det.run(image)
    assert set(ret) == {"results", "tot", "load", "pre", "net", "dec", "post", "merge"}
    rows = np.asarray(ret["results"][1], dtype=np.float32)
    assert rows.shape == (100, 56)
    want = _oracle_run(det, sd, cfg, image, flip)
    frac_rows, frac_elems = match_rows(rows, want, tol=2e-3, box_tol=5e-2)
    assert frac_rows >= 0.9 and frac_elems >= 0.97, (frac_rows, frac_elems)


def test_run_batch_and_fused_post_process_agree():
    from oracle import post_process_ref
    det, sd, cfg = _detector("bf16")
    x = torch.randn(3, 3, 512, 512, generator=torch.Generator().manual_seed(1))
    metas = [post_process_ref.make_meta(480, 640), post_process_ref.make_meta(720, 1280), post_process_ref.make_meta(512, 512)]
    a = det.run_batch(x, metas)
    b = det.run_batch_fused(x, metas)
    assert b.shape == (3, 100, 56)
    for i in range(3):
        assert np.abs(np.asarray(a[i][1]) - b[i]).max() <= 2e-3


def test_merge_outputs_soft_nms_path():
    det, _, _ = _detector("bf16", nms=True)
    rng = np.random.RandomState(0)
    d = {1: rng.uniform(0, 100, size=(100, 56)).astype(np.float32)}
    d[1][:, 2:4] = d[1][:, 0:2] + 30
    out = det.merge_outputs([d])
    assert len(out) == 100 and len(out[0]) == 56


@pytest.mark.parametrize("h,w,fix_res,flip", [(480, 640, True, False), (427, 640, False, True), (333, 500, True, True),
                                              (96, 1280, False, False), (720, 405, True, False)])
def test_device_pre_process_is_bit_exact_with_cv2_path(h, w, fix_res, flip):
    """cpb200_pre_process (warpAffine + normalise + HWC->CHW + mirrored copy in one kernel) vs the reference's
    cv2.warpAffine / numpy lines (base_detector.py:44-55), which the host branch of pre_process mirrors verbatim."""
    from centerpose_b200.config import default_cfg
    from centerpose_b200.detector import MultiPoseDetector
    cfg = default_cfg("dla_34")
    cfg.TEST.FIX_RES = fix_res; cfg.TEST.FLIP_TEST = flip
    det = _shared_detector(cfg)
    rng = np.random.RandomState(h * 7 + w)
    img = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
    det.device_preprocess = False
    ref, meta_ref = det.pre_process(img, 1.0)
    det.device_preprocess = True
    got, meta = det.pre_process(img, 1.0)
    assert got.is_cuda and got.shape == ref.shape and meta["out_height"] == meta_ref["out_height"]
    assert torch.equal(got.cpu(), ref)
    # a scale != 1 keeps cv2.resize on the host and still warps on the device
    det.device_preprocess = False
    ref2, _ = det.pre_process(img, 0.75)
    det.device_preprocess = True
    got2, _ = det.pre_process(img, 0.75)
    assert torch.equal(got2.cpu(), ref2)


_DET = {}


def _shared_detector(cfg):
    """One model for all parametrisations (the pre_process path does not depend on the weights)."""
    from centerpose_b200.detector import MultiPoseDetector
    if "d" not in _DET:
        _DET["d"] = MultiPoseDetector(cfg)
    _DET["d"].cfg = cfg
    return _DET["d"]


def test_device_merge_with_cuda_soft_nms_matches_host_path():
    """run_batch_fused(nms=True) / merge_outputs_device: decode + back-projection + soft_nms_39 entirely on the device
    vs the host pipeline (post_process in numpy + the pinned host port of lib/external/nms.pyx:172-275).  Same kept
    rows; values within 2e-3 px (the fused affine evaluates the 2x3 map in fp32)."""
    from oracle import post_process_ref
    det, sd, cfg = _detector("fp16x2", nms=True)
    x = torch.randn(2, 3, 512, 512, generator=torch.Generator().manual_seed(3))
    metas = [post_process_ref.make_meta(480, 640), post_process_ref.make_meta(720, 1280)]
    host = det.run_batch(x, metas)                       # per image {1: (100, 56)} in image pixels, no NMS yet
    fused = det.run_batch_fused(x, metas)                # (B, 100, 56), soft-NMS applied per image on the device
    assert fused.shape == (2, 100, 56)
    for i in range(2):
        want = np.asarray(det.merge_outputs([host[i]]), dtype=np.float32)
        got = fused[i]
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.abs(got - want).max() <= 2e-3 * max(1.0, np.abs(want).max()), float(np.abs(got - want).max())


def test_multiscale_fused_matches_run():
    """TEST_SCALES [1, 0.75] (the multi-scale merge of lib/detectors/multi_pose.py:73-79): device-resident pipeline vs
    run() — rows matched on bbox + score (top-K / NMS are discontinuous), then compared."""
    from tests.util import match_rows
    det, sd, cfg = _detector("fp16x2")
    cfg.TEST.TEST_SCALES = [1, 0.75]
    det.scales = cfg.TEST.TEST_SCALES
    rng = np.random.RandomState(11)
    image = rng.randint(0, 256, size=(384, 512, 3)).astype(np.uint8)
    want = np.asarray(det.run(image)["results"][1], dtype=np.float32)
    got = det.run_multiscale_fused(image)
    assert abs(len(got) - len(want)) <= 2
    rows, elems = match_rows(got, want, tol=2e-3, box_tol=5e-2)
    assert rows >= 0.97 and elems >= 0.99, (rows, elems)


def test_batched_paths_honour_loss_flags_and_topk_is_validated():
    """run_batch* apply the same cfg gating as process() (MSE_LOSS: hm_hp used raw; REG_* off: +0.5), and a TOPK the
    fused decode cannot serve is rejected at construction."""
    from centerpose_b200 import multi_pose_decode
    from centerpose_b200.config import default_cfg
    from centerpose_b200.detector import detector_factory
    det, sd, cfg = _detector("fp16x2")
    x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(5)).cuda()
    cfg.LOSS.MSE_LOSS = True; cfg.LOSS.REG_OFFSET = False
    got = det.run_batch(x)
    hm, wh, hps, reg, hm_hp, hp_off = det.model(x)
    from centerpose_b200.decode import sigmoid_
    want = multi_pose_decode(sigmoid_(hm.clone()), wh, hps, reg=None, hm_hp=hm_hp, hp_offset=hp_off, K=100)
    assert torch.equal(got, want)
    _, dets = det.process(x)
    assert torch.equal(dets, want)
    bad = default_cfg("dla_34"); bad.TEST.TOPK = 200
    with pytest.raises(ValueError):
        detector_factory[bad.TEST.TASK](bad)
