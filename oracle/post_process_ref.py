"""ORACLE (test infrastructure, not product): numpy restatement of the reference's
detection back-projection (network output cells -> original image pixels).

Follows (``/root/reference/lib``):
  * ``utils/image.py:27-60``  ``get_affine_transform``  -> :func:`get_affine_transform`
    (``cv2.getAffineTransform`` = exact 3-point solve in float64 on float32 points; restated
    with ``numpy.linalg.solve`` so the oracle has no cv2 dependency)
  * ``utils/image.py:63-66``  ``affine_transform``      -> inside :func:`transform_preds`
  * ``utils/image.py:19-24``  ``transform_preds``       -> :func:`transform_preds`
  * ``utils/post_process.py:8-19`` ``multi_pose_post_process`` -> :func:`multi_pose_post_process`
  * ``detectors/multi_pose.py:62-71`` ``MultiPoseDetector.post_process`` -> :func:`detector_post_process`
  * ``detectors/base_detector.py:36-47`` c / s / out size meta      -> :func:`make_meta`
"""
from __future__ import annotations

import numpy as np


def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _solve_affine(src, dst):
    """2x3 matrix M with M @ [x,y,1] = dst for the three src points (float64 solve)."""
    A = np.zeros((6, 6), np.float64); rhs = np.zeros(6, np.float64)
    for i in range(3):
        A[2 * i, 0:3] = [src[i, 0], src[i, 1], 1.0]
        A[2 * i + 1, 3:6] = [src[i, 0], src[i, 1], 1.0]
        rhs[2 * i] = dst[i, 0]; rhs[2 * i + 1] = dst[i, 1]
    return np.linalg.solve(A, rhs).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0):
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    scale = np.asarray(scale)
    shift = np.asarray(shift, dtype=np.float32)
    src_w = scale[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale * shift
    src[1, :] = center + src_dir + scale * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    if inv:
        return _solve_affine(dst, src)
    return _solve_affine(src, dst)


def transform_preds(coords, center, scale, output_size):
    """image.py:19-24 — float64 result; each point is first cast to float32 (image.py:64)."""
    trans = get_affine_transform(center, scale, 0, output_size, inv=1)
    pts = np.concatenate([np.asarray(coords, np.float32)[:, 0:2],
                          np.ones((coords.shape[0], 1), np.float32)], axis=1)
    return (pts.astype(np.float64) @ trans.T)


def multi_pose_post_process(dets, c, s, h, w):
    """post_process.py:8-19 — dets (B,N,56) -> list of {1: (N,56) float32 array}."""
    ret = []
    for i in range(dets.shape[0]):
        bbox = transform_preds(dets[i, :, :4].reshape(-1, 2), c[i], s[i], (w, h))
        pts = transform_preds(dets[i, :, 5:39].reshape(-1, 2), c[i], s[i], (w, h))
        top = np.concatenate([bbox.reshape(-1, 4), dets[i, :, 4:5],
                              pts.reshape(-1, 34), dets[i, :, 39:56]], axis=1).astype(np.float32)
        ret.append({1: top})
    return ret


def detector_post_process(dets, meta, scale=1.0):
    """multi_pose.py:62-71 for one image: dets (1,K,56) -> {1: (K,56) float32}."""
    dets = np.asarray(dets, np.float32).reshape(1, -1, dets.shape[-1])
    out = multi_pose_post_process(dets.copy(), [meta["c"]], [meta["s"]],
                                  meta["out_height"], meta["out_width"])[0]
    out[1] = np.array(out[1], dtype=np.float32).reshape(-1, 56)
    out[1][:, :4] /= scale
    out[1][:, 5:39] /= scale
    return out


def make_meta(height, width, scale=1.0, fix_res=True, input_h=512, input_w=512, pad=31, down_ratio=4):
    """base_detector.py:33-47,59-61 — c, s and output size for an image of (height, width)."""
    new_h = int(height * scale); new_w = int(width * scale)
    if fix_res:
        inp_h, inp_w = input_h, input_w
        c = np.array([new_w / 2., new_h / 2.], dtype=np.float32)
        s = max(height, width) * 1.0
    else:
        inp_h = (new_h | pad) + 1
        inp_w = (new_w | pad) + 1
        c = np.array([new_w // 2, new_h // 2], dtype=np.float32)
        s = np.array([inp_w, inp_h], dtype=np.float32)
    return {"c": c, "s": s, "out_height": inp_h // down_ratio, "out_width": inp_w // down_ratio,
            "inp_height": inp_h, "inp_width": inp_w}
