"""Host-side geometry of the detector: the crop/warp affine between the original image and the
network input/output grid, and its vectorised application to decoded detections.

Mirrors the behaviour of ``lib/utils/image.py:19-66`` (``get_affine_transform`` /
``affine_transform`` / ``transform_preds``) and ``lib/utils/post_process.py:8-19``
(``multi_pose_post_process``); the per-point Python loop of the reference
(``image.py:22-23``, ~1 900 iterations per image) is one matrix product here.
"""
from __future__ import annotations

import cv2
import numpy as np


def _rot90_about(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def get_affine_transform(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0):
    """2x3 float64 matrix mapping source-image pixels to the ``output_size`` grid (or back when
    ``inv``): the unique affine taking (centre, centre + 'up' * scale/2, their 90-degree partner)
    to the corresponding points of the output rectangle — same construction as the reference so
    that results agree to cv2's float64 solve."""
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    scale = np.asarray(scale)
    shift = np.asarray(shift, dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
    ang = np.pi * rot / 180.0
    sn, cs = np.sin(ang), np.cos(ang)
    up = np.array([0.0, src_w * -0.5])
    src_dir = [up[0] * cs - up[1] * sn, up[0] * sn + up[1] * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32); dst = np.zeros((3, 2), np.float32)
    src[0] = center + scale * shift
    src[1] = center + src_dir + scale * shift
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    src[2] = _rot90_about(src[0], src[1]); dst[2] = _rot90_about(dst[0], dst[1])
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def transform_preds(coords, center, scale, output_size):
    """(N,2) output-grid points -> (N,2) float64 original-image pixels."""
    t = get_affine_transform(center, scale, 0, output_size, inv=1)
    pts = np.asarray(coords, np.float32)
    return pts[:, 0:2].astype(np.float64) @ t[:, :2].T + t[:, 2]


def multi_pose_post_process(dets, c, s, h, w):
    """dets (B,N,56) in output-grid units -> list (per image) of {1: list of 56-float rows} in image
    pixels, like the reference (``post_process.py:8-19``)."""
    ret = []
    for i in range(dets.shape[0]):
        d = dets[i]
        box = transform_preds(d[:, :4].reshape(-1, 2), c[i], s[i], (w, h)).reshape(-1, 4)
        pts = transform_preds(d[:, 5:39].reshape(-1, 2), c[i], s[i], (w, h)).reshape(-1, 34)
        rows = np.concatenate([box, d[:, 4:5], pts, d[:, 39:56]], axis=1).astype(np.float32)
        ret.append({1: rows.tolist()})
    return ret
