"""COCO keypoint result formatter: the reference's ``COCOHP.convert_eval_format`` / ``save_results``
(``lib/datasets/coco_hp.py:56-90``) as free functions, so ``tools/evaluate.py`` can dump ``results.json``
without the dataset class.  Input: ``{image_id: {1: rows}}`` with rows of 56 floats
``[x1,y1,x2,y2,score, 17 x (x,y), 17 keypoint scores]`` (what ``run()['results']`` returns).

Same numbers as the reference: bbox as ``[x, y, w, h]``, every float formatted with ``'{:.2f}'`` (Python's
round-half-even on the decimal repr), keypoint visibility ``1`` where the keypoint score exceeds 0.1.
Unlike the reference the input rows are NOT modified in place (it rewrites ``dets[2:4]`` to width / height).
"""
from __future__ import annotations

import json

import numpy as np


def _f2(x) -> float:
    return float("{:.2f}".format(x))


def convert_eval_format(all_bboxes, category_id: int = 1):
    detections = []
    for image_id in all_bboxes:
        rows = np.asarray(all_bboxes[image_id][category_id], dtype=np.float64).reshape(-1, 56)
        for r in rows:
            bbox = [r[0], r[1], r[2] - r[0], r[3] - r[1]]
            kps = np.asarray(r[5:39], dtype=np.float32).reshape(17, 2)
            vis = (r[39:56] > 0.1).astype(np.int32).reshape(17, 1)
            kp51 = np.concatenate([kps, vis], axis=1).reshape(51).tolist()
            detections.append({"image_id": int(image_id), "category_id": int(category_id),
                               "bbox": [_f2(v) for v in bbox], "score": _f2(r[4]),
                               "keypoints": [_f2(v) for v in kp51]})
    return detections


def save_results(results, save_dir: str):
    with open("{}/results.json".format(save_dir), "w") as f:
        json.dump(convert_eval_format(results), f)
