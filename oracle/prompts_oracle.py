"""TEST INFRASTRUCTURE -- CPU restatement of the reference's prompt generation
(``utils/sam2_data_generator.py:43-180``), one scipy distance transform per instance like the
reference.  Only tests/ may import this; the product path is calcium_b200/prompts.py on K4.

Pinned by ``tests/golden/prompts_small.json``: prompts the reference's own functions produced in
the build container (``oracle/make_golden.py``) for instance masks of a small fluttering pouch.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
from scipy.ndimage import distance_transform_edt


def point_prompts(instance_mask: np.ndarray, min_cell_area: int = 10) -> Dict[int, Dict]:
    """generate_point_prompts_from_mask(instance_mask, 'distance_based')  (:43-131)."""
    prompts = {}
    ids = np.unique(instance_mask)
    for cid in ids[ids != 0]:
        b = (instance_mask == cid).astype(np.uint8)
        area = int(b.sum())
        if area < min_cell_area:                                  # :75
            continue
        d = distance_transform_edt(b)                             # :80
        if d.max() == 0:
            continue
        thr = np.percentile(d[d > 0], 80)                         # :86
        ys, xs = np.where(d >= thr)                               # :87
        idx = np.random.randint(len(ys))                          # :91
        y, x = ys[idx], xs[idx]
        prompts[int(cid)] = {"point": [int(x), int(y)], "label": 1, "area": area, "confidence": float(d[y, x])}
    return prompts


def negative_prompts(instance_mask: np.ndarray, num_negative: int = 5, min_distance_from_cells: int = 5) -> List[Dict]:
    """generate_negative_prompts  (:134-180)."""
    d = distance_transform_edt(1 - (instance_mask > 0).astype(np.uint8))       # :152-155
    ys, xs = np.where(d >= min_distance_from_cells)                             # :158-159
    out = []
    if len(ys) > 0:
        for idx in np.random.choice(len(ys), min(num_negative, len(ys)), replace=False):   # :166
            y, x = ys[idx], xs[idx]
            out.append({"point": [int(x), int(y)], "label": 0, "confidence": float(d[y, x])})
    return out
