"""SAM2 point prompts -- drop-ins for the reference's
``generate_point_prompts_from_mask`` / ``generate_negative_prompts`` / ``save_prompts``
(``utils/sam2_data_generator.py:43-180, 210-240``) on top of K4 (``cab200_edt_classes``).

The reference runs one full-image ``distance_transform_edt`` per instance and frame (4-5 ms each:
0.3-0.6 s per frame of a fluttering pouch).  Here:

* An instance that is a single cell -- every instance except the one the reference's id mismatch
  creates out of the background -- is bordered only by pixels of *other* ids, so its
  distance transform, its 80th-percentile threshold, its candidate pixels and their distances do not
  depend on the frame at all.  :class:`PromptTables` computes them once per (geometry, output size):
  one K4 pass over the cells mask (class = cell label), then numpy's own ``percentile`` per cell.
  Per frame a positive prompt is then one draw ``np.random.randint(n_candidates)`` and a table
  look-up.
* The frame-dependent transforms -- the "background" instance of the id mismatch, and the distance
  of background pixels to the nearest labelled cell for the negative prompts -- are two more K4
  passes (classes ``id == 1`` and ``id > 0``); their selection runs in numpy on the two int32 maps.

Random draws use the *global* legacy numpy generator with the reference's calls in the reference's
order (one ``randint`` per instance in ascending id order, then one ``choice(n, k, replace=False)``),
so with the same generator state the prompts are identical to the reference's, value for value
(tests/test_prompts.py).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def edt_squared(label_map: torch.Tensor, class_luts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K4: int32 ``[n_maps, H, W]`` squared distance to the nearest pixel of a different class.
    ``label_map``: uint16-as-int16 ``[H, W]`` device tensor; ``class_luts``: int16 ``[n_maps, L]``
    (class per label value) or ``None`` (class = label)."""
    lib = _lib.load()
    dev = label_map.device
    h, w = int(label_map.shape[0]), int(label_map.shape[1])
    n_maps = 1 if class_luts is None else int(class_luts.shape[0])
    with torch.cuda.device(dev):
        out = torch.empty((n_maps, h, w), dtype=torch.int32, device=dev)
        nb = int(lib.cab200_edt_scratch_bytes(n_maps, h, w))
        scratch = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        rc = lib.cab200_edt_classes(label_map.data_ptr(), class_luts.data_ptr() if class_luts is not None else None,
                                    int(class_luts.shape[1]) if class_luts is not None else 0, n_maps, h, w,
                                    out.data_ptr(), scratch.data_ptr(), nb, _stream_ptr(dev))
        _lib.check(rc, "cab200_edt_classes")
    return out


def instance_lut(ca: np.ndarray, threshold: float = 0.1, quirk: bool = True) -> np.ndarray:
    """Instance id per label value (0 = background, v = cell + 1) for one frame: the composition of
    ``get_active_cells`` (``core/pouch.py:563-565``), ``get_cell_masks(active_only=True)`` (``:540-546``)
    and ``create_instance_mask_from_cells`` (``utils/sam2_data_generator.py:31-35``) -- the same table K2 /
    K3 build on the device (``csrc/frame_lut.cuh``).  ``quirk`` keeps the reference's comparison of
    0-based active ids with the 1-based mask."""
    n = len(ca)
    act = np.asarray(ca) > threshold
    rank = np.cumsum(act) * act                       # 1-based rank among active cells, by cell
    lut = np.zeros(n + 1, dtype=np.uint16)
    if quirk:
        m = np.arange(n + 1)
        m[1:] = np.where(act, m[1:], 0)               # masked label value: v if cell v-1 active else 0
        nxt = np.zeros(n + 1, dtype=bool)
        nxt[:n] = act                                 # nxt[m] = cell m active (m < n)
        r1 = np.zeros(n + 1, dtype=np.int64)
        r1[:n] = rank
        lut[:] = np.where(nxt[m], r1[m], 0)
    else:
        lut[1:] = rank
    return lut


class PromptTables:
    """Static per-cell prompt data of one (geometry, output size): area, candidate pixels (distance
    >= the 80th percentile of the cell's distance transform, row-major) and their distances."""

    def __init__(self, label_mask: torch.Tensor, n_cells: int):
        self.label_mask = label_mask                                  # device, int16 view of uint16 [H, W]
        self.n = int(n_cells)
        lab = label_mask.cpu().numpy().view(np.uint16).astype(np.int64)
        self.labels_host = lab
        self.H, self.W = lab.shape
        d2 = edt_squared(label_mask)[0].cpu().numpy()
        flat = lab.ravel()
        order = np.argsort(flat, kind="stable")                       # groups pixels by label, row-major inside
        bounds = np.searchsorted(flat[order], np.arange(self.n + 2))
        self.area = np.diff(bounds)[1:self.n + 1].astype(np.int64)   # area[c] of cell c
        self.cand_off = np.zeros(self.n + 1, dtype=np.int64)
        cy, cx, cd = [], [], []
        for c in range(self.n):
            px = order[bounds[c + 1]:bounds[c + 2]]
            if len(px):
                dist = np.sqrt(d2.ravel()[px].astype(np.float64))      # scipy: sqrt of the exact integer
                thr = np.percentile(dist[dist > 0], 80)                 # utils/sam2_data_generator.py:79
                keep = px[dist >= thr]                                  # :80, np.where order = row-major
                cy.append(keep // self.W); cx.append(keep % self.W); cd.append(np.sqrt(d2.ravel()[keep].astype(np.float64)))
                self.cand_off[c + 1] = self.cand_off[c] + len(keep)
            else:
                self.cand_off[c + 1] = self.cand_off[c]
        self.cand_y = np.concatenate(cy) if cy else np.zeros(0, np.int64)
        self.cand_x = np.concatenate(cx) if cx else np.zeros(0, np.int64)
        self.cand_d = np.concatenate(cd) if cd else np.zeros(0, np.float64)


def _frame_transforms(tables: PromptTables, lut: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The two frame-dependent transforms on the device: squared distance to the nearest pixel whose
    (id == 1) / (id > 0) class differs."""
    luts = np.stack([(lut == 1), (lut > 0)]).astype(np.uint16).view(np.int16)
    d = edt_squared(tables.label_mask, torch.from_numpy(luts).to(tables.label_mask.device))
    both = d.cpu().numpy()
    return both[0], both[1]


def frame_prompts(tables: PromptTables, ca: np.ndarray, threshold: float = 0.1, quirk: bool = True,
                  num_negative: int = 5, min_distance_from_cells: int = 5, min_cell_area: int = 10
                  ) -> Tuple[Dict[int, Dict], List[Dict]]:
    """``(generate_point_prompts_from_mask(instance_mask), generate_negative_prompts(instance_mask, 5))``
    for the frame whose calcium values are ``ca`` -- the dicts of ``utils/sam2_data_generator.py:124-129``
    and ``:170-174`` -- consuming the global numpy generator exactly like those two calls."""
    n = tables.n
    lut = instance_lut(ca, threshold, quirk)
    ids = np.unique(lut)
    ids = ids[ids != 0]
    prompts: Dict[int, Dict] = {}
    # which label values carry each id: all ids but the mismatch's background id belong to one cell
    cell_of = {}
    for v in np.nonzero(lut[1:])[0]:
        cell_of.setdefault(int(lut[v + 1]), []).append(int(v))
    background_id = int(lut[0]) if lut[0] else None                    # id given to every zero pixel (cell 0 active)
    need_maps = background_id is not None or num_negative > 0
    d2_bg = d2_neg = None
    inst = None
    if need_maps:
        d2_bg, d2_neg = _frame_transforms(tables, lut)
        inst = lut[tables.labels_host]
    for iid in ids.tolist():
        if iid == background_id:
            b = inst == iid
            area = int(b.sum())
            if area < min_cell_area:
                continue
            dist = np.where(b, np.sqrt(d2_bg.astype(np.float64)), 0.0)
            if dist.max() == 0:
                continue
            thr = np.percentile(dist[dist > 0], 80)
            ys, xs = np.where(dist >= thr)
            idx = np.random.randint(len(ys))
            y, x, conf = int(ys[idx]), int(xs[idx]), float(dist[ys[idx], xs[idx]])
        else:
            (c,) = cell_of[iid]
            area = int(tables.area[c])
            if area < min_cell_area:
                continue
            o, cnt = int(tables.cand_off[c]), int(tables.cand_off[c + 1] - tables.cand_off[c])
            idx = int(np.random.randint(cnt))
            y, x, conf = int(tables.cand_y[o + idx]), int(tables.cand_x[o + idx]), float(tables.cand_d[o + idx])
        prompts[int(iid)] = {"point": [x, y], "label": 1, "area": area, "confidence": conf}

    negatives: List[Dict] = []
    if num_negative > 0:
        cells = inst > 0
        if cells.any():
            dist = np.where(cells, 0.0, np.sqrt(d2_neg.astype(np.float64)))
        else:
            # no labelled pixel at all: scipy's distance_transform_edt of an array without a zero element
            # returns the distance to the point (-1, 0) (observed with scipy 1.7 - 1.18); the reference
            # samples its negatives from that map (utils/sam2_data_generator.py:152-159)
            yy, xx = np.mgrid[0:tables.H, 0:tables.W]
            dist = np.sqrt(((yy + 1) ** 2 + xx ** 2).astype(np.float64))
        ys, xs = np.where(dist >= min_distance_from_cells)
        if len(ys) > 0:
            k = min(num_negative, len(ys))
            for idx in np.random.choice(len(ys), k, replace=False):
                y, x = ys[idx], xs[idx]
                negatives.append({"point": [int(x), int(y)], "label": 0, "confidence": float(dist[y, x])})
    return prompts, negatives


def save_prompts(prompts: Dict, output_path: str, negative_prompts: Optional[List] = None) -> str:
    """Same file as the reference's ``save_prompts`` (``utils/sam2_data_generator.py:210-240``)."""
    os.makedirs(os.path.dirname(output_path), exist_ok=True)
    data = {"positive_prompts": prompts, "negative_prompts": negative_prompts or [],
            "num_cells": len(prompts), "version": "1.0"}
    with open(output_path, "w") as f:
        json.dump(data, f, indent=2)
    return output_path
