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

import ctypes
import json
import os
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def edt_squared(label_map: torch.Tensor, class_luts: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K4: int32 ``[n_maps, H, W]`` squared distance to the nearest pixel of a different class.
    ``label_map``: uint16-as-int16 ``[H, W]`` device tensor; ``class_luts``: int16 ``[n_maps, L]``
    (class per label value) or ``None`` (class = label).  ``out`` / ``scratch``: optional preallocated buffers."""
    lib = _lib.load()
    dev = label_map.device
    h, w = int(label_map.shape[0]), int(label_map.shape[1])
    n_maps = 1 if class_luts is None else int(class_luts.shape[0])
    with torch.cuda.device(dev):
        nb = int(lib.cab200_edt_scratch_bytes(n_maps, h, w))
        if out is None:
            out = torch.empty((n_maps, h, w), dtype=torch.int32, device=dev)
        if scratch is None or scratch.numel() < nb:
            scratch = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        rc = lib.cab200_edt_classes(label_map.data_ptr(), class_luts.data_ptr() if class_luts is not None else None,
                                    int(class_luts.shape[1]) if class_luts is not None else 0, n_maps, h, w,
                                    out.data_ptr(), scratch.data_ptr(), nb, _stream_ptr(dev))
        _lib.check(rc, "cab200_edt_classes")
    return out


class _Workspace:
    """Per-thread buffers of the frame-dependent transforms (device maps + scratch + class tables, pinned host
    copy): allocating them per frame cost more than the kernel."""

    def __init__(self, tables: "PromptTables"):
        dev = tables.label_mask.device
        lib = _lib.load()
        h, w = tables.H, tables.W
        with torch.cuda.device(dev):
            self.out = torch.empty((2, h, w), dtype=torch.int32, device=dev)
            self.scratch = torch.empty(max(int(lib.cab200_edt_scratch_bytes(2, h, w)), 16), dtype=torch.uint8, device=dev)
            self.luts = torch.empty((2, tables.n + 1), dtype=torch.int16, device=dev)
            self.stream = torch.cuda.Stream(device=dev)      # pouch_prompts: one stream per worker thread
        self.host = torch.empty((2, h, w), dtype=torch.int32, pin_memory=True)
        self.host_luts = torch.empty((2, tables.n + 1), dtype=torch.int16, pin_memory=True)


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
        self.labels_u16 = label_mask.cpu().numpy().view(np.uint16).copy()
        self.present_labels = np.unique(self.labels_u16).astype(np.int64)      # label values that own a pixel
        self.H, self.W = lab.shape
        self._empty: Dict[float, Tuple[np.ndarray, np.ndarray]] = {}
        self._pool: List[_Workspace] = []                              # free K4 workspaces (pinned memory is
        self._pool_lock = threading.Lock()                             # expensive to allocate: 17 ms each)
        self._native: Dict[int, Any] = {}                              # cab200_prompt_tables per min distance
        self._native_lock = threading.Lock()
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
        self.empty_frame_candidates(5.0)
        self.cand_y = np.concatenate(cy) if cy else np.zeros(0, np.int64)
        self.cand_x = np.concatenate(cx) if cx else np.zeros(0, np.int64)
        self.cand_d = np.concatenate(cd) if cd else np.zeros(0, np.float64)


def _empty_frame_candidates(self, min_distance: float) -> Tuple[np.ndarray, np.ndarray]:
    """Negative-prompt candidates of a frame without any labelled pixel: row-major flat indices of the pixels
    with sqrt((y + 1)^2 + x^2) >= min_distance, and the squared distances of all pixels (flat)."""
    got = self._empty.get(min_distance)
    if got is None:
        yy, xx = np.mgrid[0:self.H, 0:self.W]
        d2 = ((yy + 1) ** 2 + xx ** 2).astype(np.int64)
        dist = np.sqrt(d2.astype(np.float64))
        got = (np.flatnonzero(dist >= min_distance), d2.ravel())
        self._empty[min_distance] = got
    return got


PromptTables.empty_frame_candidates = _empty_frame_candidates


def _candidate_entries(self) -> List[str]:
    """The prompt-file entry of a single-cell instance for every candidate pixel (static: the pixel fixes point,
    area and confidence), built on first use."""
    got = getattr(self, "_cand_text", None)
    if got is None:
        area_of = np.repeat(self.area, np.diff(self.cand_off))
        got = [_positive_entry({"point": [x, y], "label": 1, "area": a, "confidence": d})
               for x, y, a, d in zip(self.cand_x.tolist(), self.cand_y.tolist(), area_of.tolist(), self.cand_d.tolist())]
        self._cand_text = got
    return got


PromptTables.candidate_entries = _candidate_entries


def _acquire_workspace(self) -> "_Workspace":
    with self._pool_lock:
        if self._pool:
            return self._pool.pop()
    return _Workspace(self)


def _release_workspace(self, ws: "_Workspace") -> None:
    with self._pool_lock:
        self._pool.append(ws)


PromptTables.acquire_workspace = _acquire_workspace
PromptTables.release_workspace = _release_workspace

_TABLES: Dict[Any, PromptTables] = {}
_TABLES_LOCK = threading.Lock()


def prompt_tables(label_mask: torch.Tensor, n_cells: int) -> PromptTables:
    """Process-wide cache of :class:`PromptTables` per label map (the label maps themselves are cached per
    geometry, output size and device, so their address identifies them)."""
    key = (label_mask.data_ptr(), str(label_mask.device), tuple(label_mask.shape), int(n_cells))
    with _TABLES_LOCK:
        t = _TABLES.get(key)
        if t is None:
            if len(_TABLES) > 16:
                _TABLES.clear()
            t = _TABLES[key] = PromptTables(label_mask, n_cells)
        return t


def _frame_transforms(tables: PromptTables, lut: np.ndarray, background: bool = True, held: Optional[list] = None
                      ) -> Tuple[Optional[np.ndarray], np.ndarray]:
    """The frame-dependent transforms on the device (one K4 call): squared distance to the nearest pixel
    whose (id > 0) class differs and -- only when the id mismatch made a background instance -- the same for
    the class (id == 1).  Returns ``(d2_background_instance or None, d2_negative)`` as views of a pooled
    workspace's pinned buffer; the workspace is appended to ``held`` and goes back to the pool when the caller is
    done with the views (without ``held`` it is simply not reused)."""
    ws = tables.acquire_workspace()
    if held is not None:
        held.append(ws)
    k = 2 if background else 1
    hl = ws.host_luts.numpy()
    if background:
        hl[0] = lut == 1
        hl[1] = lut > 0
    else:
        hl[0] = lut > 0
    dev = tables.label_mask.device
    with torch.cuda.device(dev):
        ws.luts[:k].copy_(ws.host_luts[:k], non_blocking=True)
        d = edt_squared(tables.label_mask, ws.luts[:k], out=ws.out[:k], scratch=ws.scratch)
        ws.host[:k].copy_(d, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    both = ws.host.numpy()
    return (both[0], both[1]) if background else (None, both[0])


def legacy_choice(rng, n: int, k: int) -> np.ndarray:
    """``rng.choice(n, k, replace=False)`` of a legacy ``np.random.RandomState`` -- same indices, same generator
    state afterwards -- through ``cab200_legacy_choice`` (numpy's own algorithm as one loop outside the
    interpreter lock; numpy itself takes 2-5 ms for the ~10^5 draws of one frame and serialises the threads)."""
    if n < 4096:
        return rng.choice(n, k, replace=False)
    name, key, pos, has_gauss, cached = rng.get_state()
    if name != "MT19937":
        return rng.choice(n, k, replace=False)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    cpos = ctypes.c_int32(int(pos))
    out = np.empty(k, dtype=np.int64)
    _lib.check(_lib.load().cab200_legacy_choice(key.ctypes.data, ctypes.byref(cpos), int(n), int(k), out.ctypes.data),
               "cab200_legacy_choice")
    rng.set_state((name, key, int(cpos.value), has_gauss, cached))
    return out


def sample_background(rng, labels_u16: np.ndarray, labelled_lut: np.ndarray, d2: np.ndarray, min_d2: int, k: int
                      ) -> Tuple[np.ndarray, int]:
    """Candidate selection + draw of ``generate_negative_prompts`` (``utils/sam2_data_generator.py:152-166``) in
    one call outside the interpreter lock (``cab200_sample_background``): flat pixel indices of the drawn
    candidates (draw order) and the candidate count; ``rng`` ends in the state ``rng.choice`` would leave."""
    name, key, pos, has_gauss, cached = rng.get_state()
    if name != "MT19937":
        raise ValueError("legacy MT19937 generator expected")
    key = np.ascontiguousarray(key, dtype=np.uint32)
    cpos = ctypes.c_int32(int(pos))
    out = np.empty(max(k, 1), dtype=np.int64)
    count = ctypes.c_int64(0)
    lut8 = np.ascontiguousarray(labelled_lut, dtype=np.uint8)
    d2c = np.ascontiguousarray(d2, dtype=np.int32)
    lab = np.ascontiguousarray(labels_u16, dtype=np.uint16)
    _lib.check(_lib.load().cab200_sample_background(lab.ctypes.data, lut8.ctypes.data, int(lut8.size), d2c.ctypes.data,
                                                    int(lab.size), int(min_d2), key.ctypes.data, ctypes.byref(cpos),
                                                    int(k), out.ctypes.data, ctypes.byref(count)),
               "cab200_sample_background")
    n = int(count.value)
    if n > 0 and k > 0:
        rng.set_state((name, key, int(cpos.value), has_gauss, cached))
    return out[:min(k, n)], n


def global_rng():
    """The legacy global generator behind ``np.random.randint`` / ``np.random.choice`` (what the reference
    draws from); pass a private ``np.random.RandomState`` to ``frame_prompts`` to leave it alone."""
    return np.random.mtrand._rand


def frame_prompts(tables: PromptTables, ca: np.ndarray, threshold: float = 0.1, quirk: bool = True,
                  num_negative: int = 5, min_distance_from_cells: int = 5, min_cell_area: int = 10,
                  rng=None, as_text: bool = False):
    """``(generate_point_prompts_from_mask(instance_mask), generate_negative_prompts(instance_mask, 5))``
    for the frame whose calcium values are ``ca`` -- the dicts of ``utils/sam2_data_generator.py:124-129``
    and ``:170-174`` -- consuming the generator ``rng`` (default: the global legacy numpy generator) exactly
    like those two calls: one ``randint`` per instance in ascending id order (drawn as one array call per
    run of single-cell instances, which yields the same values and leaves the same state as the scalar
    calls), then one ``choice(n, k, replace=False)``.

    ``as_text=True`` returns the prompt file's text (what ``save_prompts`` writes, byte for byte) instead of the
    two structures: the entry of a single-cell instance depends on the drawn candidate pixel only, so it is a
    precomputed string and a frame with 1 400 instances costs 1 400 list look-ups instead of 1 400 dicts plus
    their JSON encoding."""
    rng = global_rng() if rng is None else rng
    lut = instance_lut(ca, threshold, quirk)
    background_id = int(lut[0]) if lut[0] else None                    # id given to every zero pixel (cell 0 active)
    pos_lut = lut > 0
    any_labelled = bool(pos_lut[tables.present_labels].any())          # (instance_mask > 0).any()
    d2_bg = d2_neg = None
    held: list = []                                                     # workspaces whose pinned views are in use
    try:
        if background_id is not None or (num_negative > 0 and any_labelled):
            d2_bg, d2_neg = _frame_transforms(tables, lut, background_id is not None, held)
        return _frame_prompts_body(tables, lut, background_id, pos_lut, any_labelled, d2_bg, d2_neg, num_negative,
                                   min_distance_from_cells, min_cell_area, rng, as_text)
    finally:
        for ws in held:
            tables.release_workspace(ws)


def _frame_prompts_body(tables, lut, background_id, pos_lut, any_labelled, d2_bg, d2_neg, num_negative,
                        min_distance_from_cells, min_cell_area, rng, as_text):

    # ---- positive prompts: single-cell instances come from the static tables ---------------------------
    cells = np.nonzero(lut[1:])[0]                                     # cells whose label value carries an id
    cid = lut[cells + 1].astype(np.int64)
    order = np.argsort(cid, kind="stable")
    cells, cid = cells[order], cid[order]
    if background_id is not None:
        keep = cid != background_id                                     # (cannot happen: that id is the background's)
        cells, cid = cells[keep], cid[keep]
    big = tables.area[cells] >= min_cell_area
    cells, cid = cells[big], cid[big]
    off = tables.cand_off[cells]
    cnt = tables.cand_off[cells + 1] - off
    prompts: Dict[int, Dict] = {}
    entries: List[str] = []                                            # as_text: the positive entries, in id order
    cand_text = tables.candidate_entries() if as_text else None

    def draw_cells(lo: int, hi: int) -> None:
        if hi <= lo:
            return
        pick = off[lo:hi] + rng.randint(0, cnt[lo:hi])
        if as_text:
            entries.extend(f'    "{iid}": {cand_text[k]}' for iid, k in zip(cid[lo:hi].tolist(), pick.tolist()))
            return
        for iid, x, y, a, c in zip(cid[lo:hi].tolist(), tables.cand_x[pick].tolist(), tables.cand_y[pick].tolist(),
                                   tables.area[cells[lo:hi]].tolist(), tables.cand_d[pick].tolist()):
            prompts[iid] = {"point": [x, y], "label": 1, "area": a, "confidence": c}

    if background_id is None:
        draw_cells(0, len(cid))
    else:
        split = int(np.searchsorted(cid, background_id))
        draw_cells(0, split)
        # the reference's dist = distance_transform_edt(instance == id) is 0 outside the instance, so its
        # percentile / candidate selection only ever sees the instance's own pixels: work on those (row-major,
        # the order np.where gives), same values, a fraction of the traffic
        flat_bg = np.flatnonzero((lut == background_id)[tables.labels_u16])
        area = int(flat_bg.size)
        if area >= min_cell_area:
            dist = np.sqrt(d2_bg.ravel()[flat_bg].astype(np.float64))
            positive = dist[dist > 0]
            if positive.size:                                            # dist_transform.max() != 0
                thr = np.percentile(positive, 80)
                cand = np.flatnonzero(dist >= thr)                      # thr > 0: pixels outside never qualify
                idx = rng.randint(len(cand))
                px = int(flat_bg[cand[idx]])
                bg = {"point": [px % tables.W, px // tables.W], "label": 1, "area": area,
                      "confidence": float(dist[cand[idx]])}
                if as_text:
                    entries.append(f'    "{background_id}": {_positive_entry(bg)}')
                else:
                    prompts[background_id] = bg
        draw_cells(split, len(cid))

    # ---- negative prompts -----------------------------------------------------------------------------
    negatives: List[Dict] = []
    if num_negative > 0:
        md = float(min_distance_from_cells)
        if not any_labelled:
            # no labelled pixel at all: scipy's distance_transform_edt of an array without a zero element
            # returns the distance to the point (-1, 0) (observed with scipy 1.7 - 1.18); the reference
            # samples its negatives from that map (utils/sam2_data_generator.py:152-159).  Static per size.
            flat, d2 = tables.empty_frame_candidates(md)
        elif md >= 1.0 and md.is_integer():
            # sqrt(d2) >= md  <=>  d2 >= md^2 for integers; labelled pixels have distance 0 < md.  Selection and
            # draw in one pass outside the interpreter lock
            picked, _ = sample_background(rng, tables.labels_u16, pos_lut, d2_neg, int(md) ** 2, num_negative)
            d2 = d2_neg.ravel()
            for px in picked.tolist():
                negatives.append({"point": [px % tables.W, px // tables.W], "label": 0,
                                  "confidence": float(np.sqrt(np.float64(d2[px])))})
            flat = ()
        else:
            labelled = pos_lut[tables.labels_u16]
            dist = np.where(labelled, 0.0, np.sqrt(d2_neg.astype(np.float64)))
            flat = np.flatnonzero(dist >= md)
            d2 = np.where(labelled, 0, d2_neg).ravel()
        if len(flat) > 0:
            k = min(num_negative, len(flat))
            for idx in legacy_choice(rng, len(flat), k):
                px = int(flat[idx])
                negatives.append({"point": [px % tables.W, px // tables.W], "label": 0,
                                  "confidence": float(np.sqrt(np.float64(d2[px])))})
    if as_text:
        return _assemble(entries, negatives), len(entries)
    return prompts, negatives


class _PromptTablesC(ctypes.Structure):
    """``cab200_prompt_tables`` of include/cab200.h."""
    _fields_ = [("n_cells", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("labels", ctypes.c_void_p), ("label_present", ctypes.c_void_p), ("area", ctypes.c_void_p),
                ("cand_off", ctypes.c_void_p), ("cand_text", ctypes.c_void_p), ("cand_text_off", ctypes.c_void_p),
                ("empty_flat", ctypes.c_void_p), ("empty_count", ctypes.c_int64), ("empty_d2", ctypes.c_void_p),
                ("label_map_dev", ctypes.c_void_p)]


def _native_tables(self, min_distance: int = 5) -> _PromptTablesC:
    """The static tables in the form ``cab200_prompts_pouch`` reads (built once; the arrays stay referenced here)."""
    key = int(min_distance)
    with self._native_lock:
        return self._native_tables_locked(key, min_distance)


def _native_tables_locked(self, key: int, min_distance: int) -> _PromptTablesC:
    got = self._native.get(key)
    if got is not None:
        return got[0]
    texts = [t.encode("ascii") for t in self.candidate_entries()]
    blob = np.frombuffer(b"".join(texts) or b"\0", dtype=np.uint8).copy()
    toff = np.concatenate([[0], np.cumsum([len(t) for t in texts])]).astype(np.int64)
    present = np.zeros(self.n + 1, dtype=np.uint8)
    present[self.present_labels[self.present_labels <= self.n]] = 1
    flat, d2 = self.empty_frame_candidates(float(min_distance))
    keep = {"blob": blob, "toff": toff, "present": present, "labels": np.ascontiguousarray(self.labels_u16),
            "area": np.ascontiguousarray(self.area, dtype=np.int64), "cand_off": np.ascontiguousarray(self.cand_off, dtype=np.int64),
            "flat": np.ascontiguousarray(flat, dtype=np.int64), "d2": np.ascontiguousarray(d2, dtype=np.int64)}
    c = _PromptTablesC(self.n, self.H, self.W, keep["labels"].ctypes.data, present.ctypes.data, keep["area"].ctypes.data,
                       keep["cand_off"].ctypes.data, blob.ctypes.data, toff.ctypes.data, keep["flat"].ctypes.data,
                       int(keep["flat"].size), keep["d2"].ctypes.data, self.label_mask.data_ptr())
    self._native[key] = (c, keep)
    return c


PromptTables.native_tables = _native_tables
PromptTables._native_tables_locked = _native_tables_locked


def pouch_prompts(tables: PromptTables, ca_frames: np.ndarray, rng, paths: Optional[List[Optional[str]]] = None,
                  threshold: float = 0.1, quirk: bool = True, num_negative: int = 5, min_distance_from_cells: int = 5,
                  min_cell_area: int = 10, return_text: bool = False):
    """:func:`frame_prompts` for ALL sampled frames of one pouch (``ca_frames`` ``[F, n]``) in one native call outside the
    interpreter lock (``cab200_prompts_pouch``): the frames' K4 passes, numpy's own draw / percentile algorithms on
    ``rng``'s Mersenne-Twister state (advanced exactly as the per-frame calls would), and the prompt files written to
    ``paths[f]`` (``None`` / missing: not written).  Returns ``(n_entries [F] -- -1 for a frame without any active
    cell, for which the reference writes nothing --, texts or None)``."""
    name, key, pos, has_gauss, cached = rng.get_state()
    if name != "MT19937":
        raise ValueError("legacy MT19937 generator expected")
    ca = np.ascontiguousarray(ca_frames, dtype=np.float64)
    F, n = ca.shape
    if n != tables.n:
        raise ValueError(f"ca_frames has {n} cells, the tables {tables.n}")
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    cpos = ctypes.c_int32(int(pos))
    nent = np.zeros(max(F, 1), dtype=np.int32)
    if paths is not None:
        enc = [(p.encode() if p else b"") + b"\0" for p in paths]
        pblob = b"".join(enc)
        poff = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    text_buf = text_off = None
    if return_text:
        text_buf = np.empty(F * (tables.n * 160 + 4096) + 16, dtype=np.uint8)
        text_off = np.zeros(F + 1, dtype=np.int64)
    ws = tables.acquire_workspace()
    dev = tables.label_mask.device
    lib = _lib.load()
    try:
        with torch.cuda.device(dev):
            ws.stream.wait_stream(torch.cuda.current_stream(dev))
            rc = lib.cab200_prompts_pouch(
                ctypes.addressof(tables.native_tables(int(min_distance_from_cells))), F, ca.ctypes.data, float(threshold),
                1 if quirk else 0, int(num_negative), int(min_distance_from_cells), int(min_cell_area), key.ctypes.data,
                ctypes.addressof(cpos), pblob if paths is not None else None, poff.ctypes.data if paths is not None else None,
                ws.out.data_ptr(), ws.scratch.data_ptr(), ws.scratch.numel(), ws.luts.data_ptr(), ws.host.data_ptr(),
                ws.host_luts.data_ptr(), ws.stream.cuda_stream, nent.ctypes.data,
                text_buf.ctypes.data if return_text else None, int(text_buf.size) if return_text else 0,
                text_off.ctypes.data if return_text else None)
            _lib.check(rc, "cab200_prompts_pouch")
    finally:
        tables.release_workspace(ws)
    rng.set_state((name, key, int(cpos.value), has_gauss, cached))
    texts = None
    if return_text:
        raw = text_buf.tobytes()
        texts = [raw[text_off[f]:text_off[f + 1]].decode("ascii") if nent[f] >= 0 else None for f in range(F)]
    return nent[:F], texts


def _fmt_point(pt, ind: str) -> str:
    return f"[\n{ind}  {pt[0]},\n{ind}  {pt[1]}\n{ind}]"


def _positive_entry(v: Dict) -> str:
    return (f'{{\n      "point": {_fmt_point(v["point"], "      ")},\n      "label": {v["label"]},\n'
            f'      "area": {v["area"]},\n      "confidence": {float.__repr__(v["confidence"])}\n    }}')


def _assemble(entries: List[str], neg: List[Dict]) -> str:
    """The file text from the positive entries (already formatted, in id order) and the negative prompts."""
    fr = float.__repr__
    pos = "{\n" + ",\n".join(entries) + "\n  }" if entries else "{}"
    if neg:
        body = ",\n".join(
            f'    {{\n      "point": {_fmt_point(v["point"], "      ")},\n      "label": {v["label"]},\n'
            f'      "confidence": {fr(v["confidence"])}\n    }}' for v in neg)
        ng = "[\n" + body + "\n  ]"
    else:
        ng = "[]"
    return (f'{{\n  "positive_prompts": {pos},\n  "negative_prompts": {ng},\n  "num_cells": {len(entries)},\n'
            f'  "version": "1.0"\n}}')


def dumps_prompts(prompts: Dict, negative_prompts: Optional[List] = None) -> str:
    """The text ``json.dump(data, f, indent=2)`` writes for the reference's prompt file
    (``utils/sam2_data_generator.py:228-238``), formatted directly: the pure-Python encoder ``indent`` switches
    ``json`` to costs more than computing the prompts (1 400 instances per frame of a fluttering pouch)."""
    return _assemble([f'    "{k}": {_positive_entry(v)}' for k, v in prompts.items()], negative_prompts or [])


def save_prompts(prompts: Dict, output_path: str, negative_prompts: Optional[List] = None) -> str:
    """Same file as the reference's ``save_prompts`` (``utils/sam2_data_generator.py:210-240``)."""
    os.makedirs(os.path.dirname(output_path), exist_ok=True)
    try:
        text = dumps_prompts(prompts, negative_prompts)
    except (KeyError, TypeError, IndexError):      # not the dicts frame_prompts builds: the generic encoder
        text = json.dumps({"positive_prompts": prompts, "negative_prompts": negative_prompts or [],
                           "num_cells": len(prompts), "version": "1.0"}, indent=2)
    with open(output_path, "w") as f:
        f.write(text)
    return output_path
