"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the reference's hot path, function by function, each citing the reference
``file:line`` it follows, plus ctypes wrappers around the plain-C restatement in
``oracle/csrc/oracle.c`` (fast enough for full-size ensembles and for the CPU baseline).

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference itself, generated in the build container by
``oracle/make_golden.py`` (imports ``/root/reference`` with the shims in
``oracle/reference_shims.py``) and committed under ``tests/golden/``.
``skimage.draw.polygon`` is a restatement of scikit-image's published algorithm (library not
installable offline): *cells_mask parity against real scikit-image is unpinned*.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

#: slot order of the packed per-pouch parameter vector (same as include/cab200.h)
PARAM_SLOTS = ["k_1", "k_a", "k_p", "k_2", "V_SERCA", "K_SERCA", "K_PLC", "K_5", "c_tot", "beta",
               "k_tau", "tau_max", "k_i", "D_c", "D_p", "dt"]
DT = 0.2                     # reference core/pouch.py:127
T_DEFAULT = int(3600 / DT)   # reference core/pouch.py:128,188
DEFAULT_FRAME_STEPS = list(range(0, 18000, 200))   # reference main.py:167


def pack_params(p: Dict[str, float], dt: float = DT) -> np.ndarray:
    """18-key reference dict -> 16 doubles.  ``D_c = D_c_ratio * D_p`` as reference
    ``core/pouch.py:194``."""
    q = dict(p)
    q["D_c"] = p["D_c_ratio"] * p["D_p"]
    q["dt"] = dt
    return np.array([q[k] for k in PARAM_SLOTS], dtype=np.float64)


def host_draws(p: Dict[str, float], n_cells: int, sim_number: int
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-cell V_PLC and r0 exactly as the reference draws them (legacy global ``np.random``).

    ``Pouch.__init__`` (``core/pouch.py:199-200``): seed(sim_number); V_PLC = U(lower, upper, (n,1)).
    ``Pouch.simulate`` (``core/pouch.py:317-331``): seed(sim_number) again; r0 = U(.5,.7,(n,1));
    k = max(1, int(frac*n)); initiators = choice(n, k, replace=False); V_PLC[init] = U(1.3,1.5,k).
    Returns (vplc (n,), r0 (n,), initiators).
    """
    np.random.seed(sim_number)
    vplc = np.random.uniform(p["lower"], p["upper"], (n_cells, 1))
    np.random.seed(sim_number)
    r0 = np.random.uniform(0.5, 0.7, size=(n_cells, 1))
    k = max(1, int(p["frac"] * n_cells))
    init = np.random.choice(n_cells, k, replace=False)
    vplc[init, 0] = np.random.uniform(1.3, 1.5, len(init))
    return vplc.flatten(), r0.flatten(), init


def step_numpy(ca, ipt, s, r, ca_lap, ipt_lap, V_PLC, dt, k_1, k_a, k_p, k_2, V_SERCA, K_SERCA,
               K_PLC, K_5, c_tot, beta, k_tau, tau_max, k_i):
    """One Euler step on ``(n,1)`` arrays: reference ``core/pouch.py:75-104`` evaluated with
    numpy's strict left-to-right IEEE semantics; ``x**3`` as ``(x*x)*x`` (numba's expansion of a
    constant integer power -- the shipped reference is the numba build)."""
    num = r * ca * ipt
    term = num / ((k_a + ca) * (k_p + ipt))
    J_channel = (k_1 * (term * term * term) + k_2) * (s - ca)
    ca_sq = ca * ca
    J_SERCA = V_SERCA * ca_sq / (ca_sq + K_SERCA * K_SERCA)
    ca_next = ca + dt * (ca_lap + J_channel - J_SERCA)
    J_PLC = V_PLC.reshape(-1, 1) * ca_sq / (ca_sq + K_PLC * K_PLC)
    ipt_next = ipt + dt * (ipt_lap + J_PLC - K_5 * ipt)
    s_next = (c_tot - ca_next) / beta
    ca_4 = ca_sq * ca_sq
    k_tau_4 = k_tau * k_tau * k_tau * k_tau
    recovery = (k_tau_4 + ca_4) / (tau_max * k_tau_4)
    inact = 1.0 - r * (k_i + ca) / k_i
    r_next = r + dt * recovery * inact
    return ca_next, ipt_next, s_next, r_next


def simulate_numpy(csr, prm: np.ndarray, c0, p0, r0, vplc, T: int = T_DEFAULT,
                   frame_steps: Sequence[int] = DEFAULT_FRAME_STEPS) -> np.ndarray:
    """Reference time loop (``core/pouch.py:338-366``) storing only ``frame_steps``.

    ``csr`` = (rowptr, colidx, vals); products via ``scipy.sparse.csr_matrix.dot`` like the
    reference (``:346-347``).  Returns frames ``[F, 4, n]``.
    """
    from scipy.sparse import csr_matrix
    rowptr, col, val = csr
    n = len(rowptr) - 1
    L = csr_matrix((val, col, rowptr), shape=(n, n))
    kw = dict(zip(PARAM_SLOTS, [float(x) for x in prm]))
    D_c, D_p, dt = kw.pop("D_c"), kw.pop("D_p"), kw.pop("dt")
    ca = np.asarray(c0, dtype=np.float64).reshape(-1, 1).copy()
    ipt = np.asarray(p0, dtype=np.float64).reshape(-1, 1).copy()
    s = (kw["c_tot"] - ca) / kw["beta"]                      # :322
    r = np.asarray(r0, dtype=np.float64).reshape(-1, 1).copy()
    V = np.asarray(vplc, dtype=np.float64).flatten()
    want = {int(t): i for i, t in enumerate(frame_steps)}
    out = np.zeros((len(frame_steps), 4, n))
    for step in range(T):
        if step in want:
            f = want[step]
            out[f, 0], out[f, 1], out[f, 2], out[f, 3] = ca[:, 0], ipt[:, 0], s[:, 0], r[:, 0]
        if step == T - 1:
            break
        ca_lap = D_c * L.dot(ca)
        ipt_lap = D_p * L.dot(ipt)
        ca, ipt, s, r = step_numpy(ca, ipt, s, r, ca_lap, ipt_lap, V, dt, **kw)
    return out


# ------------------------------------------------------------------------------------------
# rasterisation oracles
# ------------------------------------------------------------------------------------------

def _point_in_polygon(xp: np.ndarray, yp: np.ndarray, x: float, y: float) -> int:
    """scikit-image ``_shared/geometry.pxd: point_in_polygon`` (O'Rourke crossing test, eps 1e-12):
    0 outside, 1 inside, 2 vertex, 3 edge.  Third-party algorithm restated from its published
    source (scikit-image >= 0.18, the reference's ``requirements.txt:6`` lower bound)."""
    eps = 1e-12
    n = len(xp)
    x1, y1 = xp[n - 1] - x, yp[n - 1] - y
    r_cross = l_cross = 0
    for i in range(n):
        x0, y0 = xp[i] - x, yp[i] - y
        if -eps < x0 < eps and -eps < y0 < eps:
            return 2
        if (y0 > 0) != (y1 > 0):
            if (x0 * y1 - x1 * y0) / (y1 - y0) > 0:
                r_cross += 1
        if (y0 < 0) != (y1 < 0):
            if (x0 * y1 - x1 * y0) / (y1 - y0) < 0:
                l_cross += 1
        x1, y1 = x0, y0
    if (r_cross & 1) != (l_cross & 1):
        return 3
    return 1 if (r_cross & 1) else 0


def polygon_skimage(r, c, shape=None) -> Tuple[np.ndarray, np.ndarray]:
    """``skimage.draw.polygon(r, c, shape)`` restated (``skimage/draw/_draw.pyx: _polygon``):
    bounding box ``[int(max(0,min)), int(ceil(max))]`` clipped to ``shape``, row-major scan, keep
    every pixel whose ``point_in_polygon`` is non-zero (boundary inclusive).  Vectorised over the
    box but with the identical per-edge arithmetic."""
    r = np.asarray(r, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    minr, maxr = int(max(0, r.min())), int(np.ceil(r.max()))
    minc, maxc = int(max(0, c.min())), int(np.ceil(c.max()))
    if shape is not None:
        maxr, maxc = min(shape[0] - 1, maxr), min(shape[1] - 1, maxc)
    if maxr < minr or maxc < minc:
        return np.zeros(0, dtype=np.intp), np.zeros(0, dtype=np.intp)
    R, C = np.meshgrid(np.arange(minr, maxr + 1), np.arange(minc, maxc + 1), indexing="ij")
    y, x = R.astype(np.float64), C.astype(np.float64)
    eps = 1e-12
    nv = len(r)
    x1, y1 = c[nv - 1] - x, r[nv - 1] - y
    rc = np.zeros(R.shape, dtype=np.int64)
    lc = np.zeros(R.shape, dtype=np.int64)
    vertex = np.zeros(R.shape, dtype=bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(nv):
            x0, y0 = c[i] - x, r[i] - y
            is_v = (x0 > -eps) & (x0 < eps) & (y0 > -eps) & (y0 < eps)
            q = (x0 * y1 - x1 * y0) / (y1 - y0)
            live = ~vertex          # the scalar code returns at the first vertex hit
            rc += (live & ~is_v & ((y0 > 0) != (y1 > 0)) & (q > 0))
            lc += (live & ~is_v & ((y0 < 0) != (y1 < 0)) & (q < 0))
            vertex |= is_v
            x1, y1 = x0, y0
    inside = vertex | ((rc & 1) != (lc & 1)) | ((rc & 1) == 1)
    return R[inside], C[inside]


def scale_vertices(vertices, output_size) -> List[np.ndarray]:
    """Integer pixel polygons: reference ``core/pouch.py:272-281`` (== ``:430-446``)."""
    width, height = output_size
    all_x = np.concatenate([v[:, 0] for v in vertices])
    all_y = np.concatenate([v[:, 1] for v in vertices])
    x_min, x_max = np.min(all_x), np.max(all_x)
    y_min, y_max = np.min(all_y), np.max(all_y)
    out = []
    for v in vertices:
        sx = ((v[:, 0] - x_min) / (x_max - x_min) * (width - 1)).astype(int)
        sy = ((v[:, 1] - y_min) / (y_max - y_min) * (height - 1)).astype(int)
        out.append(np.column_stack([sx, sy]))
    return out


def cells_mask_oracle(vertices, output_size) -> np.ndarray:
    """``Pouch._create_cells_mask`` (reference ``core/pouch.py:260-287``): later cells overwrite."""
    width, height = output_size
    mask = np.zeros((height, width), dtype=np.int32)
    for cell_id, pts in enumerate(scale_vertices(vertices, output_size), start=1):
        rr, cc = polygon_skimage(pts[:, 1], pts[:, 0], shape=(height, width))
        mask[rr, cc] = cell_id
    return mask


def paint_label_map_cv2(vertices, output_size) -> np.ndarray:
    """The per-pixel "which cell painted last" map of ``generate_image``'s fill loop (reference
    ``core/pouch.py:441-453``), obtained with the real ``cv2.fillPoly`` on an int32 canvas in the
    same cell order.  ``image == LUT[label]`` then reproduces ``generate_image`` bit for bit."""
    import cv2
    width, height = output_size
    canvas = np.zeros((height, width), dtype=np.int32)
    for cell_id, pts in enumerate(scale_vertices(vertices, output_size), start=1):
        cv2.fillPoly(canvas, [pts.reshape((-1, 1, 2)).astype(np.int32)], color=int(cell_id))
    return canvas


def generate_image_oracle(vertices, calcium_activity, output_size) -> np.ndarray:
    """``Pouch.generate_image`` default branch, literally (reference ``core/pouch.py:412-453``):
    one ``cv2.fillPoly`` per cell in cell order with ``(gray, 255)``."""
    import cv2
    width, height = output_size
    img = np.zeros((height, width, 2), dtype=np.uint8)
    vmin = 0
    vmax = max(np.max(calcium_activity), 0.5)
    for cell_id, pts in enumerate(scale_vertices(vertices, output_size)):
        normalized = np.clip((calcium_activity[cell_id] - vmin) / (vmax - vmin), 0, 1)
        gray = int(255 * normalized)
        cv2.fillPoly(img, [pts.reshape((-1, 1, 2)).astype(np.int32)], color=(gray, 255))
    return img


def gray_lut_oracle(calcium_activity) -> np.ndarray:
    """Per-cell gray level of reference ``core/pouch.py:421-427``."""
    ca = np.asarray(calcium_activity, dtype=np.float64)
    vmax = max(np.max(ca), 0.5)
    return np.array([int(255 * np.clip((v - 0) / (vmax - 0), 0, 1)) for v in ca], dtype=np.uint8)


def active_cells_oracle(calcium, threshold: float = 0.1) -> List[int]:
    """``Pouch.get_active_cells`` (reference ``core/pouch.py:563-565``)."""
    return np.where(np.asarray(calcium) > threshold)[0].tolist()


def active_mask_oracle(cells_mask: np.ndarray, active_cells: List[int]) -> np.ndarray:
    """``Pouch.get_cell_masks(active_only=True)`` (reference ``core/pouch.py:540-546``)."""
    out = np.zeros_like(cells_mask)
    for cell_id in active_cells:
        out[cells_mask == (cell_id + 1)] = cell_id + 1
    return out


def instance_mask_oracle(cell_masks: np.ndarray, active_cells: Optional[List[int]]) -> np.ndarray:
    """``create_instance_mask_from_cells`` (reference ``utils/sam2_data_generator.py:17-40``),
    including its 0-based-vs-1-based id mismatch."""
    if active_cells is None:
        return cell_masks.astype(np.uint16)
    inst = np.zeros_like(cell_masks, dtype=np.uint16)
    for new_id, cell_id in enumerate(active_cells, start=1):
        inst[cell_masks == cell_id] = new_id
    return inst


# ------------------------------------------------------------------------------------------
# C restatement (oracle/csrc/oracle.c)
# ------------------------------------------------------------------------------------------

_LIB_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_LIB_DIR, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2, no contraction, no fast-math: strict IEEE double in source order."""
    src = os.path.join(_HERE, "csrc", "oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(_LIB_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", _LIB_PATH, src, "-lm", "-lpthread"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _ptr(a: Optional[np.ndarray], ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_simulate.restype = ctypes.c_int
        _lib.oracle_simulate_f32.restype = ctypes.c_int
        _lib.oracle_sim_batch.restype = ctypes.c_int
        _lib.oracle_raster_frame.restype = ctypes.c_int
    return _lib


def simulate_c(csr, prm, c0, p0, r0, vplc, T: int = T_DEFAULT,
               frame_steps: Sequence[int] = DEFAULT_FRAME_STEPS, f32: bool = False) -> np.ndarray:
    rowptr, col, val = (np.ascontiguousarray(csr[0], np.int32), np.ascontiguousarray(csr[1], np.int32),
                        np.ascontiguousarray(csr[2], np.float64))
    n = len(rowptr) - 1
    fs = np.ascontiguousarray(frame_steps, dtype=np.int32)
    out = np.zeros((len(fs), 4, n), dtype=np.float32 if f32 else np.float64)
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (prm, c0, p0, r0, vplc)]
    fn = lib().oracle_simulate_f32 if f32 else lib().oracle_simulate
    st = fn(ctypes.c_int(n), _ptr(rowptr, ctypes.c_int32), _ptr(col, ctypes.c_int32),
            _ptr(val, ctypes.c_double), *[_ptr(a, ctypes.c_double) for a in args],
            ctypes.c_int(T), ctypes.c_int(len(fs)), _ptr(fs, ctypes.c_int32),
            out.ctypes.data_as(ctypes.c_void_p))
    if st < 0:
        raise MemoryError("oracle_simulate")
    return out


def sim_batch_c(geom_id, g_ncell, g_rowptr_off, g_nnz_off, rowptr, col, val, cell_off, params,
                c0, p0, r0, vplc, T, frame_steps, n_threads: int = 1):
    """Ensemble on host threads; same argument meaning as ``cab200_sim_batch``.  Returns
    (frames_flat float64 [sum_p F*4*n_p], status int32 [P])."""
    geom_id = np.ascontiguousarray(geom_id, np.int32)
    g_ncell = np.ascontiguousarray(g_ncell, np.int32)
    g_rowptr_off = np.ascontiguousarray(g_rowptr_off, np.int64)
    g_nnz_off = np.ascontiguousarray(g_nnz_off, np.int64)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    cell_off = np.ascontiguousarray(cell_off, np.int64)
    params = np.ascontiguousarray(params, np.float64)
    c0, p0, r0, vplc = [np.ascontiguousarray(a, np.float64) for a in (c0, p0, r0, vplc)]
    fs = np.ascontiguousarray(frame_steps, np.int32)
    P = len(geom_id)
    frames = np.zeros(int(cell_off[-1]) * 4 * len(fs), dtype=np.float64)
    status = np.zeros(P, dtype=np.int32)
    lib().oracle_sim_batch(
        ctypes.c_int(P), _ptr(geom_id, ctypes.c_int32), ctypes.c_int(len(g_ncell)),
        _ptr(g_ncell, ctypes.c_int32), _ptr(g_rowptr_off, ctypes.c_int64),
        _ptr(g_nnz_off, ctypes.c_int64), _ptr(rowptr, ctypes.c_int32), _ptr(col, ctypes.c_int32),
        _ptr(val, ctypes.c_double), _ptr(cell_off, ctypes.c_int64), _ptr(params, ctypes.c_double),
        _ptr(c0, ctypes.c_double), _ptr(p0, ctypes.c_double), _ptr(r0, ctypes.c_double),
        _ptr(vplc, ctypes.c_double), ctypes.c_int(T), ctypes.c_int(len(fs)),
        _ptr(fs, ctypes.c_int32), _ptr(frames, ctypes.c_double), _ptr(status, ctypes.c_int32),
        ctypes.c_int(n_threads))
    return frames, status


def raster_frame_c(ca, label_img, label_mask, threshold: float = 0.1, quirk: bool = True,
                   want_active: bool = True):
    """(image (H,W,2) u8, active (H,W) i32 or None, instance (H,W) u16, n_active)."""
    ca = np.ascontiguousarray(ca, np.float64)
    li = np.ascontiguousarray(label_img, np.uint16)
    lm = np.ascontiguousarray(label_mask, np.uint16)
    H, W = li.shape
    img = np.zeros((H, W, 2), np.uint8)
    act = np.zeros((H, W), np.int32) if want_active else None
    inst = np.zeros((H, W), np.uint16)
    n_act = lib().oracle_raster_frame(
        ctypes.c_int(len(ca)), _ptr(ca, ctypes.c_double), _ptr(li, ctypes.c_uint16),
        _ptr(lm, ctypes.c_uint16), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_double(threshold),
        ctypes.c_int(1 if quirk else 0), _ptr(img, ctypes.c_uint8),
        _ptr(act, ctypes.c_int32), _ptr(inst, ctypes.c_uint16))
    return img, act, inst, n_act
