"""Ensemble driver: many independent pouches through the CUDA kernels in one go.

This is the batched form of what the reference does one pouch at a time in
``generate_simulation_batch``'s loop (reference ``main.py:202-390``): ``Pouch(...)`` ->
``simulate()`` -> per sampled time step ``generate_image`` / ``get_active_cells`` /
``get_cell_masks(active_only=True)`` / ``create_instance_mask_from_cells``.

Host side (numpy): parameter packing and the reference's random draws (same legacy ``np.random``
call sequence, so V_PLC / r0 / initiator cells are bit-identical).  Device side: K1
(``cab200_sim_batch``) and K2 (``cab200_raster_batch``) through the C ABI; torch only provides
device memory, pinned host memory and streams.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .geometry import Geometry, get_geometry

DT = 0.2                                  # reference core/pouch.py:127
DURATION = 3600                           # reference core/pouch.py:128
T_DEFAULT = int(DURATION / DT)            # 18000, reference core/pouch.py:188
DEFAULT_FRAME_STEPS = list(range(0, 18000, 200))   # reference main.py:167
REQUIRED_PARAMS = ["D_c_ratio", "D_p", "K_5", "K_PLC", "K_SERCA", "V_SERCA", "beta", "c_tot",
                   "frac", "k_1", "k_2", "k_a", "k_i", "k_p", "k_tau", "lower", "tau_max", "upper"]


def pack_params(p: Dict[str, float], dt: float = DT) -> np.ndarray:
    """18-key dict -> the 16 doubles K1 consumes; ``D_c = D_c_ratio * D_p`` (core/pouch.py:194)."""
    q = dict(p)
    q["D_c"] = p["D_c_ratio"] * p["D_p"]
    q["dt"] = dt
    return np.array([q[k] for k in _lib.PARAM_SLOTS], dtype=np.float64)


def draw_vplc(p: Dict[str, float], n_cells: int, sim_number: int) -> np.ndarray:
    """``Pouch.__init__``'s per-cell V_PLC draw (reference ``core/pouch.py:199-200``), ``(n,1)``."""
    np.random.seed(sim_number)
    return np.random.uniform(p["lower"], p["upper"], (n_cells, 1))


def draw_initial_state(p: Dict[str, float], n_cells: int, sim_number: int, vplc: np.ndarray
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """``Pouch.simulate``'s preamble (reference ``core/pouch.py:317-331``): reseed, r0 ~ U(.5,.7),
    initiator cells get V_PLC ~ U(1.3,1.5).  Mutates ``vplc`` in place like the reference mutates
    ``self.V_PLC``; returns (r0 (n,), initiator ids)."""
    np.random.seed(sim_number)
    r0 = np.random.uniform(0.5, 0.7, size=(n_cells, 1))
    n_init = max(1, int(p["frac"] * n_cells))
    init = np.random.choice(n_cells, n_init, replace=False)
    vplc[init, 0] = np.random.uniform(1.3, 1.5, len(init))
    return r0.flatten(), init


def validate_for_simulation(p: Dict[str, float]) -> None:
    """The two ``ValueError`` checks at the top of ``Pouch.simulate`` (core/pouch.py:309-313)."""
    if p["c_tot"] <= 0 or p["beta"] <= 0 or p["D_p"] <= 0:
        raise ValueError("Invalid simulation parameters: c_tot, beta, and D_p must be positive")
    if not 0 <= p["frac"] <= 1:
        raise ValueError(f"Invalid fraction of initiator cells: {p['frac']}. Must be between 0 and 1")


@dataclass
class PouchSpec:
    """One pouch of an ensemble, fully drawn on the host."""
    params: Dict[str, float]
    geometry: Geometry
    sim_number: int = 0
    vplc: Optional[np.ndarray] = None     # (n,) final per-cell V_PLC (initiators applied)
    r0: Optional[np.ndarray] = None       # (n,)
    c0: Optional[np.ndarray] = None       # (n,), default zeros (core/pouch.py:320)
    p0: Optional[np.ndarray] = None

    @classmethod
    def draw(cls, params: Dict[str, float], size: str = "medium", sim_number: int = 0,
             geometry_dir: Optional[str] = None) -> "PouchSpec":
        validate_for_simulation(params)
        geo = get_geometry(size, geometry_dir)
        v = draw_vplc(params, geo.n_cells, sim_number)
        r0, _ = draw_initial_state(params, geo.n_cells, sim_number, v)
        return cls(params=params, geometry=geo, sim_number=sim_number, vplc=v.flatten(), r0=r0)


# ---- per-device geometry cache ---------------------------------------------------------------

class _DeviceGeometry:
    def __init__(self, geo: Geometry, device: torch.device):
        rowptr, col, val = geo.csr
        self.geo = geo
        self.n = geo.n_cells
        self.max_row = int(np.diff(rowptr).max()) if self.n else 0
        offdiag = np.ones(len(col), dtype=bool)
        rows = np.repeat(np.arange(self.n), np.diff(rowptr))
        offdiag &= rows != col
        self.unit = bool(np.all(val[offdiag] == 1.0))
        if self.unit and self.n:
            # the unit-weight kernel adds the diagonal term from registers at one point per warp (csrc/k1_layout.cuh);
            # a geometry whose rows cannot be grouped that way (several diagonal entries in a row; rows with many
            # entries before and others with many after the diagonal in a geometry of one warp) runs the general kernel
            lay = np.zeros(4, dtype=np.int64)
            lib = _lib.load()
            if lib.cab200_sim_layout(self.n, 1, _lib.FP64_STRICT, _lib.as_i64p(lay)) == 0:
                probe = np.zeros(2 * self.n, dtype=np.uint16)
                if lib.cab200_plan_slots(self.n, _lib.as_i32p(rowptr), _lib.as_i32p(col), int(lay[2]), 1, int(lay[1]), 0, 1, 0,
                                         probe.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None) != 0:
                    self.unit = False
        self.nnz = len(col)
        if self.nnz == 0:        # a graph without any entry still needs non-null device pointers
            col, val = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.float64)
        self.rowptr = torch.from_numpy(rowptr).to(device)
        self.col = torch.from_numpy(col).to(device)
        self.val = torch.from_numpy(val).to(device)
        self.label_maps: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self.edge_maps: Dict[Tuple[int, int, int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self.plans: Dict[Tuple[int, int], torch.Tensor] = {}
        self._capacity: Dict[Tuple[int, int], int] = {}

    def plan(self, precision: int, n_cta: int = 1) -> torch.Tensor:
        """Device copy of the K1 layout plan (``cab200_plan_slots`` / ``cab200_plan_cluster``) for this
        geometry; pure layout (bank-conflict avoidance, cluster partition), computed once on the host
        and cached on disk."""
        t = self.plans.get((precision, n_cta))
        if t is None:
            t = torch.from_numpy(slot_plan(self.geo, self.unit, precision, n_cta=n_cta).view(np.int16)).to(
                self.rowptr.device)
            self.plans[(precision, n_cta)] = t
        return t

    def capacity(self, n_cta: int, precision: int = _lib.FP64_STRICT) -> int:
        """Clusters of ``n_cta`` CTAs of this geometry's kernel that are resident at once on the geometry's device
        (``cab200_cluster_capacity``; <= 0: that cluster size does not apply).  On a 148-SM B200: 74 clusters of 2,
        45 of 3, 33 of 4, 15 of 8 -- a cluster lives inside one GPC."""
        key = (n_cta, precision)
        c = self._capacity.get(key)
        if c is None:
            with torch.cuda.device(self.rowptr.device):
                c = int(_lib.load().cab200_cluster_capacity(self.n, self.max_row, n_cta, precision))
            self._capacity[key] = c
        return c

    def auto_cluster(self, count: int = 1 << 30, n_sm: int = 148) -> int:
        """CTAs per pouch when the caller leaves the choice to the library: the option with the shortest predicted
        time for ``count`` pouches of this geometry, time = waves x step time, waves = ceil(count / resident pouches).
        Step times as measured on B200 (scripts/time_cluster.py): one CTA holds up to 1536 cells without spilling at
        0.88 ns per cell and step (1.32 us for 1500 cells), beyond that 1.94 ns per cell; a cluster CTA costs 1.4 ns per
        cell and step plus nothing else (the halo exchange latency sits inside that figure) -- so a lone medium pouch
        finishes in 13 ms on 4 CTAs instead of 24 ms on one, 148 of them are fastest on one CTA each, and 40 large
        pouches run faster as clusters of 3 (45 resident) than as clusters of 4 (33 resident, two waves).  Pouches of
        up to 1024 cells are latency bound per step and stay on one CTA; above 4 x 1536 cells only a cluster of 8
        fits (up to 12 288 cells)."""
        if not self.unit or self.n <= 1024:
            return 1
        if self.n > 4 * 1536:
            return 8
        count = max(1, min(int(count), 1 << 20))
        options = []
        if self.n <= 4096:
            per_cell = 0.88e-3 if self.n <= 1536 else 1.94e-3
            options.append((-(-count // n_sm) * per_cell * self.n, 1))
        for nc in (2, 3, 4):
            cells = -(-self.n // nc)
            if cells > 1536 or (nc == 3 and cells <= 1024):      # clusters of 3 exist for the 512 x 3 shape only
                continue
            cap = self.capacity(nc)
            if cap > 0:
                options.append((-(-count // cap) * 1.4e-3 * cells, nc))
        return min(options)[1] if options else 4


PLAN_ITERS = int(os.environ.get("CALCIUM_B200_PLAN_ITERS", "300000"))
_PLANS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plans")


def plan_layout(n: int, unit: bool, precision: int) -> Tuple[int, int]:
    """(value copies, cells per thread) the library will use for an n-cell geometry."""
    lay = np.zeros(4, dtype=np.int64)
    _lib.check(_lib.load().cab200_sim_layout(n, 1 if unit else 0, precision, _lib.as_i64p(lay)),
               "cab200_sim_layout")
    return int(lay[2]), int(lay[1])


def plan_path(geo: Geometry, copies: int, cpt: int, tracked: bool, n_cta: int = 1, entry_bytes: int = 16) -> str:
    """Plans are keyed by the CSR structure and the layout they were made for.  ``tracked`` plans
    (long offline annealing, scripts/make_plans.py) ship with the package under ``plans/``; others
    are computed on first use and cached next to the synthetic geometry cache."""
    rowptr, col, _ = geo.csr
    h = hashlib.sha1()
    h.update(rowptr.tobytes()); h.update(col.tobytes())
    h.update(f"copies={copies}:cpt={cpt}:v5-warpdiag".encode() if n_cta == 1         # v5: one diagonal point per warp
             else f"copies={copies}:cpt={cpt}:ncta={n_cta}:v3-warpdiag".encode())
    if entry_bytes != 16:
        h.update(f":entry={entry_bytes}".encode())      # FP32 mode: 8-byte entries, another bank-conflict model
    name = f"plan_{h.hexdigest()[:16]}.npy"
    if tracked:
        return os.path.join(_PLANS_DIR, name)
    from .geometry import _synthetic_cache_path
    return os.path.join(os.path.dirname(_synthetic_cache_path("x")), name)


def cluster_layout(n: int, n_cta: int, precision: int) -> Tuple[int, int]:
    lay = np.zeros(4, dtype=np.int64)
    _lib.check(_lib.load().cab200_cluster_layout(n, n_cta, precision, _lib.as_i64p(lay)),
               "cab200_cluster_layout")
    return int(lay[2]), int(lay[1])


def slot_plan(geo: Geometry, unit: bool, precision: int, iters: Optional[int] = None,
              n_cta: int = 1) -> np.ndarray:
    """uint16 plan for ``geo``: ``cab200_plan_slots`` (one CTA per pouch, ``[2n]``) or
    ``cab200_plan_cluster`` (``n_cta`` CTAs per pouch); see include/cab200.h."""
    lib = _lib.load()
    rowptr, col, _ = geo.csr
    n = geo.n_cells
    copies, cpt = plan_layout(n, unit, precision) if n_cta == 1 else cluster_layout(n, n_cta, precision)
    # one CTA per pouch: the plan is annealed against the bank-conflict model of the entry size the kernel loads
    # (16-byte (c, p) pairs in FP64, 8-byte ones in the FP32 mode); cluster plans use the FP64 model for both
    entry_bytes = 8 if (precision == _lib.FP32 and n_cta == 1) else 16
    paths = [plan_path(geo, copies, cpt, True, n_cta, entry_bytes), plan_path(geo, copies, cpt, False, n_cta, entry_bytes)]
    if iters is None:
        for path in paths:
            if os.path.exists(path):
                try:
                    plan = np.load(path)
                    if plan.dtype == np.uint16 and plan.ndim == 1 and (n_cta > 1 or plan.shape == (2 * n,)):
                        return plan
                except Exception:  # noqa: BLE001 - a corrupt cache entry is just recomputed
                    pass
    n_it = PLAN_ITERS if iters is None else iters
    if n_cta == 1:
        plan = np.zeros(2 * n, dtype=np.uint16)
        _lib.check(lib.cab200_plan_slots_entry(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1 if unit else 0, cpt,
                                               entry_bytes, n_it, 1, 0,
                                               plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None),
                   "cab200_plan_slots_entry")
    else:
        cap = 3 * n + n_cta + n_cta * n
        plan = np.zeros(cap, dtype=np.uint16)
        got = lib.cab200_plan_cluster(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), n_cta, copies, cpt, n_it, 1,
                                      plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), cap, None)
        if got < 0:
            _lib.check(int(got), "cab200_plan_cluster")
        plan = plan[:int(got)].copy()
    if iters is None:
        try:
            os.makedirs(os.path.dirname(paths[1]), exist_ok=True)
            tmp = paths[1] + f".{os.getpid()}.tmp.npy"
            np.save(tmp, plan)
            os.replace(tmp, paths[1])
        except OSError:
            pass
    return plan


_DEV_GEO: Dict[Tuple[int, int], _DeviceGeometry] = {}
_DEV_GEO_LOCK = threading.Lock()


def resolve_device(device) -> torch.device:
    """``torch.device`` with an explicit index (``torch.device("cuda")`` means the CURRENT device, which is not
    necessarily device 0: caches keyed on the index must not mix the two up)."""
    d = torch.device(device) if device is not None else torch.device("cuda")
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def device_geometry(geo: Geometry, device: torch.device) -> _DeviceGeometry:
    device = resolve_device(device)
    key = (id(geo), device.index)
    with _DEV_GEO_LOCK:
        dg = _DEV_GEO.get(key)
        if dg is None:
            dg = _DeviceGeometry(geo, device)
            _DEV_GEO[key] = dg
        return dg


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def build_label_maps(geo: Geometry, output_size: Tuple[int, int], device: torch.device
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two static uint16 label maps of one (geometry, output size), both built on ``device``.

    ``label_mask`` = ``Pouch.cells_mask`` (reference ``core/pouch.py:260-287``): K0,
    ``cab200_cells_mask`` (scikit-image's polygon rule).
    ``label_img`` = which cell's ``cv2.fillPoly`` painted each pixel last in ``generate_image``'s
    loop (reference ``core/pouch.py:441-453``): K0b, ``cab200_paint_label_map`` (OpenCV's fillPoly
    rule: outlines + scanline interior).  Rendering a frame is then a pure gather (K2).
    """
    dg = device_geometry(geo, device)
    key = (int(output_size[0]), int(output_size[1]))
    if key in dg.label_maps:
        return dg.label_maps[key]
    width, height = key
    if geo.n_cells > 65534:
        raise ValueError("label maps are uint16: at most 65534 cells")
    if (height * width) % 2:
        raise ValueError("output_size with an odd pixel count is not supported")
    offs, pts = geo.scaled_polygons(key)
    lib = _lib.load()
    d_offs = torch.from_numpy(offs).to(device)
    d_pts = torch.from_numpy(pts).to(device)
    label_img = torch.empty((height, width), dtype=torch.int16, device=device)
    label_mask = torch.empty((height, width), dtype=torch.int16, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.cab200_paint_label_map(geo.n_cells, d_offs.data_ptr(), d_pts.data_ptr(), height,
                                              width, label_img.data_ptr(), _stream_ptr(device)),
                   "cab200_paint_label_map")
        _lib.check(lib.cab200_cells_mask(geo.n_cells, d_offs.data_ptr(), d_pts.data_ptr(), height,
                                         width, label_mask.data_ptr(), _stream_ptr(device)),
                   "cab200_cells_mask")
        torch.cuda.current_stream(device).synchronize()
    dg.label_maps[key] = (label_img, label_mask)
    return dg.label_maps[key]


def edge_mode(with_border: bool, edge_blur: bool) -> int:
    """K2b mode for ``generate_image``'s optional branches (reference ``core/pouch.py:458-481``); 0 = default."""
    if edge_blur:
        return _lib.EDGE_BLUR_BORDER if with_border else _lib.EDGE_BLUR
    return _lib.EDGE_BORDER if with_border else 0


def blur_code(blur_type: str) -> int:
    """'motion' or -- like the reference's ``_create_blur_kernel`` (``:494-513``) -- mean for anything else."""
    return _lib.BLUR_MOTION if blur_type == "motion" else _lib.BLUR_MEAN


def build_edge_maps(geo: Geometry, output_size: Tuple[int, int], device: torch.device,
                    blur_kernel_size: int = 3, blur_type: str = "mean"
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The three static uint8 maps of ``generate_image``'s edge branches for one (geometry, output size,
    blur kernel), built on ``device`` by K0c (``cab200_edge_maps``): ``edges`` = the reference's
    ``edges_mask`` (every cell outlined with ``cv2.polylines``, ``core/pouch.py:455-456``), ``dilated`` =
    ``cv2.dilate(cv2.filter2D(edges_mask, -1, kernel), ones((3, 3)))`` (``:459-461``), whose /255 is the
    blend weight, and ``alpha`` = channel 1 of the blended image (frame independent)."""
    dg = device_geometry(geo, device)
    ks, bt = int(blur_kernel_size), blur_code(blur_type)
    key = (int(output_size[0]), int(output_size[1]), ks, bt)
    if key in dg.edge_maps:
        return dg.edge_maps[key]
    width, height = key[:2]
    offs, pts = geo.scaled_polygons(key[:2])
    lib = _lib.load()
    d_offs = torch.from_numpy(offs).to(device)
    d_pts = torch.from_numpy(pts).to(device)
    limg, _ = build_label_maps(geo, key[:2], device)
    edges = torch.empty((height, width), dtype=torch.uint8, device=device)
    dilated = torch.empty((height, width), dtype=torch.uint8, device=device)
    alpha = torch.empty((height, width), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.cab200_edge_maps(geo.n_cells, d_offs.data_ptr(), d_pts.data_ptr(), limg.data_ptr(), height, width,
                                        ks, bt, edges.data_ptr(), dilated.data_ptr(), alpha.data_ptr(),
                                        _stream_ptr(device)),
                   "cab200_edge_maps")
        torch.cuda.current_stream(device).synchronize()
    dg.edge_maps[key] = (edges, dilated, alpha)
    return dg.edge_maps[key]


_RAW_PIN: Dict[int, torch.Tensor] = {}
_RAW_PIN_LOCK = threading.Lock()


def encode_raw_frames(images: torch.Tensor, arena_bytes_per_pixel: float = 1.0, keep_pinned: bool = False
                      ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """K3r (``cab200_encode_raw_batch``): the finished 16-bit gray + alpha ``.png`` of every frame of ``images``
    (device uint8 ``[..., H, W, 2]`` -- frames that went through the edge blur or the deterministic defects and are
    no longer flat-shaded) encoded on the device; what crosses PCIe is the files, not 1 MiB of raw pixels per frame.
    Returns ``(arena, offset, size, flags)`` on the host, the last three of the leading shape of ``images``.  An arena
    that turns out too small (``arena_bytes_per_pixel``; blurred 512 x 512 frames take 0.2-0.7 bytes per pixel) is
    allocated again at the size the encoder asked for.  ``keep_pinned``: return a view of the per-device pinned landing
    buffer instead of a copy (the caller must be done with it before the next call on this device)."""
    lib = _lib.load()
    dev = images.device
    if images.dtype != torch.uint8 or images.dim() < 3 or images.shape[-1] != 2 or not images.is_contiguous():
        raise ValueError("encode_raw_frames expects a contiguous uint8 tensor [..., H, W, 2]")
    lead = tuple(images.shape[:-3])
    h, w = int(images.shape[-3]), int(images.shape[-2])
    n = int(np.prod(lead)) if lead else 1
    with torch.cuda.device(dev):
        n_slots = 2 * torch.cuda.get_device_properties(dev).multi_processor_count
        nb = int(lib.cab200_encode_raw_scratch_bytes(h, w, n_slots))
        if nb == 0:
            raise _lib.Cab200Error(f"cab200_encode_raw_scratch_bytes: unsupported frame size {w}x{h}")
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        records = torch.empty((max(n, 1), 2), dtype=torch.int64, device=dev)
        cursor = torch.zeros(2, dtype=torch.int64, device=dev)
        cap = (int(n * h * w * arena_bytes_per_pixel) + (1 << 20) + 15) // 16 * 16
        for _ in range(2):
            arena = torch.empty(cap, dtype=torch.uint8, device=dev)
            _lib.check(lib.cab200_encode_raw_batch(images.data_ptr(), n, h, w, arena.data_ptr(), cap, cursor.data_ptr(),
                                                   records.data_ptr(), scratch.data_ptr(), nb, n_slots, _stream_ptr(dev)),
                       "cab200_encode_raw_batch")
            used = int(cursor[0].item())
            if used <= cap:
                break
            cap = (used + (1 << 20) + 15) // 16 * 16
        rec = records[:n].cpu().numpy()
        # pinned landing buffer, cached per device (a pageable copy of half a gigabyte runs at a fraction of PCIe speed,
        # and a pinned allocation costs 0.3 s per GB)
        with _RAW_PIN_LOCK:
            pin = _RAW_PIN.pop(dev.index, None)
        if pin is None or pin.numel() < used:
            pin = torch.empty(max(used, 1 << 20) * 5 // 4, dtype=torch.uint8, pin_memory=True)
        pin[:used].copy_(arena[:used], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = pin[:used].numpy() if keep_pinned else pin[:used].numpy().copy()
        with _RAW_PIN_LOCK:
            _RAW_PIN[dev.index] = pin
    off = rec[:, 0].copy().reshape(lead)
    size = (rec[:, 1] & 0xFFFFFFFF).reshape(lead)
    flags = ((rec[:, 1] >> 32) & 0xFFFFFFFF).reshape(lead)
    if np.any(flags != _lib.ENC_OK):
        raise _lib.Cab200Error("cab200_encode_raw_batch: a frame did not fit its scratch slot or the arena")
    return host, off, size, flags


def _u16_numpy(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint16)


# ---- the ensemble ------------------------------------------------------------------------------

class Ensemble:
    """P independent pouches on one GPU.

    ``simulate()`` integrates all of them with one ``cab200_sim_batch`` call; ``rasterize()`` turns
    the captured frames into images / instance masks with ``cab200_raster_batch``.  All results
    stay on the device until asked for.
    """

    def __init__(self, specs: Sequence[PouchSpec], *, output_size: Tuple[int, int] = (512, 512),
                 T: int = T_DEFAULT, frame_steps: Optional[Sequence[int]] = None,
                 device: Optional[torch.device] = None, precision: int = _lib.FP64_STRICT,
                 dt: float = DT, use_plan: bool = True, cluster="auto", coalesced_capture="auto",
                 huffman: str = "tuned"):
        if not torch.cuda.is_available():
            raise RuntimeError("calcium_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = resolve_device(device)
        self.specs = list(specs)
        self.P = len(self.specs)
        self.T = int(T)
        self.dt = dt
        self.precision = precision
        self.output_size = (int(output_size[0]), int(output_size[1]))
        fs = DEFAULT_FRAME_STEPS if frame_steps is None else list(frame_steps)
        self.frame_steps = np.array(sorted(int(t) for t in fs if 0 <= int(t) < self.T), dtype=np.int32)
        self.F = len(self.frame_steps)
        self._frame_index = {int(t): i for i, t in enumerate(self.frame_steps)}

        # geometry table
        geos: List[Geometry] = []
        gid = np.zeros(self.P, dtype=np.int32)
        for i, s in enumerate(self.specs):
            for j, g in enumerate(geos):
                if g is s.geometry:
                    gid[i] = j
                    break
            else:
                geos.append(s.geometry)
                gid[i] = len(geos) - 1
        self.geometries = geos
        self.geom_id = gid
        self.dgeo = [device_geometry(g, self.device) for g in geos]
        self.g_ncell = np.array([d.n for d in self.dgeo], dtype=np.int32)
        self.g_max_row = np.array([d.max_row for d in self.dgeo], dtype=np.int32)
        self.g_unit = np.array([1 if d.unit else 0 for d in self.dgeo], dtype=np.int32)
        rp_sizes = [d.n + 1 for d in self.dgeo]
        nnz_sizes = [int(d.col.numel()) for d in self.dgeo]
        self.g_rowptr_off = np.concatenate([[0], np.cumsum(rp_sizes)[:-1]]).astype(np.int64)
        self.g_nnz_off = np.concatenate([[0], np.cumsum(nnz_sizes)[:-1]]).astype(np.int64)
        if len(self.dgeo) == 1:
            self.d_rowptr, self.d_col, self.d_val = self.dgeo[0].rowptr, self.dgeo[0].col, self.dgeo[0].val
        else:
            self.d_rowptr = torch.cat([d.rowptr for d in self.dgeo])
            self.d_col = torch.cat([d.col for d in self.dgeo])
            self.d_val = torch.cat([d.val for d in self.dgeo])
        ncell = self.g_ncell[gid].astype(np.int64)
        self.cell_off = np.concatenate([[0], np.cumsum(ncell)]).astype(np.int64)
        self.total_cells = int(self.cell_off[-1])
        # CTAs per pouch: "auto" (one CTA up to 1536 cells, else a cluster), or 1 / 2 / 4 for all
        n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        counts = np.bincount(gid, minlength=len(self.dgeo))
        self.g_cluster = np.array([d.auto_cluster(int(counts[gi]), n_sm) if cluster == "auto"
                                   else (int(cluster) if d.unit else 1)
                                   for gi, d in enumerate(self.dgeo)], dtype=np.int32)
        self.use_plan = use_plan
        self.huffman = huffman                          # K3's DEFLATE code: "tuned" (static dynamic-Huffman block) | "fixed"
        self.coalesced_capture = coalesced_capture      # "auto" | True | False, see _alloc_frames_tmp
        self._frames_tmp: Optional[torch.Tensor] = None
        self.d_plan: Optional[torch.Tensor] = None
        self.g_plan_off = np.full(len(self.dgeo), -1, dtype=np.int64)
        plans, off = [], 0
        for gi, d in enumerate(self.dgeo):
            nc = int(self.g_cluster[gi])
            if use_plan or nc > 1:          # a cluster cannot run without its partition plan
                try:
                    t = d.plan(precision, nc)
                except _lib.Cab200Error:
                    # some part of the cluster partition has rows that share no diagonal point (k1_layout.cuh): the
                    # pouch stays on one CTA when the library chose the cluster itself and one CTA can hold it
                    if nc == 1 or cluster != "auto" or d.n > 4096:
                        raise
                    nc = 1
                    self.g_cluster[gi] = 1
                    if not use_plan:
                        continue
                    t = d.plan(precision, 1)
                self.g_plan_off[gi] = off
                plans.append(t)
                off += int(t.numel())
        if plans:
            self.d_plan = plans[0] if len(plans) == 1 else torch.cat(plans)

        # host staging (pinned) for the per-pouch inputs: params | c0 | p0 | r0 | vplc
        self._h_in = torch.empty(self.P * _lib.NPARAM + 4 * self.total_cells, dtype=torch.float64,
                                 pin_memory=True)
        h = self._h_in.numpy()
        hp = h[: self.P * _lib.NPARAM].reshape(self.P, _lib.NPARAM)
        hc = h[self.P * _lib.NPARAM:].reshape(4, self.total_cells)
        for i, s in enumerate(self.specs):
            a, b = self.cell_off[i], self.cell_off[i + 1]
            n = b - a
            hp[i] = pack_params(s.params, dt)
            hc[0, a:b] = 0.0 if s.c0 is None else np.asarray(s.c0, dtype=np.float64).reshape(n)
            hc[1, a:b] = 0.0 if s.p0 is None else np.asarray(s.p0, dtype=np.float64).reshape(n)
            hc[2, a:b] = np.asarray(s.r0, dtype=np.float64).reshape(n)
            hc[3, a:b] = np.asarray(s.vplc, dtype=np.float64).reshape(n)
        self.h2d_bytes = self._h_in.numel() * 8
        self._d_in: Optional[torch.Tensor] = None
        self.frames: Optional[torch.Tensor] = None
        self.status: Optional[torch.Tensor] = None
        self._scratch = torch.empty(
            int(self.lib.cab200_sim_scratch_bytes(self.P, self.F)
                + self.lib.cab200_raster_scratch_bytes(self.P)
                + 2 * self.lib.cab200_encode_scratch_bytes(self.P, len(self.dgeo))) + 128,   # two halves: K3 chunks on two streams
            dtype=torch.uint8, device=self.device)
        self._label_img: Optional[torch.Tensor] = None
        self._edge_maps: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._label_mask: Optional[torch.Tensor] = None
        self._enc_tables: Optional[torch.Tensor] = None
        self._enc_geom: Optional[Tuple[torch.Tensor, np.ndarray, np.ndarray]] = None

    # -- sizes ---------------------------------------------------------------------------------
    @property
    def cell_timesteps(self) -> int:
        """Work of one ``simulate()``: sum over pouches of n * (T - 1)."""
        return self.total_cells * (self.T - 1)

    @property
    def frame_elems(self) -> int:
        return self.total_cells * 4 * self.F

    # -- K1 ------------------------------------------------------------------------------------
    def _alloc_frames_tmp(self) -> Optional[torch.Tensor]:
        """Staging for the coalesced capture path (``frames_tmp`` of ``cab200_sim_batch``): used when a
        large share of the steps is captured (the reference's full ``disc_dynamics`` history), if the
        slot-ordered copy of the largest same-geometry group fits in free device memory."""
        if self.coalesced_capture is False or self.F == 0:
            return None
        if self.coalesced_capture == "auto" and self.F * 8 < self.T:
            return None
        need = 0
        for g, d in enumerate(self.dgeo):
            if self.g_plan_off[g] < 0:
                continue
            cnt = int((self.geom_id == g).sum())
            need = max(need, int(self.lib.cab200_sim_frames_tmp_bytes(d.n, int(self.g_cluster[g]), int(d.unit),
                                                                      self.precision, cnt, self.F)))
        free, _ = torch.cuda.mem_get_info(self.device)
        if need == 0 or need > free - (2 << 30):
            return None
        return torch.empty(need, dtype=torch.uint8, device=self.device)

    def upload(self) -> None:
        """Host -> device copy of params / initial state (from pinned memory, async)."""
        with torch.cuda.device(self.device):
            if self._d_in is None:
                self._d_in = torch.empty(self._h_in.numel(), dtype=torch.float64, device=self.device)
            self._d_in.copy_(self._h_in, non_blocking=True)

    def simulate(self, upload: bool = True) -> torch.Tensor:
        """Integrate every pouch for T-1 steps; returns the flat frames tensor (device)."""
        if upload or self._d_in is None:
            self.upload()
        real = torch.float64 if self.precision == _lib.FP64_STRICT else torch.float32
        with torch.cuda.device(self.device):
            # kernels on side streams that still read the frames this launch overwrites (K3 chunks of
            # run_files_to_host): waited for here, after the upload, so that the upload overlaps them
            readers = getattr(self, "_frames_readers", None)
            if readers:
                cur_ = torch.cuda.current_stream(self.device)
                for ev_ in readers:
                    cur_.wait_event(ev_)
                self._frames_readers = []
            if self.frames is None or self.frames.dtype != real:
                self.frames = torch.empty(max(self.frame_elems, 1), dtype=real, device=self.device)
                self.status = torch.zeros(self.P, dtype=torch.int32, device=self.device)
                self._frames_tmp = self._alloc_frames_tmp()
            d = self._d_in
            np_ = self.P * _lib.NPARAM
            tc = self.total_cells
            base = d.data_ptr()
            rc = self.lib.cab200_sim_batch(
                self.P, _lib.as_i32p(self.geom_id), len(self.dgeo), _lib.as_i32p(self.g_ncell),
                _lib.as_i32p(self.g_max_row), _lib.as_i32p(self.g_unit),
                _lib.as_i64p(self.g_rowptr_off), _lib.as_i64p(self.g_nnz_off),
                self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.d_val.data_ptr(),
                _lib.as_i32p(self.g_cluster), _lib.as_i64p(self.g_plan_off),
                self.d_plan.data_ptr() if self.d_plan is not None else None,
                _lib.as_i64p(self.cell_off),
                base, base + 8 * np_, base + 8 * (np_ + tc), base + 8 * (np_ + 2 * tc),
                base + 8 * (np_ + 3 * tc),
                self.T, self.F, _lib.as_i32p(self.frame_steps),
                self.frames.data_ptr(), self.status.data_ptr(), self.precision,
                self._scratch.data_ptr(), self.lib.cab200_sim_scratch_bytes(self.P, self.F),
                self._frames_tmp.data_ptr() if self._frames_tmp is not None else None,
                self._frames_tmp.numel() if self._frames_tmp is not None else 0,
                _stream_ptr(self.device))
            _lib.check(rc, "cab200_sim_batch")
        return self.frames

    def continue_from_last_frame(self) -> None:
        """Makes the state of the last captured frame (it must be step T-1) the initial state of the next
        ``simulate(upload=False)``: c, p and r are copied on the device into the input block; s is a function of c
        (``core/pouch.py:95`` and ``:322`` are the same expression), so the continuation is bit-identical to one
        long run.  This is how horizons whose every-step history exceeds HBM are run: segment by segment, each
        segment's frames consumed (copied out, rasterised, or overwritten) before the next one lands."""
        if self.frames is None or self._d_in is None or self.F == 0 or int(self.frame_steps[-1]) != self.T - 1:
            raise RuntimeError("continue_from_last_frame needs a finished simulate() whose last captured step is T-1")
        if self.precision != _lib.FP64_STRICT:
            raise RuntimeError("continue_from_last_frame: FP64 only (FP32 frames are rounded copies of the inputs)")
        np_, tc = self.P * _lib.NPARAM, self.total_cells
        with torch.cuda.device(self.device):
            for g in range(len(self.dgeo)):
                idx = np.nonzero(self.geom_id == g)[0]
                n = int(self.g_ncell[g])
                contiguous = len(idx) > 0 and np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx)))
                groups = [idx] if contiguous else [np.array([i]) for i in idx]
                for grp in groups:
                    a = int(self.cell_off[grp[0]])
                    cnt = len(grp)
                    last = self.frames[a * 4 * self.F: (a + cnt * n) * 4 * self.F].view(cnt, self.F, 4, n)[:, -1]
                    for dst, state in ((0, 0), (1, 1), (2, 3)):           # c0 <- c, p0 <- p, r0 <- r
                        self._d_in[np_ + dst * tc + a: np_ + dst * tc + a + cnt * n].view(cnt, n).copy_(last[:, state])

    def frames_of(self, i: int) -> torch.Tensor:
        """Device view ``[F, 4, n_i]`` of pouch i's captured frames."""
        a, b = int(self.cell_off[i]), int(self.cell_off[i + 1])
        n = b - a
        return self.frames[a * 4 * self.F: b * 4 * self.F].view(self.F, 4, n)

    def frame_index(self, time_step: int) -> int:
        try:
            return self._frame_index[int(time_step)]
        except KeyError:
            raise KeyError(f"time step {time_step} was not captured; captured steps: "
                           f"{self.frame_steps[:8].tolist()}... ({self.F} frames)") from None

    def check_status(self) -> np.ndarray:
        st = self.status.cpu().numpy()
        if np.any(st & _lib.ST_BADGEOM):
            raise _lib.Cab200Error("integrator rejected the geometry (CSR row too long, column out "
                                   "of range, or off-diagonal weights declared unit but are not)")
        return st

    # -- K2 ------------------------------------------------------------------------------------
    def label_maps(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._label_img is None:
            maps = [build_label_maps(g, self.output_size, self.device) for g in self.geometries]
            self._label_img = torch.stack([m[0] for m in maps]).contiguous()
            self._label_mask = torch.stack([m[1] for m in maps]).contiguous()
        return self._label_img, self._label_mask

    def raster_out_bytes(self, image=True, instance=True, active=False, count: Optional[int] = None) -> int:
        w, h = self.output_size
        c = self.P if count is None else count
        return c * self.F * h * w * (2 * bool(image) + 2 * bool(instance) + 4 * bool(active))

    def edge_maps(self, blur_kernel_size: int = 3, blur_type: str = "mean"
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Per-geometry stacks ``[G][H][W]`` u8 of the static outline, blend-weight and alpha maps (K0c)."""
        key = (int(blur_kernel_size), blur_code(blur_type))
        if key not in self._edge_maps:
            maps = [build_edge_maps(g, self.output_size, self.device, key[0], blur_type) for g in self.geometries]
            self._edge_maps[key] = tuple(torch.stack([m[k] for m in maps]).contiguous() for k in range(3))
        return self._edge_maps[key]

    def rasterize(self, image: bool = True, instance: bool = True, active: bool = False,
                  pouches: Optional[Tuple[int, int]] = None, threshold: float = 0.1,
                  quirk: bool = True, out: Optional[Dict[str, torch.Tensor]] = None,
                  with_border: bool = False, edge_blur: bool = False, blur_kernel_size: int = 3,
                  blur_type: str = "mean") -> Dict[str, torch.Tensor]:
        """K2 over pouches ``[a, b)`` (default all).  Returns device tensors
        ``image [p,F,H,W,2] u8``, ``instance [p,F,H,W] i16 (bit pattern of uint16)``,
        ``active [p,F,H,W] i32``, ``n_active [p,F] i32``.  ``out`` may hold preallocated tensors.
        With ``edge_blur`` / ``with_border`` (``generate_image``'s optional branches, reference
        ``core/pouch.py:458-481``) the image comes from K2b (``cab200_raster_edges_batch``) instead."""
        if self.frames is None:
            raise RuntimeError("simulate() first")
        if self.precision != _lib.FP64_STRICT:
            raise RuntimeError("rasterize() reads float64 frames; run the FP64 mode")
        a, b = (0, self.P) if pouches is None else pouches
        cnt = b - a
        w, h = self.output_size
        limg, lmsk = self.label_maps()
        res: Dict[str, torch.Tensor] = {} if out is None else out
        with torch.cuda.device(self.device):
            def need(name, shape, dtype):
                t = res.get(name)
                if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                    t = torch.empty(shape, dtype=dtype, device=self.device)
                    res[name] = t
                return t
            img = need("image", (cnt, self.F, h, w, 2), torch.uint8) if image else None
            ins = need("instance", (cnt, self.F, h, w), torch.int16) if instance else None
            act = need("active", (cnt, self.F, h, w), torch.int32) if active else None
            nact = need("n_active", (cnt, self.F), torch.int32)
            gid = np.ascontiguousarray(self.geom_id[a:b])
            coff = np.ascontiguousarray(self.cell_off[a:b + 1])
            sim_scr = int(self.lib.cab200_sim_scratch_bytes(self.P, self.F))
            sim_scr = (sim_scr + 15) // 16 * 16
            mode = edge_mode(with_border, edge_blur) if image else 0
            rc = self.lib.cab200_raster_batch(
                cnt, _lib.as_i32p(gid), _lib.as_i32p(self.g_ncell), _lib.as_i64p(coff),
                limg.data_ptr(), lmsk.data_ptr(), h, w, self.F, self.frames.data_ptr(),
                float(threshold), 1 if quirk else 0,
                img.data_ptr() if (img is not None and not mode) else None,
                act.data_ptr() if act is not None else None,
                ins.data_ptr() if ins is not None else None,
                nact.data_ptr(), self._scratch.data_ptr() + sim_scr,
                self.lib.cab200_raster_scratch_bytes(cnt), _stream_ptr(self.device))
            _lib.check(rc, "cab200_raster_batch")
            if mode:
                edges, dilated, alpha = self.edge_maps(blur_kernel_size, blur_type)
                rc = self.lib.cab200_raster_edges_batch(
                    cnt, _lib.as_i32p(gid), _lib.as_i32p(self.g_ncell), _lib.as_i64p(coff),
                    limg.data_ptr(), edges.data_ptr(), dilated.data_ptr(), alpha.data_ptr(), h, w, self.F,
                    self.frames.data_ptr(),
                    mode, int(blur_kernel_size), blur_code(blur_type), img.data_ptr(),
                    self._scratch.data_ptr() + sim_scr, self.lib.cab200_raster_scratch_bytes(cnt),
                    _stream_ptr(self.device))
                _lib.check(rc, "cab200_raster_edges_batch")
        return res

    # -- K3 ------------------------------------------------------------------------------------
    def encode_tables(self) -> torch.Tensor:
        """Device copy of ``cab200_encode_tables`` for this ensemble's output size."""
        if self._enc_tables is None:
            w, h = self.output_size
            nb = int(self.lib.cab200_encode_tables_bytes())
            host = np.zeros(nb, dtype=np.uint8)
            _lib.check(self.lib.cab200_encode_tables(h, w, _lib.HUFFMAN_TUNED if self.huffman == "tuned"
                                                     else _lib.HUFFMAN_FIXED, host.ctypes.data), "cab200_encode_tables")
            self._enc_tables = torch.from_numpy(host).to(self.device)
        return self._enc_tables

    def encode_geometry_tables(self) -> Tuple[torch.Tensor, np.ndarray, np.ndarray]:
        """Static per-geometry tables of K3 (``cab200_encode_geometry``: checksum sums and event lists),
        their byte offsets and the largest event count per row segment."""
        if getattr(self, "_enc_geom", None) is None:
            w, h = self.output_size
            limg, lmsk = self.label_maps()
            tabs = self.encode_tables()
            sizes = [int(self.lib.cab200_encode_geometry_bytes(int(n), h, w)) for n in self.g_ncell]
            off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
            buf = torch.empty(int(sum(sizes)), dtype=torch.uint8, device=self.device)
            max_ev = np.zeros(len(sizes), dtype=np.int32)
            with torch.cuda.device(self.device):
                for g, n in enumerate(self.g_ncell):
                    one = np.zeros(1, dtype=np.int32)
                    _lib.check(self.lib.cab200_encode_geometry(
                        int(n), limg[g].data_ptr(), lmsk[g].data_ptr(), h, w, tabs.data_ptr(),
                        buf.data_ptr() + int(off[g]), _lib.as_i32p(one), _stream_ptr(self.device)),
                        "cab200_encode_geometry")
                    max_ev[g] = one[0]
            self._enc_geom = (buf, off, max_ev)
        return self._enc_geom

    def encode(self, image: bool = True, mask: bool = True, pouches: Optional[Tuple[int, int]] = None,
               threshold: float = 0.1, quirk: bool = True, out: Optional[Dict[str, torch.Tensor]] = None,
               arena_bytes: Optional[int] = None, stage_bytes: int = 0, parse: str = "events",
               scratch_half: int = 0) -> Dict[str, torch.Tensor]:
        """K3 over pouches ``[a, b)``: every frame rasterised and encoded on the device into the files
        the reference's batch path stores -- a 16-bit gray+alpha ``.png`` (``save_image``,
        utils/image_processing.py:252-269) and a ``.npz`` holding ``mask`` (the instance mask,
        utils/sam2_data_generator.py:203).  Returns device tensors ``arena`` (uint8), ``cursor``
        (int64[1], bytes used), ``records`` (int64 ``[p, F, 2, 2]``: for image / mask the arena offset
        and ``size | flags << 32``) and ``n_active [p, F]``.  See :func:`records_to_host`.
        ``parse="pixels"`` forces the per-pixel parse (same files; the event parse is the fast path).
        ``scratch_half`` (0 / 1): which half of the launch scratch (pouch table) the call uses -- calls that may
        run concurrently on two streams (:meth:`run_files_to_host`) must not share one."""
        if self.frames is None:
            raise RuntimeError("simulate() first")
        if self.precision != _lib.FP64_STRICT:
            raise RuntimeError("encode() reads float64 frames; run the FP64 mode")
        a, b = (0, self.P) if pouches is None else pouches
        cnt = b - a
        w, h = self.output_size
        limg, lmsk = self.label_maps()
        tabs = self.encode_tables()
        gtab, gtab_off, gmax_ev = self.encode_geometry_tables()
        res: Dict[str, torch.Tensor] = {} if out is None else out
        if arena_bytes is None:
            # a flat-shaded polygon frame encodes to ~1/10 of its raw size; 1/3 leaves ample room
            arena_bytes = max(1 << 20, cnt * self.F * h * w * 6 // 3)
        arena_bytes = (int(arena_bytes) + 15) // 16 * 16
        with torch.cuda.device(self.device):
            arena = res.get("arena")
            if arena is None or arena.numel() < arena_bytes:
                arena = torch.empty(arena_bytes, dtype=torch.uint8, device=self.device)
                res["arena"] = arena
            cursor = res.get("cursor")
            if cursor is None:
                res["cursor16"] = torch.zeros(2, dtype=torch.int64, device=self.device)   # 16 bytes: publishable
                cursor = res["cursor16"][:1]
                res["cursor"] = cursor
            else:
                cursor.zero_()
            rec = res.get("records")
            if rec is None or tuple(rec.shape) != (cnt, self.F, 2, 2):
                rec = torch.empty((cnt, self.F, 2, 2), dtype=torch.int64, device=self.device)
                res["records"] = rec
            nact = res.get("n_active")
            if nact is None or tuple(nact.shape) != (cnt, self.F):
                pad = torch.zeros((cnt * self.F + 3) // 4 * 4, dtype=torch.int32, device=self.device)
                nact = pad[: cnt * self.F].view(cnt, self.F)       # storage padded to 16 bytes
                res["n_active"], res["n_active16"] = nact, pad
            gid = np.ascontiguousarray(self.geom_id[a:b])
            coff = np.ascontiguousarray(self.cell_off[a:b + 1])
            scr_off = int(self.lib.cab200_sim_scratch_bytes(self.P, self.F)) + int(
                self.lib.cab200_raster_scratch_bytes(self.P))
            scr_off = (scr_off + 15) // 16 * 16
            if scratch_half:
                scr_off += (int(self.lib.cab200_encode_scratch_bytes(self.P, len(self.dgeo))) + 15) // 16 * 16
            rc = self.lib.cab200_encode_batch(
                cnt, _lib.as_i32p(gid), len(self.dgeo), _lib.as_i32p(self.g_ncell), _lib.as_i64p(coff),
                limg.data_ptr(), lmsk.data_ptr(), h, w, self.F, self.frames.data_ptr(),
                float(threshold), 1 if quirk else 0, 1 if image else 0, 1 if mask else 0,
                tabs.data_ptr(), gtab.data_ptr(), _lib.as_i64p(gtab_off),
                _lib.as_i32p(gmax_ev) if parse == "events" else None,
                arena.data_ptr(), arena.numel(), cursor.data_ptr(),
                rec.data_ptr(), nact.data_ptr(), int(stage_bytes),
                self._scratch.data_ptr() + scr_off, self.lib.cab200_encode_scratch_bytes(cnt, len(self.dgeo)),
                _stream_ptr(self.device))
            _lib.check(rc, "cab200_encode_batch")
        return res

    @staticmethod
    def records_to_host(records: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``records`` (int64 [..., 2]) -> (offset, size, flags) arrays."""
        off = records[..., 0]
        size = records[..., 1] & 0xFFFFFFFF
        flags = (records[..., 1] >> 32) & 0xFFFFFFFF
        return off, size, flags

    def encoded_files(self, enc: Dict[str, torch.Tensor]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Host copies ``(arena[:used], offset, size, flags)`` of an :meth:`encode` result (synchronises).
        The file of pouch i, frame f, kind k (0 image, 1 mask) is ``arena[offset[i,f,k]:][:size[i,f,k]]``."""
        used = int(enc["cursor"].item())
        arena = enc["arena"][: min(used, enc["arena"].numel())].cpu().numpy()
        off, size, flags = self.records_to_host(enc["records"].cpu().numpy())
        if np.any(flags & _lib.ENC_ARENA_FULL):
            raise _lib.Cab200Error(f"encode arena too small ({enc['arena'].numel()} bytes for {used} needed)")
        return arena, off, size, flags

    def run_files_to_host(self, chunk_pouches: int = 37, image: bool = True, mask: bool = True,
                           consumer=None, threshold: float = 0.1, quirk: bool = True,
                           arena_bytes_per_pixel: float = 0.25, wait: bool = True,
                           stage_bytes: int = 0, frames: str = "none") -> Dict[str, int]:
        """The whole hot path with HOST buffers on both sides, in the form the reference's batch path
        stores its results: pinned host inputs -> device, K1, then K3 chunk by chunk -- every sampled
        frame as a finished ``.png`` and ``.npz`` byte string -- copied into pinned host memory together
        with the float64 frames and the per-frame records.  Four ring slots: while K3 encodes chunk k,
        the host learns the size of chunk k-1 (a 16-byte read) and queues its copy-out on a second
        stream.  ``consumer(chunk: EncodedChunk)`` is called once a chunk has landed; if it returns an object
        with a ``result()`` method (a future of a writer thread that still reads the chunk), the chunk's host
        buffer is not reused before that has returned.  A chunk that outgrows its arena slot (about 0.16 bytes
        per pixel are typical, 0.25 reserved) is encoded again into an arena of the size it asked for.  With
        ``wait=False`` the last chunks are left in flight (their copy-out then overlaps whatever the
        caller queues next, e.g. the next ensemble's integration); :meth:`flush_files` completes them.
        ``frames``: which part of the sampled float64 state also goes to the host (``_fring["h_frames"]``):
        ``"none"`` (default: the files are the product; the batch path fetches the calcium state of a chunk
        itself when it generates prompts), ``"ca"`` -- the calcium state ``[P, F, n]`` (ensembles of one
        pouch size, otherwise like ``"all"``), ``"all"`` -- c, p, s, r as laid out on the device.
        With :meth:`set_relay` (multi-GPU, one process per GPU) chunks of a rank whose host path is slow
        travel over NVLink through a partner GPU.  Returns the byte counts moved (with ``wait=False``: of
        the copies queued during this call)."""
        dev = self.device
        w, h = self.output_size
        cur = torch.cuda.current_stream(dev)
        cp = max(1, min(chunk_pouches, self.P))
        cap = self.files_chunk_capacity(cp, arena_bytes_per_pixel)
        uniform = len(set(int(x) for x in self.g_ncell[self.geom_id])) == 1
        if frames == "ca" and not uniform:
            frames = "all"
        if frames not in ("ca", "all", "none"):
            raise ValueError(f"frames must be 'ca', 'all' or 'none', got {frames!r}")
        key = (cp, cap, image, mask, frames)
        st = getattr(self, "_fring", None)
        if st is None or st["key"] != key:
            self.flush_files()
            self.release_files_ring()
            # pinned rings are expensive to allocate (~0.3 s per GB) and identical from one ensemble of a batch to
            # the next: a process-wide cache keyed by device and shape hands them on (release_files_ring)
            pool_key = (dev.index, cp, cap, self.F)
            with _RING_LOCK:
                slots = _RING_CACHE.pop(pool_key, None)
            if slots is None:
                slots = []
                for _ in range(4):
                    slots.append({
                        "dev": {}, "h_arena": torch.empty(cap, dtype=torch.uint8, pin_memory=True),
                        "h_rec": torch.empty((cp, self.F, 2, 2), dtype=torch.int64, pin_memory=True),
                        "h_cur": torch.empty(2, dtype=torch.int64, pin_memory=True),
                        "h_nact": torch.empty((cp * self.F + 3) // 4 * 4, dtype=torch.int32, pin_memory=True),
                        "meta": torch.cuda.Event(), "data": torch.cuda.Event(), "busy": None})
            st = {"key": key, "pool_key": pool_key, "cs": torch.cuda.Stream(device=dev), "fs": torch.cuda.Stream(device=dev),
                  # K3 chunks alternate between two streams: the next chunk's CTAs fill the SMs the previous
                  # chunk's last wave leaves idle (one stream: 4 x 3.33 ms per 148 pouches, one launch: 12.9 ms)
                  "ks": [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)], "k3_done": [None, None],
                  # the few bytes the host waits for (sizes, records) are published from a high-priority stream:
                  # behind the other K3 stream's thousands of pending CTAs they would arrive milliseconds late
                  "ps": torch.cuda.Stream(device=dev, priority=-1),
                  "slots": slots, "launched": [], "copying": [],
                  "frames_done": torch.cuda.Event(),
                  "h_frames": (torch.empty(max(self.frame_elems, 1), dtype=torch.float64, pin_memory=True)
                               if frames == "all" else
                               torch.empty((self.P, self.F, int(self.g_ncell[0])), dtype=torch.float64, pin_memory=True)
                               if frames == "ca" else None)}
            self._fring = st
        cs = st["cs"]
        d2h = 0
        launched: List[Dict[str, Any]] = st["launched"]      # K3 queued, size not yet known to the host
        copying: List[Dict[str, Any]] = st["copying"]        # copy-out queued

        def queue_copy(e):
            nonlocal d2h
            slot = e["slot"]
            slot["meta"].synchronize()
            w_ = slot.pop("writer", None)
            if w_ is not None:
                w_.result()                      # a writer thread still read this slot's previous chunk
            used = int(slot["h_cur"][0])
            if used > cap:
                # the chunk outgrew its arena slot: encode it again into an arena of the size it asked for
                # (synchronous and rare; the files of a flat-shaded frame take ~0.16 bytes per pixel)
                consumer_, threshold_, quirk_ = e["consumer"]
                for ks_ in st["ks"]:
                    ks_.synchronize()            # the launch scratch is shared with the chunks in flight
                big = self.encode(image=image, mask=mask, pouches=(e["lo"], e["hi"]), threshold=threshold_, quirk=quirk_,
                                  out={}, arena_bytes=used + (1 << 20), stage_bytes=stage_bytes)
                torch.cuda.current_stream(dev).synchronize()
                used = int(big["cursor"].item())
                n_ = e["hi"] - e["lo"]
                slot["h_rec"][:n_].copy_(big["records"])
                slot["h_nact"][: n_ * self.F].copy_(big["n_active"].reshape(-1))
                e["big_arena"] = big["arena"][:used].cpu()
                e["used"] = used
                d2h += used
                copying.append(e)
                return
            e["used"] = used
            relay = getattr(self, "_relay", None)
            consumer_ = e["consumer"][0]
            if relay is not None and relay.active and relay.role == "sender" and consumer_ is None and used:
                # slow host path: the chunk goes over NVLink into the partner GPU's staging ring, which
                # forwards it to pinned host memory (calcium_b200/relay.py); acknowledged in finish()
                e["ticket"] = relay.send(slot["dev"]["arena"].data_ptr(), used, cs)
                slot["data"].record(cs)
            else:
                with torch.cuda.stream(cs):
                    if used:
                        slot["h_arena"][:used].copy_(slot["dev"]["arena"][:used], non_blocking=True)
                    slot["data"].record(cs)
            d2h += used
            copying.append(e)

        def finish(e):
            slot = e["slot"]
            slot["data"].synchronize()
            if "ticket" in e:
                self._relay.wait_ack(e.pop("ticket"))      # the partner has landed the chunk in host memory
            consumer, threshold, quirk = e["consumer"]
            if consumer is not None:
                n = e["hi"] - e["lo"]
                arena_host = e.pop("big_arena", None)
                arena_host = slot["h_arena"][: e["used"]] if arena_host is None else arena_host
                ret = consumer(EncodedChunk(self, e["lo"], e["hi"], arena_host.numpy(),
                                            slot["h_rec"][:n].numpy(),
                                            slot["h_nact"][: n * self.F].view(n, self.F).numpy(), threshold, quirk))
                if ret is not None and hasattr(ret, "result"):
                    slot["writer"] = ret         # the slot's host buffers stay untouched until it has returned
            slot["busy"] = None

        st["queue_copy"], st["finish"] = queue_copy, finish
        with torch.cuda.device(dev):
            cur.wait_event(st["frames_done"])        # a previous call's frame copy-out reads self.frames
            self.simulate(upload=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            fs = st["fs"]                            # own stream: waits for K1, and an arena copy of the
            fs.wait_event(ev)                        # previous call must not queue behind that wait
            with torch.cuda.stream(fs):
                if frames == "all":
                    st["h_frames"].copy_(self.frames, non_blocking=True)
                elif frames == "ca":
                    n0 = int(self.g_ncell[0])
                    ca = self.frames.view(self.P, self.F, 4, n0)[:, :, 0, :].contiguous()
                    ca.record_stream(fs)
                    st["h_frames"].copy_(ca, non_blocking=True)
                st["frames_done"].record(fs)
            if st["h_frames"] is not None:
                d2h += st["h_frames"].numel() * 8
            k = st.get("k", 0)
            for ks in st["ks"]:
                ks.wait_event(ev)                    # K1 (and whatever the caller queued before it)
            for lo in range(0, self.P, cp):
                hi = min(lo + cp, self.P)
                slot = st["slots"][k % 4]
                ks = st["ks"][k % 2]                  # slot k % 4 always meets stream k % 2
                if slot["busy"] is not None:             # its previous chunk must be out of the way
                    e = slot["busy"]
                    if e in launched:
                        launched.remove(e); queue_copy(e)
                    if e in copying:
                        copying.remove(e); finish(e)
                ks.wait_event(slot["data"])
                with torch.cuda.stream(ks):
                    enc = self.encode(image=image, mask=mask, pouches=(lo, hi), threshold=threshold, quirk=quirk,
                                      out=slot["dev"], arena_bytes=cap, stage_bytes=stage_bytes, scratch_half=k % 2)
                    n = hi - lo
                    k3_ev = torch.cuda.Event()
                    k3_ev.record(ks)
                st["k3_done"][k % 2] = k3_ev
                ps = st["ps"]
                ps.wait_event(k3_ev)
                with torch.cuda.stream(ps):
                    # the sizes reach the host by zero-copy stores, not through the copy engine (it is busy
                    # with megabytes of earlier chunks, and the host needs these bytes to queue the next copy)
                    for dst, src in ((slot["h_rec"], enc["records"]), (slot["h_nact"], enc["n_active16"]),
                                     (slot["h_cur"], enc["cursor16"])):
                        nbytes = src.numel() * src.element_size()
                        _lib.check(self.lib.cab200_publish_to_host(dst.data_ptr(), src.data_ptr(), nbytes // 16 * 16,
                                                                   _stream_ptr(dev)), "cab200_publish_to_host")
                    slot["meta"].record(ps)
                d2h += 8 + enc["records"].numel() * 8 + enc["n_active"].numel() * 4
                e = {"lo": lo, "hi": hi, "slot": slot, "consumer": (consumer, threshold, quirk)}
                slot["busy"] = e
                launched.append(e)
                # copy-outs are queued as soon as a chunk's size is known; the launch loop itself only blocks
                # when every ring slot is in flight (blocking earlier would hold back the chunk that is to
                # follow the pair running concurrently on the two K3 streams)
                while launched and (len(launched) >= 4 or launched[0]["slot"]["meta"].query()):
                    queue_copy(launched.pop(0))
                k += 1
            # the next integration overwrites the frames these chunks read: simulate() waits for them
            self._frames_readers = [done for done in st["k3_done"] if done is not None]
            st["k"] = k
            if wait:
                self.flush_files()
        return {"h2d_bytes": self.h2d_bytes, "d2h_bytes": d2h}

    def files_chunk_capacity(self, chunk_pouches: int = 37, arena_bytes_per_pixel: float = 0.25) -> int:
        """Arena bytes :meth:`run_files_to_host` reserves per chunk (the relay ring's slot size)."""
        w, h = self.output_size
        cp = max(1, min(chunk_pouches, self.P))
        return (int(cp * self.F * h * w * arena_bytes_per_pixel) + (1 << 20) + 15) // 16 * 16

    def set_relay(self, endpoint) -> None:
        """Attach a :class:`calcium_b200.relay.Endpoint` (or None) for the chunk copy-out."""
        self._relay = endpoint

    def release_files_ring(self) -> None:
        """Hands the pinned ring of :meth:`run_files_to_host` to the process-wide cache (call when done with the
        ensemble; the next ensemble of the same shape on this device takes it over instead of allocating)."""
        st = getattr(self, "_fring", None)
        if st is None:
            return
        self.flush_files()
        for slot in st["slots"]:
            w_ = slot.pop("writer", None)
            if w_ is not None:
                w_.result()
            slot["busy"] = None
        with _RING_LOCK:
            _RING_CACHE.setdefault(st["pool_key"], st["slots"])
        self._fring = None

    def flush_files(self) -> int:
        """Completes every chunk :meth:`run_files_to_host` left in flight (copy-out finished, consumer
        called); returns the bytes of the chunks whose size was only learnt here."""
        st = getattr(self, "_fring", None)
        if st is None or "queue_copy" not in st:
            return 0
        late = 0
        with torch.cuda.device(self.device):
            while st["launched"]:
                e = st["launched"].pop(0)
                st["queue_copy"](e)
                late += int(e["used"])
            while st["copying"]:
                st["finish"](st["copying"].pop(0))
            st["cs"].synchronize()
            st["fs"].synchronize()
        return late

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()

    # -- end to end: host in, host out -----------------------------------------------------------
    def run_to_host(self, chunk_pouches: int = 16, image: bool = True, instance: bool = True,
                    consumer=None, threshold: float = 0.1, quirk: bool = True) -> Dict[str, int]:
        """The whole hot path with HOST buffers on both sides, as a user of the reference gets it:
        pinned host inputs -> device, K1, then K2 chunk by chunk with every result (frames, images,
        instance masks, active counts) copied back into pinned host memory.  Device->host copies
        run on a second stream and overlap the rasterisation of the next chunk (two device and two
        host buffers).  ``consumer(lo, hi, arrays)`` is called with numpy views of the pinned
        buffers once a chunk has landed (e.g. to hand them to file writers); without a consumer
        the data is simply left in the pinned buffers.  Returns the byte counts moved."""
        dev = self.device
        w, h = self.output_size
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._h_frames = torch.empty(max(self.frame_elems, 1), dtype=torch.float64, pin_memory=True)
            self._ring: List[Dict[str, Any]] = []
        cs = self._copy_stream
        cp = max(1, min(chunk_pouches, self.P))
        if not self._ring or self._ring[0]["cp"] != cp or self._ring[0]["key"] != (image, instance):
            self._ring = []
            for _ in range(2):
                slot: Dict[str, Any] = {"cp": cp, "key": (image, instance), "dev": {}, "host": {},
                                        "done": torch.cuda.Event()}
                if image:
                    slot["host"]["image"] = torch.empty((cp, self.F, h, w, 2), dtype=torch.uint8, pin_memory=True)
                if instance:
                    slot["host"]["instance"] = torch.empty((cp, self.F, h, w), dtype=torch.int16, pin_memory=True)
                slot["host"]["n_active"] = torch.empty((cp, self.F), dtype=torch.int32, pin_memory=True)
                self._ring.append(slot)
        d2h = 0
        with torch.cuda.device(dev):
            self.simulate(upload=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            cs.wait_event(ev)
            with torch.cuda.stream(cs):
                self._h_frames.copy_(self.frames, non_blocking=True)
            d2h += self.frames.numel() * self.frames.element_size()
            k = 0
            landed = []
            for lo in range(0, self.P, cp):
                hi = min(lo + cp, self.P)
                slot = self._ring[k % 2]
                cur.wait_event(slot["done"])            # its previous copy-out has finished
                if consumer is not None and len(landed) >= 2:
                    plo, phi, pslot = landed.pop(0)
                    pslot["done"].synchronize()
                    consumer(plo, phi, {n_: t[: phi - plo].numpy() for n_, t in pslot["host"].items()})
                if hi - lo != cp:
                    slot["dev"] = {}                    # ragged tail: fresh, smaller device tensors
                out = self.rasterize(image=image, instance=instance, pouches=(lo, hi),
                                     threshold=threshold, quirk=quirk, out=slot["dev"])
                ev = torch.cuda.Event()
                ev.record(cur)
                cs.wait_event(ev)
                with torch.cuda.stream(cs):
                    for name, t in out.items():
                        slot["host"][name][: hi - lo].copy_(t, non_blocking=True)
                        d2h += t.numel() * t.element_size()
                    slot["done"].record(cs)
                landed.append((lo, hi, slot))
                k += 1
            cur.wait_stream(cs)
            if consumer is not None:
                for plo, phi, pslot in landed:
                    pslot["done"].synchronize()
                    consumer(plo, phi, {n_: t[: phi - plo].numpy() for n_, t in pslot["host"].items()})
        return {"h2d_bytes": self.h2d_bytes, "d2h_bytes": d2h}


_RING_CACHE: Dict[Any, List[Dict[str, Any]]] = {}
_RING_LOCK = threading.Lock()


class EncodedChunk:
    """The files of pouches ``[lo, hi)`` as K3 left them in (pinned) host memory.  Valid until the
    ring slot is reused, i.e. until the consumer callback returns."""

    def __init__(self, ens: "Ensemble", lo: int, hi: int, arena: np.ndarray, records: np.ndarray,
                 n_active: np.ndarray, threshold: float, quirk: bool):
        self.ens, self.lo, self.hi = ens, lo, hi
        self.arena = arena
        self.offset, self.size, self.flags = Ensemble.records_to_host(records)
        self.n_active = n_active
        self._threshold, self._quirk = threshold, quirk
        self._raw: Dict[int, Dict[str, np.ndarray]] = {}

    def file(self, i: int, f: int, kind: int):
        """Bytes of the ``.png`` (kind 0) / ``.npz`` (kind 1) of chunk-local pouch i, frame f; ``None``
        when there is none (not requested, or a mask of a frame without active cells)."""
        fl = int(self.flags[i, f, kind])
        if fl & _lib.ENC_OK:
            o = int(self.offset[i, f, kind])
            return self.arena[o:o + int(self.size[i, f, kind])]
        if fl & _lib.ENC_SKIPPED:
            return None
        if fl & _lib.ENC_STAGE_FULL:
            return self._host_encoded(i, f, kind)
        raise _lib.Cab200Error(f"frame ({self.lo + i}, {f}) kind {kind}: encode flags {fl}")

    def write_files(self, image_paths, mask_paths, threads: int = 16) -> Tuple[List[Tuple[int, int]], int]:
        """Writes the chunk's files with the native writer pool (``cab200_write_files``), straight from the arena.
        ``image_paths[i][f]`` / ``mask_paths[i][f]`` (``None`` entries and a ``None`` list are skipped) name the
        files of chunk-local pouch i, frame f.  Returns the ``(i, f)`` of the masks written and the bytes written."""
        lib = _lib.load()
        paths: List[bytes] = []
        offs: List[int] = []
        sizes: List[int] = []
        late: List[Tuple[str, bytes]] = []
        masks_written: List[Tuple[int, int]] = []
        n, F = self.flags.shape[0], self.flags.shape[1]
        for kind, table in ((0, image_paths), (1, mask_paths)):
            if table is None:
                continue
            fl = self.flags[:, :, kind]
            for i in range(n):
                row = table[i]
                for f in range(F):
                    flag = int(fl[i, f])
                    if row[f] is None or flag & _lib.ENC_SKIPPED:
                        continue
                    if flag & _lib.ENC_OK:
                        paths.append(row[f].encode())
                        offs.append(int(self.offset[i, f, kind])); sizes.append(int(self.size[i, f, kind]))
                    else:
                        late.append((row[f], bytes(self.file(i, f, kind))))     # staging overflow: host-encoded
                    if kind == 1:
                        masks_written.append((i, f))
        blob = b"\0".join(paths) + b"\0"
        poff = np.zeros(len(paths), dtype=np.int64)
        if paths:
            np.cumsum([len(p) + 1 for p in paths[:-1]], out=poff[1:])
        o = np.asarray(offs, dtype=np.int64)
        z = np.asarray(sizes, dtype=np.int64)
        written = np.zeros(1, dtype=np.int64)
        arena = np.ascontiguousarray(self.arena)
        _lib.check(lib.cab200_write_files(len(paths), blob, _lib.as_i64p(poff), arena.ctypes.data, _lib.as_i64p(o),
                                          _lib.as_i64p(z), int(threads), _lib.as_i64p(written)), "cab200_write_files")
        for path, data in late:
            with open(path, "wb") as fh:
                fh.write(data)
        return masks_written, int(written[0]) + sum(len(d) for _, d in late)

    def _host_encoded(self, i: int, f: int, kind: int) -> bytes:
        """A file that did not fit K3's staging buffer: rasterise the pouch with K2 and encode on the host."""
        import io as _io
        from . import io as cio
        raw = self._raw.get(i)
        if raw is None:
            out = self.ens.rasterize(image=True, instance=True, pouches=(self.lo + i, self.lo + i + 1),
                                     threshold=self._threshold, quirk=self._quirk)
            raw = {"image": out["image"][0].cpu().numpy(), "instance": _u16_numpy(out["instance"][0])}
            self._raw[i] = raw
        if kind == 0:
            img = raw["image"][f]
            return cio.encode_gray_alpha_png16(img[..., 0].astype(np.uint16) << 2, img[..., 1].astype(np.uint16) * 257)
        buf = _io.BytesIO()
        np.savez_compressed(buf, mask=raw["instance"][f])
        return buf.getvalue()


def make_specs(n: int, size: str = "medium", sim_types: Optional[Sequence[str]] = None,
               first_seed: int = 0, geometry_dir: Optional[str] = None,
               randomize: bool = True) -> List[PouchSpec]:
    """The deterministic ensemble of SURVEY.md section 8(d): pouch i uses seed = sim_number =
    first_seed + i, sim type round-robin over ``sim_types`` (default: all four, in the order of
    reference ``core/parameters.py:69-90``), parameters from ``generate_random_params(seed)``."""
    from .parameters import SimulationParameters
    types = list(sim_types) if sim_types else SimulationParameters.get_available_sim_types()
    specs = []
    for i in range(n):
        seed = first_seed + i
        sp = SimulationParameters(types[i % len(types)])
        prm = sp.generate_random_params(seed=seed) if randomize else sp.get_params_dict()
        specs.append(PouchSpec.draw(prm, size=size, sim_number=seed, geometry_dir=geometry_dir))
    return specs
