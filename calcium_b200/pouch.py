"""``Pouch`` -- drop-in for the reference's simulation class, computed on the GPU.

Same constructor, attributes, methods, return types and exceptions as the reference's
``core/pouch.py:107-611``; the arithmetic runs in ``libcalcium_b200`` (K1 integrator, K2 frame
rasteriser, K0 cells_mask) through :class:`calcium_b200.ensemble.Ensemble`.

What differs, deliberately:
  * ``disc_dynamics`` is not a dense ``(n, 4, T)`` array (864 MB for a medium pouch, reference
    ``core/pouch.py:204``) but a :class:`FrameStore`: it answers the reference's indexing
    ``disc_dynamics[cell, state, t]`` / ``[:, state, t]`` from frames kept on the GPU.  Time steps
    are captured on demand: the default set is the reference's batch sampling
    (``range(0, 18000, 200)``, ``main.py:167``) plus the last step; asking for any other step
    re-runs the (deterministic, ~30 ms) integration with that step added, so every ``t`` works.
  * ``_unpack_parameters`` also sets ``D_p`` (the reference forgets it and cannot construct a
    ``Pouch`` as shipped, SURVEY.md defect D1).
  * geometry comes from ``calcium_b200.geometry`` (cached; synthetic when the MSELab files are absent).
"""
from __future__ import annotations

import os
import warnings
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ensemble import (DEFAULT_FRAME_STEPS, DT, DURATION, REQUIRED_PARAMS, Ensemble, PouchSpec,
                       build_label_maps, draw_initial_state, draw_vplc, _u16_numpy)
from .geometry import GeometryLoader


class FrameStore:
    """Array-like view of a pouch's trajectory, shape ``(n_cells, 4, T)`` like the reference's
    ``disc_dynamics`` (states: 0 Ca2+, 1 IP3, 2 ER Ca2+, 3 IP3R inactivation).  Values live on the
    GPU as captured frames; indexing returns numpy (host copies)."""

    def __init__(self, pouch: "Pouch"):
        self._pouch = pouch
        self._host: Dict[int, np.ndarray] = {}      # t -> (4, n) host copy

    @property
    def shape(self) -> Tuple[int, int, int]:
        return (self._pouch.n_cells, 4, self._pouch.T)

    ndim = 3
    dtype = np.dtype(np.float64)

    def _frame(self, t: int) -> np.ndarray:
        T = self._pouch.T
        if t < 0:
            t += T
        if not 0 <= t < T:
            raise IndexError(f"index {t} is out of bounds for axis 2 with size {T}")
        f = self._host.get(t)
        if f is None:
            f = self._pouch._fetch_frame(t)
            self._host[t] = f
        return f

    def _invalidate(self) -> None:
        self._host.clear()

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (3 - len(key))
        if len(key) != 3:
            raise IndexError("too many indices for disc_dynamics")
        cells, state, tk = key
        if isinstance(tk, (int, np.integer)):
            return self._frame(int(tk)).T[cells, state]          # (n,4)[cells, state]
        if isinstance(tk, slice):
            ts = list(range(*tk.indices(self._pouch.T)))
        else:
            ts = [int(t) for t in np.asarray(tk).ravel()]
        T = self._pouch.T
        ts = [t + T if t < 0 else t for t in ts]
        for t in ts:
            if not 0 <= t < T:
                raise IndexError(f"index {t} is out of bounds for axis 2 with size {T}")
        self._pouch._ensure_steps(ts)                             # may re-run the capture and drop the host copies
        missing = [t for t in dict.fromkeys(ts) if t not in self._host]
        if missing:                                               # one capture pass + ONE device->host copy
            for t, f in zip(missing, self._pouch._fetch_frames(missing)):
                self._host[t] = f
        block = np.stack([self._host[t] for t in ts], axis=-1)    # (4, n, len)
        return np.transpose(block, (1, 0, 2))[cells, state]

    def materialize(self) -> np.ndarray:
        """The full dense ``(n, 4, T)`` array (every step captured) -- the reference's layout."""
        return self[:, :, :]

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype)


class Pouch:
    """Simulates calcium ion dynamics in epithelial tissue (GPU drop-in for the reference class)."""

    DEFAULT_TIME_STEP = DT          # seconds, reference core/pouch.py:127
    DEFAULT_DURATION = DURATION     # seconds, reference core/pouch.py:128
    REQUIRED_PARAMS = list(REQUIRED_PARAMS)

    def __init__(self, params: Optional[Dict[str, float]] = None, size: str = "medium",
                 sim_number: int = 0, save: bool = False, save_name: str = "default",
                 geometry_dir: Optional[str] = None, output_size: Tuple[int, int] = (512, 512),
                 jpeg_quality: int = 90, *, frame_steps: Optional[Sequence[int]] = None,
                 device: Optional[Any] = None):
        self.size = size
        self.save_name = save_name
        self.sim_number = sim_number
        self.save = save
        self.output_size = output_size
        self.jpeg_quality = jpeg_quality

        self.param_dict = params if params is not None else self._get_default_params()
        self._validate_parameters()

        self._geometry = GeometryLoader(geometry_dir).geometry(size)
        self.new_vertices = self._geometry.vertices
        self.n_cells = self._geometry.n_cells
        self.dt = self.DEFAULT_TIME_STEP
        self.T = int(self.DEFAULT_DURATION / self.dt)

        self._unpack_parameters()
        self.D_c = self.D_c_ratio * self.D_p
        self.D_p = self.D_p

        # per-cell PLC activation strength, same draw as reference core/pouch.py:199-200
        self.V_PLC = draw_vplc(self.param_dict, self.n_cells, sim_number)

        self._device = torch.device(device) if device is not None else None
        steps = DEFAULT_FRAME_STEPS + [self.T - 1] if frame_steps is None else list(frame_steps)
        self._steps = sorted({int(t) for t in steps if 0 <= int(t) < self.T})
        self._ens: Optional[Ensemble] = None
        self._r0: Optional[np.ndarray] = None
        self._simulated = False
        self.disc_dynamics = FrameStore(self)
        self._cells_mask: Optional[np.ndarray] = None

    # ---- reference attributes that are expensive: materialised on first use ---------------------
    @property
    def adj_matrix(self) -> np.ndarray:
        return self._geometry.adjacency

    @property
    def laplacian_matrix(self) -> np.ndarray:
        return self._geometry.laplacian

    @property
    def laplacian_sparse(self):
        from scipy.sparse import csr_matrix
        rowptr, col, val = self._geometry.csr
        return csr_matrix((val, col, rowptr), shape=(self.n_cells, self.n_cells))

    @property
    def cells_mask(self) -> np.ndarray:
        """(H, W) int32 pixel -> cell id map (0 = background), reference ``core/pouch.py:260-287``."""
        if self._cells_mask is None:
            _, lmask = build_label_maps(self._geometry, self.output_size, self._dev())
            self._cells_mask = _u16_numpy(lmask).astype(np.int32)
        return self._cells_mask

    def _create_cells_mask(self) -> np.ndarray:
        return self.cells_mask.copy()

    # ---- parameters (reference core/pouch.py:209-258) ----------------------------------------------
    def _get_default_params(self) -> Dict[str, float]:
        return {"K_PLC": 0.2, "K_5": 0.66, "k_1": 1.11, "k_a": 0.08, "k_p": 0.13, "k_2": 0.0203,
                "V_SERCA": 0.9, "K_SERCA": 0.1, "c_tot": 2.0, "beta": 0.185, "k_i": 0.4,
                "D_p": 0.005, "tau_max": 800.0, "k_tau": 1.5, "lower": 0.5, "upper": 0.7,
                "frac": 0.007680491551459293, "D_c_ratio": 0.1}

    def _validate_parameters(self) -> None:
        provided, required = set(self.param_dict.keys()), set(self.REQUIRED_PARAMS)
        missing, extra = required - provided, provided - required
        if missing or extra:
            parts = ["Parameter validation failed:"]
            if missing:
                parts.append(f"  Missing: {', '.join(sorted(missing))}")
            if extra:
                parts.append(f"  Unexpected: {', '.join(sorted(extra))}")
            raise ValueError("\n".join(parts))

    def _unpack_parameters(self) -> None:
        for key in self.REQUIRED_PARAMS:
            setattr(self, key, self.param_dict[key])

    # ---- simulation ----------------------------------------------------------------------------
    def _dev(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("calcium_b200 needs a CUDA device (no CPU fallback)")
            self._device = torch.device("cuda", torch.cuda.current_device())
        elif self._device.type == "cuda" and self._device.index is None:
            if not torch.cuda.is_available():
                raise RuntimeError("calcium_b200 needs a CUDA device (no CPU fallback)")
            self._device = torch.device("cuda", torch.cuda.current_device())     # "cuda" = the current device
        return self._device

    def _run(self) -> None:
        spec = PouchSpec(params=self.param_dict, geometry=self._geometry, sim_number=self.sim_number,
                         vplc=self.V_PLC.flatten(), r0=self._r0)
        self._ens = Ensemble([spec], output_size=self.output_size, T=self.T,
                             frame_steps=self._steps, device=self._dev(), dt=self.dt)
        self._ens.simulate()
        self._ens.check_status()
        self.disc_dynamics._invalidate()

    def simulate(self) -> bool:
        """Run the simulation (reference ``core/pouch.py:289-372``).  ``ValueError`` for invalid
        parameters; other failures are warned about and re-raised, like the reference."""
        if self.c_tot <= 0 or self.beta <= 0 or self.D_p <= 0:
            raise ValueError("Invalid simulation parameters: c_tot, beta, and D_p must be positive")
        if not 0 <= self.frac <= 1:
            raise ValueError(f"Invalid fraction of initiator cells: {self.frac}. Must be between 0 and 1")
        try:
            self._r0, _ = draw_initial_state(self.param_dict, self.n_cells, self.sim_number, self.V_PLC)
            self._run()
            self._simulated = True
            return True
        except Exception as e:  # noqa: BLE001 - mirrors the reference's warn-and-reraise
            warnings.warn(f"Simulation error: {str(e)}")
            raise

    def _ensure_steps(self, steps: Sequence[int]) -> None:
        missing = [int(t) for t in steps if int(t) not in self._steps_set()]
        if missing and self._simulated:
            self._steps = sorted(set(self._steps) | set(missing))
            self._run()
        elif missing:
            self._steps = sorted(set(self._steps) | set(missing))

    def _steps_set(self):
        cached = getattr(self, "_steps_cache", None)
        if cached is None or cached[0] is not self._steps:
            cached = (self._steps, set(self._steps))
            self._steps_cache = cached
        return cached[1]

    def _fetch_frames(self, ts: Sequence[int]) -> np.ndarray:
        """(len(ts), 4, n) host copy of the states at steps ``ts``: the steps are captured in one pass and
        fetched with one device->host copy."""
        if not self._simulated:
            return np.zeros((len(ts), 4, self.n_cells))
        self._ensure_steps(ts)
        idx = torch.as_tensor([self._ens.frame_index(t) for t in ts], dtype=torch.long, device=self._ens.device)
        return self._ens.frames_of(0).index_select(0, idx).cpu().numpy()

    def _fetch_frame(self, t: int) -> np.ndarray:
        """(4, n) host copy of the state at step t."""
        if not self._simulated:
            # the reference's array is all zeros before simulate(); so is ours
            return np.zeros((4, self.n_cells))
        self._ensure_steps([t])
        fi = self._ens.frame_index(t)
        return self._ens.frames_of(0)[fi].cpu().numpy()

    # ---- images and masks ------------------------------------------------------------------------
    def _raster(self, time_step: int, threshold: float, image: bool, instance: bool, active: bool,
                quirk: bool = True, **edge) -> Dict[str, np.ndarray]:
        self._ensure_steps([time_step])
        ens = self._ens
        fi = ens.frame_index(time_step)
        # one frame: rasterise through a 1-frame view of the ensemble
        sub = _SingleFrame(ens, fi)
        return sub.rasterize(image=image, instance=instance, active=active, threshold=threshold,
                             quirk=quirk, **edge)

    def generate_image(self, time_step: int, output_path: Optional[str] = None,
                       with_border: bool = False, colormap: str = "gray", edge_blur: bool = False,
                       blur_kernel_size: int = 3, blur_type: str = "mean") -> np.ndarray:
        """(H, W, 2) uint8 gray + alpha image of step ``time_step`` (reference
        ``core/pouch.py:374-492``).  The default branch is K2 (``cab200_raster_batch``); the optional
        ``edge_blur`` / ``with_border`` branches (``:456-481``) are K2b (``cab200_raster_edges_batch``)
        over the static outline / blend-weight maps of K0c."""
        if time_step < 0 or time_step >= self.T:
            raise ValueError(f"Time step must be between 0 and {self.T-1}, got {time_step}")
        try:
            edge = dict(with_border=bool(with_border), edge_blur=bool(edge_blur),
                        blur_kernel_size=int(blur_kernel_size), blur_type=blur_type)
            if not self._simulated:
                img = self._blank_image(**edge)
            else:
                img = self._raster(time_step, 0.1, True, False, False, **edge)["image"]
            if output_path is not None:
                self._save_image(img, output_path)
            return img
        except Exception as e:  # noqa: BLE001 - same wrapping as the reference (:491-492)
            raise RuntimeError(f"Error generating image: {str(e)}")

    def _blank_image(self, with_border=False, edge_blur=False, blur_kernel_size=3, blur_type="mean") -> np.ndarray:
        """``generate_image`` before ``simulate()``: the reference's ``disc_dynamics`` is all zeros then, so
        every cell is painted gray 0 / alpha 255; the edge branches go through K2b on a zero frame."""
        dev = self._dev()
        if with_border or edge_blur:
            return _raster_one(self._geometry, self.output_size, dev, torch.zeros((4, self.n_cells), dtype=torch.float64,
                                                                                  device=dev),
                               0.1, True, True, False, False, with_border, edge_blur, blur_kernel_size,
                               blur_type)["image"]
        limg, _ = build_label_maps(self._geometry, self.output_size, dev)
        alpha = (_u16_numpy(limg) > 0).astype(np.uint8) * 255
        return np.stack([np.zeros_like(alpha), alpha], axis=-1)

    def _create_blur_kernel(self, size: int, blur_type: str) -> np.ndarray:
        if blur_type == "motion":
            k = np.zeros((size, size), np.float32)
            k[size // 2, :] = 1.0 / size
            return k
        return np.ones((size, size), np.float32) / (size * size)

    def _save_image(self, img: np.ndarray, output_path: str) -> None:
        from .io import save_gray_alpha_png
        save_gray_alpha_png(img, output_path, bit_depth=10)

    def get_cell_masks(self, active_only: bool = False, time_step: Optional[int] = None,
                       threshold: float = 0.1) -> np.ndarray:
        """Cell id mask (reference ``core/pouch.py:515-548``)."""
        if active_only:
            if time_step is None:
                raise ValueError("time_step must be provided when active_only=True")
            if not self._simulated:
                return np.zeros_like(self.cells_mask)
            return self._raster(time_step, threshold, False, False, True)["active"]
        return self.cells_mask.copy()

    def get_active_cells(self, time_step: int, threshold: float = 0.1) -> List[int]:
        """0-based ids of cells with Ca2+ above ``threshold`` (reference ``core/pouch.py:550-565``)."""
        calcium = self.disc_dynamics[:, 0, time_step]
        return np.where(calcium > threshold)[0].tolist()

    def get_instance_mask(self, time_step: int, threshold: float = 0.1, quirk: bool = True) -> np.ndarray:
        """``create_instance_mask_from_cells(get_cell_masks(active_only=True, t), get_active_cells(t))``
        (reference ``utils/sam2_data_generator.py:322-325``) in one GPU pass; uint16."""
        return self._raster(time_step, threshold, False, True, False, quirk=quirk)["instance"]

    def generate_label_data(self, time_step: int, threshold: float = 0.1) -> Dict[str, Any]:
        active = self.get_active_cells(time_step, threshold)
        ca = self.disc_dynamics[:, 0, time_step]
        return {
            "active_cells": active,
            "calcium_values": {cid: float(ca[cid]) for cid in active},
            "n_active": len(active),
            "simulation_info": {"sim_number": self.sim_number, "size": self.size,
                                "time_step": time_step, "total_cells": self.n_cells},
        }

    def __repr__(self) -> str:
        return (f"Pouch(size='{self.size}', n_cells={self.n_cells}, "
                f"sim_number={self.sim_number}, T={self.T})")

    def __str__(self) -> str:
        return (f"Calcium Signaling Simulation\n"
                f"  Geometry: {self.size} ({self.n_cells} cells)\n"
                f"  Duration: {self.T * self.dt:.0f}s ({self.T} steps @ {self.dt}s)\n"
                f"  Output: {self.output_size[0]}×{self.output_size[1]} pixels\n"
                f"  Sim ID: {self.sim_number}")


def _raster_one(geometry, output_size, dev, frame: torch.Tensor, threshold: float, quirk: bool,
                image: bool, instance: bool, active: bool, with_border: bool = False, edge_blur: bool = False,
                blur_kernel_size: int = 3, blur_type: str = "mean") -> Dict[str, np.ndarray]:
    """K2 (and K2b for the edge branches) over ONE frame ``frame`` = contiguous ``[4][n]`` float64 on ``dev``."""
    from .ensemble import _stream_ptr, blur_code, build_edge_maps, edge_mode
    lib = _lib.load()
    w, h = output_size
    limg, lmsk = build_label_maps(geometry, output_size, dev)
    n = geometry.n_cells
    out: Dict[str, np.ndarray] = {}
    mode = edge_mode(with_border, edge_blur) if image else 0
    with torch.cuda.device(dev):
        img = torch.empty((h, w, 2), dtype=torch.uint8, device=dev) if image else None
        ins = torch.empty((h, w), dtype=torch.int16, device=dev) if instance else None
        act = torch.empty((h, w), dtype=torch.int32, device=dev) if active else None
        nact = torch.empty(1, dtype=torch.int32, device=dev)
        scr_bytes = int(lib.cab200_raster_scratch_bytes(1))
        scratch = torch.empty(scr_bytes + 16, dtype=torch.uint8, device=dev)
        gid = np.zeros(1, dtype=np.int32)
        ncell = np.array([n], dtype=np.int32)
        coff = np.array([0, n], dtype=np.int64)
        rc = lib.cab200_raster_batch(
            1, _lib.as_i32p(gid), _lib.as_i32p(ncell), _lib.as_i64p(coff),
            limg.data_ptr(), lmsk.data_ptr(), h, w, 1, frame.data_ptr(), float(threshold),
            1 if quirk else 0,
            img.data_ptr() if (img is not None and not mode) else None,
            act.data_ptr() if act is not None else None,
            ins.data_ptr() if ins is not None else None,
            nact.data_ptr(), scratch.data_ptr(), scr_bytes, _stream_ptr(dev))
        _lib.check(rc, "cab200_raster_batch")
        if mode:
            edges, dilated, alpha = build_edge_maps(geometry, output_size, dev, blur_kernel_size, blur_type)
            rc = lib.cab200_raster_edges_batch(
                1, _lib.as_i32p(gid), _lib.as_i32p(ncell), _lib.as_i64p(coff),
                limg.data_ptr(), edges.data_ptr(), dilated.data_ptr(), alpha.data_ptr(), h, w, 1, frame.data_ptr(),
                mode, int(blur_kernel_size), blur_code(blur_type), img.data_ptr(),
                scratch.data_ptr(), scr_bytes, _stream_ptr(dev))
            _lib.check(rc, "cab200_raster_edges_batch")
        if img is not None:
            out["image"] = img.cpu().numpy()
        if ins is not None:
            out["instance"] = ins.cpu().numpy().view(np.uint16)
        if act is not None:
            out["active"] = act.cpu().numpy()
        out["n_active"] = int(nact.cpu().item())
    return out


class _SingleFrame:
    """Rasterises one captured frame of a 1-pouch ensemble (K2 / K2b with n_frames = 1)."""

    def __init__(self, ens: Ensemble, fi: int):
        self.ens, self.fi = ens, fi

    def rasterize(self, image: bool, instance: bool, active: bool, threshold: float, quirk: bool,
                  **edge) -> Dict[str, np.ndarray]:
        ens = self.ens
        frame = ens.frames_of(0)[self.fi].contiguous()      # a contiguous [4][n] block for this frame
        return _raster_one(ens.geometries[0], ens.output_size, ens.device, frame, threshold, quirk,
                           image, instance, active, **edge)
