"""The reference's defect pipeline.

``apply_all_defects(image, cells_mask, config)`` (``utils/image_processing.py:130-175``) chains background,
optical, sensor and post-processing defects.  Two kinds:

* **deterministic** defects that are a function of the pixel value and ONE statistic of the frame, or a product
  with a static map -- uniform background fluorescence (``artifacts/background.py:46-56``), chromatic aberration
  (a copy below three channels, ``artifacts/optical.py:22-24``), vignetting (``:49-83``), dynamic range
  compression (``artifacts/sensor.py:117-135``), quantization (``:138-156``), brightness / contrast
  (``utils/image_processing.py:97-127``).  Between two defects the reference's image is uint8 again, so each is
  a 256-entry table row chosen by the frame's statistic (``np.max(image)`` or ``image.max() > 1.0``) or a
  float64 multiply.  The rows are built HERE, on the host, by evaluating the reference's numpy expressions for
  every possible (statistic, pixel value) pair -- same numpy, same rounding -- and **K5**
  (``cab200_defect_pass``) applies them on the GPU to whole chunks of frames.

* **stochastic / library** defects -- the non-uniform background (a ``np.random.rand`` field through a 99x99
  ``cv2.GaussianBlur``, ``artifacts/background.py:24-45``), Poisson / readout / Gaussian noise
  (``artifacts/sensor.py:8-104``), defocus and partial defocus (``utils/image_processing.py:34-93``).  Their
  cost and their *identity* is the legacy Mersenne-Twister stream (``np.random.poisson`` consumes a
  data-dependent number of draws per pixel -- inherently sequential) and OpenCV's own Gaussian filter, so they
  run on the host, frame by frame, with the pouch's generator threaded through exactly like the prompt
  draws (:func:`apply_all_defects_host`): same calls, same order, same state afterwards as the reference.
  ``spontaneous_luminescence`` and ``cell_fragments`` index channel 2 of the two-channel frame and raise in the
  reference itself -- here too, after consuming the same draws, so the frame is skipped like the
  reference's loop skips it (``main.py:342-351``).

``random_defect_config`` is ``generate_random_defect_config`` (``main.py:29-92``): the per-frame configuration
the reference draws from the stdlib generator when the caller passes no ``defect_configs``.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

STOCHASTIC = ("poisson_noise", "readout_noise", "gaussian_noise", "partial_defocus")
BROKEN_IN_REFERENCE = ("spontaneous_luminescence", "cell_fragments")


def host_defects(config: Optional[Dict[str, Any]]) -> List[str]:
    """Names of the defects ``config`` switches on that K5 cannot apply (they draw from the generator, call
    OpenCV's Gaussian filter, or raise in the reference): a frame with any of them takes
    :func:`apply_all_defects_host`."""
    if not config:
        return []
    bad = [k for k in STOCHASTIC + BROKEN_IN_REFERENCE if config.get(k, False) and (k != "partial_defocus" or config.get("defocus_blur", False))]
    if config.get("background_fluorescence", False) and config.get("non_uniform_background", True):
        bad.append("non_uniform_background")
    if config.get("defocus_blur", False):
        bad.append("defocus_blur")
    return bad


unsupported_defects = host_defects      # round-1 name


def _values() -> Tuple[np.ndarray, np.ndarray]:
    m = np.arange(256, dtype=np.uint8).reshape(256, 1)          # the frame statistic np.max(image), a uint8
    v = np.arange(256, dtype=np.uint8).reshape(1, 256)          # the pixel value
    return m, v


def table_background(intensity: float) -> np.ndarray:
    """``np.clip(result[:, :, i] + intensity * max_val, 0, max_val)`` assigned into the uint8 image
    (``artifacts/background.py:46-56``), for every (max_val, value)."""
    m, v = _values()
    uniform = float(intensity) * m.astype(np.float64)             # python float * np.uint8 scalar -> float64
    out = np.clip(v + uniform, 0, m)
    return out.astype(np.uint8)


def table_gamma(gamma: float) -> np.ndarray:
    """``apply_dynamic_range_compression`` (``artifacts/sensor.py:117-135``); row 1: ``image.max() > 1.0``."""
    v = np.arange(256, dtype=np.uint8)
    with np.errstate(all="ignore"):
        row0 = np.power(v.copy(), gamma).astype(np.uint8)                      # already "normalised": no scaling
        row1 = (np.power(v / 255.0, gamma) * 255.0).astype(np.uint8)
    return np.stack([row0, row1])


def table_quantization(bits: int) -> np.ndarray:
    """``apply_quantization`` (``artifacts/sensor.py:138-156``) for every (max_val, value)."""
    m, v = _values()
    levels = 2 ** int(bits)
    with np.errstate(all="ignore"):
        step = m / levels                                                       # uint8 / int -> float64
        quantized = np.round(v / step) * step
        out = np.clip(quantized, 0, m)
        out = np.where(np.isfinite(out), out, 0.0)                              # max_val == 0: 0/0 (never a rendered frame)
    return out.astype(np.uint8)


def table_brightness_contrast(brightness: float, contrast: float) -> np.ndarray:
    """``adjust_brightness_contrast`` (``utils/image_processing.py:97-127``), float32 like the reference; row 1:
    ``image.max() > 1.0``."""
    rows = []
    for scaled in (False, True):
        img_float = np.arange(256, dtype=np.uint8).astype(np.float32)
        if scaled:
            img_float /= 255.0
        img_float = (img_float - 0.5) * contrast + 0.5
        img_float += brightness
        img_float = np.clip(img_float, 0, 1.0)
        if scaled:
            img_float *= 255.0
        with np.errstate(all="ignore"):
            rows.append(img_float.astype(np.uint8))
    return np.stack(rows)


def vignette_map(output_size: Tuple[int, int], strength: float) -> np.ndarray:
    """The float64 factor of ``add_vignetting`` (``artifacts/optical.py:62-72``), ``[H][W]``."""
    cols, rows = int(output_size[0]), int(output_size[1])
    x = np.linspace(-1, 1, cols)
    y = np.linspace(-1, 1, rows)
    xv, yv = np.meshgrid(x, y)
    r = np.sqrt(xv ** 2 + yv ** 2)
    return np.ascontiguousarray(1 - strength * np.clip(r, 0, 1) ** 2)


_PLAN_KEYS = ("background_fluorescence", "non_uniform_background", "background_intensity", "vignetting",
              "vignetting_strength", "dynamic_range_compression", "gamma", "quantization", "bit_depth",
              "adjust_brightness_contrast", "brightness", "contrast")
_PLAN_CACHE: Dict[Any, List[Tuple[int, Optional[np.ndarray]]]] = {}


def plan(config: Optional[Dict[str, Any]], output_size: Tuple[int, int]) -> List[Tuple[int, Optional[np.ndarray]]]:
    """The passes of ``apply_all_defects`` for the supported defects of ``config``, in the reference's order:
    ``(kind, table or float64 map)``.  Cached per (configuration, output size): a batch usually carries one
    configuration for all its pouches."""
    if not config:
        return []
    key = (tuple((k, repr(config.get(k))) for k in _PLAN_KEYS), int(output_size[0]), int(output_size[1]))
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        if len(_PLAN_CACHE) > 64:
            _PLAN_CACHE.clear()
        hit = _PLAN_CACHE[key] = _build_plan(config, output_size)
    return hit


def _build_plan(config: Dict[str, Any], output_size: Tuple[int, int]) -> List[Tuple[int, Optional[np.ndarray]]]:
    ops: List[Tuple[int, Optional[np.ndarray]]] = []
    if config.get("background_fluorescence", False) and not config.get("non_uniform_background", True):
        ops.append((_lib.DEFECT_LUT_MAX, table_background(config.get("background_intensity", 0.1))))
    if config.get("vignetting", False):
        ops.append((_lib.DEFECT_MUL, vignette_map(output_size, config.get("vignetting_strength", 0.5))))
    if config.get("dynamic_range_compression", False):
        ops.append((_lib.DEFECT_LUT_FLAG, table_gamma(config.get("gamma", 0.7))))
    if config.get("quantization", False):
        ops.append((_lib.DEFECT_LUT_MAX, table_quantization(config.get("bit_depth", 8))))
    if config.get("adjust_brightness_contrast", False):
        ops.append((_lib.DEFECT_LUT_FLAG, table_brightness_contrast(config.get("brightness", 0), config.get("contrast", 1.0))))
    return ops


def apply_defects(image: torch.Tensor, config: Optional[Dict[str, Any]], output_size: Tuple[int, int]) -> torch.Tensor:
    """``apply_all_defects`` (supported defects only) in place on ``image``: a CUDA uint8 tensor whose last three
    dimensions are ``(H, W, 2)``; every leading index is one frame."""
    ops = plan(config, output_size)
    if not ops:
        return image
    if not image.is_cuda or image.dtype != torch.uint8 or not image.is_contiguous() or image.shape[-1] != 2:
        raise ValueError("apply_defects expects a contiguous CUDA uint8 tensor [..., H, W, 2]")
    h, w = int(image.shape[-3]), int(image.shape[-2])
    if (w, h) != (int(output_size[0]), int(output_size[1])):
        raise ValueError("image shape does not match output_size")
    n_frames = image.numel() // (h * w * 2)
    lib = _lib.load()
    dev = image.device
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        mx = [torch.empty(max(n_frames, 1), dtype=torch.int32, device=dev) for _ in range(2)]
        for lo in range(0, n_frames, 65535):
            cnt = min(65535, n_frames - lo)
            ptr = image.data_ptr() + lo * h * w * 2
            cur = 0
            _lib.check(lib.cab200_defect_pass(ptr, cnt, h, w, _lib.DEFECT_MAX_ONLY, None, None, None,
                                              mx[cur].data_ptr(), stream), "cab200_defect_pass")
            for kind, data in ops:
                d = torch.from_numpy(data).to(dev)
                is_map = kind == _lib.DEFECT_MUL
                _lib.check(lib.cab200_defect_pass(ptr, cnt, h, w, kind, None if is_map else d.data_ptr(),
                                                  d.data_ptr() if is_map else None, mx[cur].data_ptr(),
                                                  mx[1 - cur].data_ptr(), stream), "cab200_defect_pass")
                cur = 1 - cur
        torch.cuda.current_stream(dev).synchronize()          # the uploaded tables may be freed now
    return image


def apply_all_defects(image: np.ndarray, cells_mask: Optional[np.ndarray], config: Optional[Dict[str, Any]],
                      device: Optional[torch.device] = None) -> np.ndarray:
    """Drop-in for the reference's ``apply_all_defects(image, cells_mask, config)``
    (``utils/image_processing.py:130-175``) for one ``(H, W, 2)`` uint8 frame: the supported defects of ``config``
    applied on the GPU (K5); a configuration with stochastic / OpenCV defects takes the host chain with the
    global numpy generator, like the reference."""
    if host_defects(config):
        return apply_all_defects_host(image, cells_mask, config, np.random)
    if image.ndim != 3 or image.shape[2] != 2 or image.dtype != np.uint8:
        raise ValueError(f"expected an (H, W, 2) uint8 frame, got {image.shape} {image.dtype}")
    if not torch.cuda.is_available():
        raise RuntimeError("calcium_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
    apply_defects(t, config, (image.shape[1], image.shape[0]))
    return t.cpu().numpy()


# ---- the per-frame random configuration ----------------------------------------------------------

def random_defect_config(rng=None) -> Dict[str, Any]:
    """``generate_random_defect_config`` (reference ``main.py:29-92``): same keys, same stdlib-``random`` calls
    in the same order (the reference draws one per frame when no ``defect_configs`` are given, ``:273-277``).
    Only the three background defects are ever switched on by it."""
    import random as _random
    r = rng if rng is not None else _random
    draws = (("background_fluorescence", "choice", (True, False)), ("background_intensity", "uniform", (0.05, 0.2)),
             ("non_uniform_background", "choice", (True, False)), ("spontaneous_luminescence", "choice", (True, False)),
             ("spontaneous_min", "uniform", (0.03, 0.1)), ("spontaneous_max", "uniform", (0.1, 0.2)),
             ("spontaneous_probability", "uniform", (0.1, 0.3)), ("cell_fragments", "choice", (True, False)),
             ("fragment_count", "randint", (3, 10)), ("fragment_min_size", "randint", (5, 15)),
             ("fragment_max_size", "randint", (15, 40)), ("fragment_min_intensity", "uniform", (0.05, 0.15)),
             ("fragment_max_intensity", "uniform", (0.15, 0.3)), ("chromatic_aberration", None, False),
             ("chromatic_offset", "randint", (1, 5)), ("vignetting", None, False),
             ("vignetting_strength", "uniform", (0.2, 0.7)), ("poisson_noise", None, False),
             ("poisson_scaling", "uniform", (0.5, 2.0)), ("readout_noise", None, False),
             ("readout_strength", "uniform", (0.02, 0.1)), ("gaussian_noise", None, False),
             ("gaussian_mean", None, 0), ("gaussian_sigma", "uniform", (0.05, 0.15)),
             ("dynamic_range_compression", None, False), ("gamma", "uniform", (0.6, 0.9)),
             ("quantization", None, False), ("bit_depth", "choice", (6, 8, 10)), ("defocus_blur", None, False),
             ("defocus_kernel_size", "choice", (5, 7, 9, 11)), ("defocus_sigma", "uniform", (1.0, 5.0)),
             ("partial_defocus", None, False), ("defocus_radius", "randint", (50, 200)),
             ("adjust_brightness_contrast", None, False), ("brightness", "uniform", (-0.1, 0.1)),
             ("contrast", "uniform", (0.8, 1.2)))
    cfg: Dict[str, Any] = {}
    for key, how, arg in draws:
        if how is None:
            cfg[key] = arg
        elif how == "choice":
            cfg[key] = r.choice(list(arg))
        elif how == "uniform":
            cfg[key] = r.uniform(*arg)
        else:
            cfg[key] = r.randint(*arg)
    return cfg


# ---- the host chain: stochastic and OpenCV defects, generator threaded ----------------------------

def _lut(image: np.ndarray, table: np.ndarray, by_max: bool) -> np.ndarray:
    row = int(image.max()) if by_max else int(image.max() > 1.0)
    return table[row][image]


_GRADIENT07: Dict[Tuple[int, int], np.ndarray] = {}


def _nonuniform_background(image, intensity, rs):                       # artifacts/background.py:21-45, 58
    import cv2
    max_val = np.max(image)
    rows, cols = image.shape[:2]
    g07 = _GRADIENT07.get((rows, cols))
    if g07 is None:                      # static per frame size: the same numpy expression, evaluated once
        yy, xx = np.mgrid[0:rows, 0:cols]
        gradient = np.sin(xx / cols * 3 * np.pi) * np.sin(yy / rows * 2 * np.pi) * 0.5 + 0.5
        g07 = gradient * 0.7
        g07.setflags(write=False)
        _GRADIENT07[(rows, cols)] = g07
    noise = cv2.GaussianBlur(rs.rand(rows, cols), (99, 99), 30)
    background = (g07 + noise * 0.3) * intensity * max_val
    out = image.copy()
    for ch in range(image.shape[2]):
        out[:, :, ch] = np.clip(out[:, :, ch] + background, 0, max_val)
    return out


def _spontaneous(image, cells_mask, lo, hi, probability, rs):           # artifacts/background.py:61-112
    ids = np.unique(cells_mask)
    ids = ids[ids > 0]
    chosen = rs.choice(ids, int(len(ids) * probability), replace=False)
    for _ in chosen:
        rs.uniform(lo, hi)
        if image.shape[2] < 3:          # the reference writes channels 0, 1, 2 (:99-107): IndexError on channel 2
            raise IndexError(f"index 2 is out of bounds for axis 2 with size {image.shape[2]}")
        raise NotImplementedError("spontaneous luminescence on frames with three or more channels")
    return image.copy()


def _fragments(image, count, size_lo, size_hi, int_lo, int_hi, rs):     # artifacts/background.py:115-167
    rows, cols = image.shape[:2]
    for _ in range(count):
        rs.randint(0, cols); rs.randint(0, rows)
        rs.randint(size_lo, size_hi)
        rs.uniform(int_lo, int_hi)
        if image.shape[2] < 3:          # channels 0, 1, 2 again (:153-161)
            raise IndexError(f"index 2 is out of bounds for axis 2 with size {image.shape[2]}")
        raise NotImplementedError("cell fragments on frames with three or more channels")
    return image.copy()


def _poisson(image, scaling, rs):                                       # artifacts/sensor.py:8-45
    big = image.max() > 1.0
    lam = (image / 255.0 if big else image.copy()) * 255.0 * scaling
    noisy = np.zeros_like(lam)
    for ch in range(image.shape[2]):
        noisy[:, :, ch] = rs.poisson(lam[:, :, ch])
    noisy = noisy / (255.0 * scaling)
    if big:
        noisy = noisy * 255.0
    return np.clip(noisy, 0, image.max()).astype(image.dtype)


def _readout(image, pattern, strength, rs):                             # artifacts/sensor.py:48-76
    if pattern is None:
        pattern = np.repeat(rs.normal(0, 1, (image.shape[0], image.shape[1], 1)), image.shape[2], axis=2)
    max_val = image.max()
    return np.clip(image + pattern * max_val * strength, 0, max_val).astype(image.dtype)


def _gaussian(image, mean, sigma, rs):                                  # artifacts/sensor.py:79-104
    max_val = image.max()
    noise = rs.normal(mean, max_val * sigma, image.shape)
    return np.clip(image + noise, 0, max_val).astype(image.dtype)


def _defocus(image, kernel_size, sigma):                                # utils/image_processing.py:34-52
    import cv2
    k = max(3, kernel_size)
    k += (k % 2 == 0)
    return cv2.GaussianBlur(image, (k, k), sigma)


def _partial_defocus(image, center, radius, strength, rs):              # utils/image_processing.py:55-93
    import cv2
    rows, cols = image.shape[:2]
    if center is None:
        center = (rs.randint(cols // 4, 3 * cols // 4), rs.randint(rows // 4, 3 * rows // 4))
    blurred = _defocus(image, 9, strength)
    mask = np.zeros((rows, cols), dtype=np.float32)
    cv2.circle(mask, center, radius, 1.0, -1)
    mask = cv2.GaussianBlur(mask, (radius // 2 * 2 + 1, radius // 2 * 2 + 1), radius / 3)
    mask = np.repeat(np.expand_dims(mask, axis=2), image.shape[2], axis=2)
    return (image * (1 - mask) + blurred * mask).astype(image.dtype)


def apply_all_defects_host(image: np.ndarray, cells_mask: Optional[np.ndarray], config: Optional[Dict[str, Any]],
                           rs=np.random) -> np.ndarray:
    """``apply_all_defects`` (reference ``utils/image_processing.py:130-175``) for one ``(H, W, 2)`` uint8 frame
    with every defect, on the host: the stochastic ones draw from ``rs`` (a ``np.random.RandomState`` -- the
    pouch's own -- or the ``np.random`` module) with the reference's calls in the reference's order, the
    deterministic ones go through the same tables K5 uses.  Raises ``IndexError`` where the reference does."""
    if image.ndim != 3 or image.shape[2] != 2 or image.dtype != np.uint8:
        raise ValueError(f"expected an (H, W, 2) uint8 frame, got {image.shape} {image.dtype}")
    cfg = config or {}
    g = cfg.get
    out = image.copy()
    size = (image.shape[1], image.shape[0])
    # 1. background (artifacts/background.py:170-207)
    if g("background_fluorescence", False):
        if g("non_uniform_background", True):
            out = _nonuniform_background(out, g("background_intensity", 0.1), rs)
        else:
            out = _lut(out, table_background(g("background_intensity", 0.1)), True)
    if g("spontaneous_luminescence", False):
        out = _spontaneous(out, cells_mask, g("spontaneous_min", 0.05), g("spontaneous_max", 0.15),
                           g("spontaneous_probability", 0.2), rs)
    if g("cell_fragments", False):
        out = _fragments(out, g("fragment_count", 5), g("fragment_min_size", 10), g("fragment_max_size", 30),
                         g("fragment_min_intensity", 0.1), g("fragment_max_intensity", 0.3), rs)
    # 2. optical (artifacts/optical.py:86-112); chromatic aberration copies frames below three channels
    if g("vignetting", False):
        vm = vignette_map(size, g("vignetting_strength", 0.5))
        out = (out * vm[:, :, None]).astype(np.uint8)
    # 3. sensor (artifacts/sensor.py:159-198)
    if g("poisson_noise", False):
        out = _poisson(out, g("poisson_scaling", 1.0), rs)
    if g("readout_noise", False):
        out = _readout(out, g("readout_pattern", None), g("readout_strength", 0.05), rs)
    if g("gaussian_noise", False):
        out = _gaussian(out, g("gaussian_mean", 0), g("gaussian_sigma", 0.1), rs)
    if g("dynamic_range_compression", False):
        out = _lut(out, table_gamma(g("gamma", 0.7)), False)
    if g("quantization", False):
        out = _lut(out, table_quantization(g("bit_depth", 8)), True)
    # 4. post-processing (utils/image_processing.py:155-172)
    if g("defocus_blur", False):
        if g("partial_defocus", False):
            out = _partial_defocus(out, g("defocus_center", None), g("defocus_radius", 100), g("defocus_sigma", 3), rs)
        else:
            out = _defocus(out, g("defocus_kernel_size", 9), g("defocus_sigma", 3))
    if g("adjust_brightness_contrast", False):
        out = _lut(out, table_brightness_contrast(g("brightness", 0), g("contrast", 1.0)), False)
    return out
