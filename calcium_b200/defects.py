"""The deterministic part of the reference's defect pipeline on the GPU (partial).

``apply_all_defects(image, cells_mask, config)`` (``utils/image_processing.py:130-175``) chains background,
optical, sensor and post-processing defects.  Supported here -- the ones that are deterministic and work on the
two-channel (gray, alpha) frames ``generate_image`` returns -- in the reference's order:

  background_fluorescence with non_uniform_background=False   artifacts/background.py:46-56
  chromatic_aberration (returns a copy below three channels)  artifacts/optical.py:22-24
  vignetting                                                  artifacts/optical.py:49-83
  dynamic_range_compression                                   artifacts/sensor.py:117-135
  quantization                                                artifacts/sensor.py:138-156
  adjust_brightness_contrast                                  utils/image_processing.py:97-127

Not supported (``unsupported_defects`` names them; the batch path reports and skips them): the non-uniform
background, Poisson / readout / Gaussian noise and partial defocus draw from the unseeded global generator;
``spontaneous_luminescence`` and ``cell_fragments`` index channel 2 of the two-channel frame and raise in the
reference itself; ``defocus_blur`` is OpenCV's fixed-point Gaussian blur.

Between two defects the reference's image is uint8 again, so a supported defect is a 256-entry table row chosen
by one statistic of the frame (``np.max(image)`` or ``image.max() > 1.0``), or a product with a static float64
map.  The rows are built HERE, on the host, by evaluating the reference's numpy expressions for every possible
(statistic, pixel value) pair -- same numpy, same rounding -- and K5 (``cab200_defect_pass``) applies them.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

STOCHASTIC = ("poisson_noise", "readout_noise", "gaussian_noise", "partial_defocus")
BROKEN_IN_REFERENCE = ("spontaneous_luminescence", "cell_fragments")


def unsupported_defects(config: Optional[Dict[str, Any]]) -> List[str]:
    """Names of the defects ``config`` switches on that this module does not apply."""
    if not config:
        return []
    bad = [k for k in STOCHASTIC + BROKEN_IN_REFERENCE if config.get(k, False) and (k != "partial_defocus" or config.get("defocus_blur", False))]
    if config.get("background_fluorescence", False) and config.get("non_uniform_background", True):
        bad.append("non_uniform_background")
    if config.get("defocus_blur", False):
        bad.append("defocus_blur")
    return bad


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
    applied on the GPU; ``cells_mask`` is only read by defects that are not supported.  Raises
    ``NotImplementedError`` when ``config`` switches on a defect this module does not apply."""
    bad = unsupported_defects(config)
    if bad:
        raise NotImplementedError("defects not supported on the GPU path: " + ", ".join(bad))
    if image.ndim != 3 or image.shape[2] != 2 or image.dtype != np.uint8:
        raise ValueError(f"expected an (H, W, 2) uint8 frame, got {image.shape} {image.dtype}")
    if not torch.cuda.is_available():
        raise RuntimeError("calcium_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
    apply_defects(t, config, (image.shape[1], image.shape[0]))
    return t.cpu().numpy()
