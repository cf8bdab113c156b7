"""CPU restatement of the deterministic part of the reference's defect pipeline -- TEST INFRASTRUCTURE ONLY.

Whole-image numpy code following ``apply_all_defects`` (``utils/image_processing.py:130-175``) and the functions it
calls, statement by statement, for the defects ``calcium_b200.defects`` supports (the product builds 256-entry
tables from the same formulas; this applies them to the image like the reference does).  Pinned by
``tests/golden/defects_xsmall.npz`` (outputs of the unmodified reference, ``oracle/make_golden_defects.py``)."""
from __future__ import annotations

import numpy as np


def background_uniform(image: np.ndarray, intensity: float) -> np.ndarray:          # artifacts/background.py:21,46-58
    result = image.copy()
    max_val = np.max(image)
    uniform_value = intensity * max_val
    for i in range(image.shape[2]):
        result[:, :, i] = np.clip(result[:, :, i] + uniform_value, 0, max_val)
    return result.astype(image.dtype)


def vignetting(image: np.ndarray, strength: float) -> np.ndarray:                  # artifacts/optical.py:59-83
    rows, cols = image.shape[:2]
    x = np.linspace(-1, 1, cols)
    y = np.linspace(-1, 1, rows)
    xv, yv = np.meshgrid(x, y)
    r = np.sqrt(xv ** 2 + yv ** 2)
    vignette = 1 - strength * np.clip(r, 0, 1) ** 2
    result = image.copy()
    for i in range(image.shape[2]):
        result[:, :, i] = image[:, :, i] * vignette
    return result.astype(image.dtype)


def dynamic_range_compression(image: np.ndarray, gamma: float) -> np.ndarray:      # artifacts/sensor.py:117-135
    normalized = image / 255.0 if image.max() > 1.0 else image.copy()
    compressed = np.power(normalized, gamma)
    if image.max() > 1.0:
        compressed = compressed * 255.0
    return compressed.astype(image.dtype)


def quantization(image: np.ndarray, bits: int) -> np.ndarray:                      # artifacts/sensor.py:138-156
    max_val = image.max()
    levels = 2 ** bits
    step = max_val / levels
    quantized = np.round(image / step) * step
    return np.clip(quantized, 0, max_val).astype(image.dtype)


def brightness_contrast(image: np.ndarray, brightness: float, contrast: float) -> np.ndarray:   # utils/image_processing.py:97-127
    img_float = image.astype(np.float32)
    if img_float.max() > 1.0:
        img_float /= 255.0
    img_float = (img_float - 0.5) * contrast + 0.5
    img_float += brightness
    img_float = np.clip(img_float, 0, 1.0)
    if image.max() > 1.0:
        img_float *= 255.0
    return img_float.astype(image.dtype)


def apply_supported_defects(image: np.ndarray, config: dict) -> np.ndarray:        # utils/image_processing.py:130-175
    result = image.copy()
    if config.get("background_fluorescence", False) and not config.get("non_uniform_background", True):
        result = background_uniform(result, config.get("background_intensity", 0.1))
    if config.get("vignetting", False):
        result = vignetting(result, config.get("vignetting_strength", 0.5))
    if config.get("dynamic_range_compression", False):
        result = dynamic_range_compression(result, config.get("gamma", 0.7))
    if config.get("quantization", False):
        result = quantization(result, config.get("bit_depth", 8))
    if config.get("adjust_brightness_contrast", False):
        result = brightness_contrast(result, config.get("brightness", 0), config.get("contrast", 1.0))
    return result
