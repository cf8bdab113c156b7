"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/defects_xsmall.npz``.

Run in the build container (``python -m oracle.make_golden_defects``), where ``/root/reference`` is mounted: the
*unmodified* ``apply_all_defects`` (``utils/image_processing.py:130-175``) applied to frames the reference itself
rendered (``tests/golden/raster_xsmall.npz``) with configurations that enable only the deterministic defects that
work on two-channel frames: uniform background fluorescence, vignetting, dynamic range compression, quantization,
brightness / contrast (and chromatic aberration, a no-op below three channels)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_shims as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    "bg": {"background_fluorescence": True, "background_intensity": 0.13, "non_uniform_background": False},
    "vig": {"vignetting": True, "vignetting_strength": 0.55},
    "gamma": {"dynamic_range_compression": True, "gamma": 0.7},
    "quant6": {"quantization": True, "bit_depth": 6},
    "quant10": {"quantization": True, "bit_depth": 10},
    "bc": {"adjust_brightness_contrast": True, "brightness": 0.07, "contrast": 1.15},
    "bc_dark": {"adjust_brightness_contrast": True, "brightness": -0.1, "contrast": 0.8},
    "chroma": {"chromatic_aberration": True, "chromatic_offset": 3},
    "all": {"background_fluorescence": True, "background_intensity": 0.08, "non_uniform_background": False,
            "chromatic_aberration": True, "chromatic_offset": 2, "vignetting": True, "vignetting_strength": 0.4,
            "dynamic_range_compression": True, "gamma": 0.85, "quantization": True, "bit_depth": 6,
            "adjust_brightness_contrast": True, "brightness": 0.03, "contrast": 1.1},
    "vig_quant": {"vignetting": True, "vignetting_strength": 0.7, "quantization": True, "bit_depth": 8},
}


def main() -> None:
    R.load_reference()
    from utils.image_processing import apply_all_defects
    z = np.load(os.path.join(OUT, "raster_xsmall.npz"))
    out = {"configs": np.array(json.dumps(CONFIGS))}
    for tag in ("160x120", "512x512"):
        cm = z[f"cells_mask_{tag}"].astype(np.int32)
        for t in (0, 10000):
            img = z[f"image_{tag}_{t}"]
            for name, cfg in CONFIGS.items():
                if tag == "512x512" and name not in ("all", "vig"):
                    continue
                res = apply_all_defects(img, cm, dict(cfg))
                assert res.dtype == np.uint8 and res.shape == img.shape
                out[f"{name}_{tag}_{t}"] = res
    # a frame whose every pixel is 0 or 1 takes the "already normalised" branches (image.max() <= 1)
    tiny = (z["image_160x120_10000"] > 128).astype(np.uint8)
    out["tiny_in"] = tiny
    for name in ("gamma", "bc", "quant6", "bg"):
        out[f"tiny_{name}"] = apply_all_defects(tiny, z["cells_mask_160x120"].astype(np.int32), dict(CONFIGS[name]))
    np.savez_compressed(os.path.join(OUT, "defects_xsmall.npz"), **out)
    print("wrote defects_xsmall.npz", os.path.getsize(os.path.join(OUT, "defects_xsmall.npz")), len(out))


if __name__ == "__main__":
    main()
