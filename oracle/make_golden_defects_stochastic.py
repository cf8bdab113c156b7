"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/defects_stochastic.npz``.

Run in the build container (``python -m oracle.make_golden_defects_stochastic``), where ``/root/reference`` is
mounted: the *unmodified* ``apply_all_defects`` (``utils/image_processing.py:130-175``) on frames the reference
itself rendered (``tests/golden/raster_xsmall.npz``), with the global numpy generator seeded before every call,
for configurations that switch on the stochastic and OpenCV defects (non-uniform background, Poisson / readout /
Gaussian noise, defocus, partial defocus) alone and chained with deterministic ones.  Recorded per case: the
output frame and a digest of the generator state the call left behind.  Also: the configurations
``generate_random_defect_config`` (``main.py:29-92``) draws from a seeded stdlib generator, and what the
reference does for the two defects that raise on two-channel frames: the exception type and the
generator state at that point."""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_shims as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    "bg_nonuniform": {"background_fluorescence": True, "background_intensity": 0.17},
    "poisson": {"poisson_noise": True, "poisson_scaling": 0.8},
    "poisson_hi": {"poisson_noise": True, "poisson_scaling": 1.9},
    "readout": {"readout_noise": True, "readout_strength": 0.07},
    "gaussian": {"gaussian_noise": True, "gaussian_mean": 0, "gaussian_sigma": 0.11},
    "defocus": {"defocus_blur": True, "defocus_kernel_size": 7, "defocus_sigma": 2.3},
    "defocus_even": {"defocus_blur": True, "defocus_kernel_size": 4, "defocus_sigma": 1.2},
    "partial": {"defocus_blur": True, "partial_defocus": True, "defocus_radius": 37, "defocus_sigma": 2.0},
    "partial_center": {"defocus_blur": True, "partial_defocus": True, "defocus_center": (60, 50), "defocus_radius": 30,
                       "defocus_sigma": 3.0},
    "chain": {"background_fluorescence": True, "background_intensity": 0.09, "vignetting": True,
              "vignetting_strength": 0.4, "poisson_noise": True, "poisson_scaling": 1.3, "readout_noise": True,
              "readout_strength": 0.04, "gaussian_noise": True, "gaussian_sigma": 0.06,
              "dynamic_range_compression": True, "gamma": 0.8, "quantization": True, "bit_depth": 6,
              "defocus_blur": True, "defocus_kernel_size": 5, "defocus_sigma": 1.5,
              "adjust_brightness_contrast": True, "brightness": 0.04, "contrast": 1.1},
}
RAISING = {
    "spont": {"spontaneous_luminescence": True, "spontaneous_probability": 0.2},
    "frag": {"cell_fragments": True, "fragment_count": 4},
    "bg_then_spont": {"background_fluorescence": True, "non_uniform_background": False, "spontaneous_luminescence": True},
}


def state_digest() -> str:
    st = np.random.get_state()
    return hashlib.sha1(st[1].tobytes() + str(st[2:]).encode()).hexdigest()


def main() -> None:
    R.load_reference()
    from utils.image_processing import apply_all_defects
    z = np.load(os.path.join(OUT, "raster_xsmall.npz"))
    out = {"configs": np.array(json.dumps(CONFIGS)), "raising": np.array(json.dumps(RAISING))}
    digests = {}
    cm = z["cells_mask_160x120"].astype(np.int32)
    for t in (0, 10000):
        img = z[f"image_160x120_{t}"]
        for i, (name, cfg) in enumerate(CONFIGS.items()):
            np.random.seed(1000 + i + t)
            res = apply_all_defects(img, cm, dict(cfg))
            assert res.dtype == np.uint8 and res.shape == img.shape, (name, res.dtype, res.shape)
            out[f"{name}_{t}"] = res
            digests[f"{name}_{t}"] = state_digest()
        for i, (name, cfg) in enumerate(RAISING.items()):
            np.random.seed(2000 + i + t)
            try:
                apply_all_defects(img, cm, dict(cfg))
                digests[f"{name}_{t}"] = "no exception " + state_digest()
            except Exception as e:  # noqa: BLE001
                digests[f"{name}_{t}"] = type(e).__name__ + " " + state_digest()
    # the all-0/1 frame (image.max() <= 1 branches)
    tiny = (z["image_160x120_10000"] > 128).astype(np.uint8)
    for i, name in enumerate(("poisson", "readout", "gaussian", "bg_nonuniform")):
        np.random.seed(3000 + i)
        out[f"tiny_{name}"] = apply_all_defects(tiny, cm, dict(CONFIGS[name]))
        digests[f"tiny_{name}"] = state_digest()
    out["digests"] = np.array(json.dumps(digests))
    # the per-frame random configuration (stdlib generator, seeded here)
    import main as ref_main
    random.seed(424242)
    out["random_configs"] = np.array(json.dumps([ref_main.generate_random_defect_config() for _ in range(5)]))
    np.savez_compressed(os.path.join(OUT, "defects_stochastic.npz"), **out)
    print("wrote defects_stochastic.npz", os.path.getsize(os.path.join(OUT, "defects_stochastic.npz")), len(out))
    print(digests)


if __name__ == "__main__":
    main()
