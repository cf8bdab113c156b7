"""the deterministic defects of apply_all_defects
(reference utils/image_processing.py:130-175) that work on two-channel frames.  CPU: the oracle restatement and the
product's table builder against outputs of the unmodified reference (tests/golden/defects_xsmall.npz, recorded by
oracle/make_golden_defects.py).  GPU: K5 (cab200_defect_pass) against the same fixtures."""
import json
import os

import numpy as np
import pytest

from oracle import defects_oracle as DO


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(os.path.join(golden_dir, "defects_xsmall.npz"))
    rz = np.load(os.path.join(golden_dir, "raster_xsmall.npz"))
    cases = []
    for key in z.files:
        if key in ("configs", "tiny_in"):
            continue
        if key.startswith("tiny_"):
            cases.append((key, key[5:], z["tiny_in"], (160, 120)))
        else:
            name, tag, t = key.rsplit("_", 2)
            cases.append((key, name, rz[f"image_{tag}_{t}"], tuple(map(int, tag.split("x")))))
    return z, json.loads(str(z["configs"])), cases


def test_oracle_defects_match_the_reference(gold):
    z, cfgs, cases = gold
    assert len(cases) >= 28
    for key, name, img, _ in cases:
        assert np.array_equal(DO.apply_supported_defects(img, cfgs[name]), z[key]), key
    assert any(not np.array_equal(z[key], img) for key, _, img, _ in cases)


def test_tables_reproduce_the_reference(gold):
    """The 256-entry rows the product builds on the host + the static vignette map, applied with plain numpy
    indexing: what K5 does on the device."""
    pytest.importorskip("torch")
    from calcium_b200 import defects as D, _lib
    z, cfgs, cases = gold
    for key, name, img, osz in cases:
        cur = img.copy()
        for kind, data in D.plan(cfgs[name], osz):
            m = int(cur.max())
            if kind == _lib.DEFECT_LUT_MAX:
                cur = data[m][cur]
            elif kind == _lib.DEFECT_LUT_FLAG:
                cur = data[1 if m > 1 else 0][cur]
            else:
                cur = (cur * data[..., None]).astype(np.int64).astype(np.uint8)
        assert np.array_equal(cur, z[key]), key


def test_host_defects_are_named():
    pytest.importorskip("torch")
    from calcium_b200 import defects as D
    assert D.unsupported_defects(None) == [] and D.unsupported_defects({"vignetting": True}) == []
    cfg = {"background_fluorescence": True, "spontaneous_luminescence": True, "gaussian_noise": True, "defocus_blur": True,
           "cell_fragments": False}
    assert sorted(D.unsupported_defects(cfg)) == ["defocus_blur", "gaussian_noise", "non_uniform_background",
                                                  "spontaneous_luminescence"]
    assert D.plan(cfg, (64, 64)) == []                 # the non-uniform background is stochastic: nothing to apply
    assert len(D.plan({"chromatic_aberration": True, "quantization": True}, (64, 64))) == 1   # aberration: no-op on 2 channels
    # such a configuration takes the host chain; this one raises like the reference does (spontaneous luminescence
    # writes channel 2 of a two-channel frame)
    with pytest.raises(IndexError):
        D.apply_all_defects(np.full((8, 8, 2), 9, np.uint8), np.arange(64, dtype=np.int32).reshape(8, 8), cfg)
