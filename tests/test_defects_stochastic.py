"""Row f3, stochastic part: ``calcium_b200.defects.apply_all_defects_host`` against outputs of the UNMODIFIED
reference ``apply_all_defects`` recorded with a seeded generator (``tests/golden/defects_stochastic.npz``,
``oracle/make_golden_defects_stochastic.py``): identical frames AND identical generator state afterwards, the same
exception where the reference raises on two-channel frames, and the same per-frame random configurations."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from calcium_b200 import defects as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLD, "defects_stochastic.npz"))
    r = np.load(os.path.join(GOLD, "raster_xsmall.npz"))
    return z, r


def _digest(rs) -> str:
    st = rs.get_state()
    return hashlib.sha1(st[1].tobytes() + str(st[2:]).encode()).hexdigest()


def test_stochastic_defects_equal_the_reference_bit_for_bit(gold):
    z, r = gold
    configs = json.loads(str(z["configs"]))
    digests = json.loads(str(z["digests"]))
    cm = r["cells_mask_160x120"].astype(np.int32)
    for t in (0, 10000):
        img = r[f"image_160x120_{t}"]
        for i, (name, cfg) in enumerate(configs.items()):
            if "defocus_center" in cfg:
                cfg = dict(cfg, defocus_center=tuple(cfg["defocus_center"]))
            rs = np.random.RandomState(1000 + i + t)               # a private generator, like the batch path uses
            got = D.apply_all_defects_host(img, cm, cfg, rs)
            assert got.dtype == np.uint8 and np.array_equal(got, z[f"{name}_{t}"]), (name, t)
            assert _digest(rs) == digests[f"{name}_{t}"], (name, t, "generator state")
    tiny = (r["image_160x120_10000"] > 128).astype(np.uint8)
    for i, name in enumerate(("poisson", "readout", "gaussian", "bg_nonuniform")):
        rs = np.random.RandomState(3000 + i)
        assert np.array_equal(D.apply_all_defects_host(tiny, cm, configs[name], rs), z[f"tiny_{name}"]), name
        assert _digest(rs) == digests[f"tiny_{name}"]


def test_defects_that_raise_in_the_reference_raise_here_with_the_same_draws(gold):
    z, r = gold
    raising = json.loads(str(z["raising"]))
    digests = json.loads(str(z["digests"]))
    cm = r["cells_mask_160x120"].astype(np.int32)
    for t in (0, 10000):
        img = r[f"image_160x120_{t}"]
        for i, (name, cfg) in enumerate(raising.items()):
            rs = np.random.RandomState(2000 + i + t)
            with pytest.raises(IndexError):
                D.apply_all_defects_host(img, cm, cfg, rs)
            assert digests[f"{name}_{t}"] == "IndexError " + _digest(rs), (name, t)


def test_global_generator_and_dropin_signature(gold):
    """``apply_all_defects`` (the drop-in) with the module-level generator, like the reference's callers."""
    z, r = gold
    configs = json.loads(str(z["configs"]))
    cm = r["cells_mask_160x120"].astype(np.int32)
    img = r["image_160x120_10000"]
    i = list(configs).index("gaussian")
    np.random.seed(1000 + i + 10000)
    assert np.array_equal(D.apply_all_defects(img, cm, configs["gaussian"]), z["gaussian_10000"])


def test_random_defect_config_draws_like_the_reference(gold):
    z, _ = gold
    want = json.loads(str(z["random_configs"]))
    random.seed(424242)
    got = [D.random_defect_config() for _ in range(5)]
    assert [list(c) for c in got] == [list(c) for c in want]          # same keys, same order
    assert json.loads(json.dumps(got)) == want
    assert D.host_defects(got[0]) == [k for k in ("spontaneous_luminescence", "cell_fragments") if got[0][k]] + (
        ["non_uniform_background"] if got[0]["background_fluorescence"] and got[0]["non_uniform_background"] else [])
