"""Prompt generation: calcium_b200/prompts.py on K4 against prompts the reference's own
``generate_point_prompts_from_mask`` / ``generate_negative_prompts`` produced (tests/golden/prompts_small.*,
recorded by oracle/make_golden.py), and the oracle restatement against the same fixtures."""
import json
import os

import numpy as np
import pytest

CASES = [(t, thr) for t in (2000, 9000, 16000) for thr in (0.1, 0.0)]


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(os.path.join(golden_dir, "prompts_small.npz"))
    j = json.load(open(os.path.join(golden_dir, "prompts_small.json")))
    return z, j


def _norm(pos):
    return {int(k): v for k, v in pos.items()}


@pytest.mark.parametrize("t,thr", CASES)
def test_oracle_prompts_match_the_reference(gold, t, thr):
    from oracle import prompts_oracle as PO
    z, j = gold
    g = j[f"{t}_{thr}"]
    inst = z[f"instance_{t}_{thr}"]
    np.random.seed(g["seed"])
    assert PO.point_prompts(inst) == _norm(g["positive"])
    assert PO.negative_prompts(inst, 5) == g["negative"]


@pytest.mark.parametrize("t,thr", CASES)
def test_instance_lut_reproduces_the_reference_instance_mask(gold, t, thr):
    """Host restatement of the device LUT: lut[cells_mask] is the reference's instance mask (id mismatch included)."""
    pytest.importorskip("torch")
    from calcium_b200.prompts import instance_lut
    z, j = gold
    ca = z["ca"][list(z["t"]).index(t)]
    lut = instance_lut(ca, threshold=thr, quirk=True)
    assert np.array_equal(lut[z["cells_mask"]], z[f"instance_{t}_{thr}"])
    if thr == 0.0:
        assert lut[0] == 1          # cell 0 active: every zero pixel becomes instance 1


def test_fixture_covers_the_background_instance(gold):
    z, j = gold
    big = [k for k, g in j.items() if "1" in g["positive"] and g["positive"]["1"]["area"] > 10000]
    assert big, "no frame in which the id mismatch labels the background"
    assert any(len(g["positive"]) > 100 for g in j.values())


@pytest.mark.gpu
def test_k4_is_scipys_distance_transform(cuda_device):
    """cab200_edt_classes against scipy.ndimage.distance_transform_edt: squared distances exact."""
    import torch
    from scipy.ndimage import distance_transform_edt
    from calcium_b200 import build
    build.build()
    from calcium_b200.prompts import edt_squared
    rng = np.random.default_rng(2)
    for (h, w) in ((64, 80), (200, 333), (512, 512)):
        lab = np.zeros((h, w), np.uint16)
        for k in range(1, 40):                       # random rectangles and discs, some nested, some at the border
            y, x = rng.integers(0, h), rng.integers(0, w)
            ry, rx = rng.integers(1, h // 3), rng.integers(1, w // 3)
            yy, xx = np.ogrid[:h, :w]
            m = ((yy - y) / ry) ** 2 + ((xx - x) / rx) ** 2 <= 1 if k % 2 else (abs(yy - y) <= ry) & (abs(xx - x) <= rx)
            lab[m] = k
        dlab = torch.from_numpy(lab.view(np.int16)).to(cuda_device)
        d2 = edt_squared(dlab)[0].cpu().numpy()
        for k in np.unique(lab):
            b = lab == k
            if b.all():
                continue
            ref = distance_transform_edt(b.astype(np.uint8))
            assert np.array_equal(np.sqrt(d2[b].astype(np.float64)), ref[b]), (h, w, int(k))
        # class tables: binary foreground / background in one call
        luts = np.stack([(np.arange(40) > 0), (np.arange(40) % 3 == 0)]).astype(np.uint16)
        dd = edt_squared(dlab, torch.from_numpy(luts.view(np.int16)).to(cuda_device)).cpu().numpy()
        for i in range(2):
            cls = luts[i][lab]
            for v in (0, 1):
                b = cls == v
                if b.any() and not b.all():
                    ref = distance_transform_edt(b.astype(np.uint8))
                    assert np.array_equal(np.sqrt(dd[i][b].astype(np.float64)), ref[b])


@pytest.mark.gpu
@pytest.mark.parametrize("t,thr", CASES)
def test_gpu_prompts_match_the_reference(cuda_device, gold, t, thr):
    """Same generator state in, same prompts out -- points, areas, confidences, negatives -- as the reference."""
    import torch
    from calcium_b200 import build
    build.build()
    from calcium_b200 import ensemble as E
    from calcium_b200.prompts import PromptTables, frame_prompts
    z, j = gold
    g = j[f"{t}_{thr}"]
    spec = E.make_specs(1, size="small")[0]
    ens = E.Ensemble([spec], T=10, frame_steps=[0], output_size=(256, 256), device=cuda_device)
    lmsk = ens.label_maps()[1][0]
    assert np.array_equal(lmsk.cpu().numpy().view(np.uint16), z["cells_mask"])
    tables = PromptTables(lmsk, spec.geometry.n_cells)
    ca = z["ca"][list(z["t"]).index(t)]
    np.random.seed(g["seed"])
    pos, neg = frame_prompts(tables, ca, threshold=thr, quirk=True)
    assert pos == _norm(g["positive"])
    assert neg == g["negative"]
