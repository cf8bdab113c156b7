"""K5 (cab200_defect_pass) on the GPU against outputs of the unmodified reference's apply_all_defects
(tests/golden/defects_xsmall.npz) -- SURVEY.md section 8 row f3, partial.  The CPU half (oracle, table builder) is
tests/test_defects.py."""
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


@pytest.mark.gpu
def test_k5_matches_the_reference(cuda_device, gold):
    import torch
    from calcium_b200 import build
    build.build()
    from calcium_b200 import apply_all_defects
    from calcium_b200.defects import apply_defects
    z, cfgs, cases = gold
    for key, name, img, osz in cases:
        assert np.array_equal(apply_all_defects(img, None, cfgs[name]), z[key]), key
    # many frames with different maxima in one call: every frame uses its own statistic
    img = np.load(os.path.join(os.path.dirname(__file__), "golden", "raster_xsmall.npz"))["image_160x120_10000"]
    frames = np.stack([img, img // 100, img // 3, np.zeros_like(img), img // 2 + 100])
    want = np.stack([DO.apply_supported_defects(f, cfgs["all"]) if f.max() > 0 else f for f in frames])
    t = torch.from_numpy(frames).to(cuda_device)
    apply_defects(t, cfgs["all"], (160, 120))
    got = t.cpu().numpy()
    for i in (0, 1, 2, 4):
        assert np.array_equal(got[i], want[i]), i


@pytest.mark.gpu
def test_batch_path_applies_deterministic_defects(cuda_device, tmp_path):
    """generate_simulation_batch(defect_configs=[...]) (reference main.py:273-293): the stored PNGs hold the frames
    after K5; a stochastic defect in the configuration is reported and skipped."""
    cv2 = pytest.importorskip("cv2")
    from calcium_b200 import generate_simulation_batch, ensemble as E
    cfg = {"vignetting": True, "vignetting_strength": 0.6, "quantization": True, "bit_depth": 6, "gaussian_noise": True}
    res = generate_simulation_batch(2, str(tmp_path), pouch_sizes=["xsmall"], sim_types=["Intercellular waves"],
                                    time_steps=[0, 9000], defect_configs=[cfg, None], image_size="160x120",
                                    generate_masks=False)
    assert [s["image_count"] for s in res["simulations"]] == [2, 2]
    specs = [E.PouchSpec.draw(res["simulations"][i]["parameters"], size="xsmall", sim_number=i) for i in range(2)]
    ens = E.Ensemble(specs, frame_steps=[0, 9000], device=cuda_device, output_size=(160, 120))
    ens.simulate()
    clean = ens.rasterize(image=True, instance=False)["image"].cpu().numpy()
    files = sorted(os.listdir(os.path.join(str(tmp_path), "images", "train")))
    for i in range(2):
        for f, t in enumerate((0, 9000)):
            name = [x for x in files if f"_{i}_t{t:05d}_" in x][0]
            png = cv2.imread(os.path.join(str(tmp_path), "images", "train", name), cv2.IMREAD_UNCHANGED)
            want = DO.apply_supported_defects(clean[i, f], cfg) if i == 0 else clean[i, f]
            assert np.array_equal(png[..., 0], want[..., 0].astype(np.uint16) * 4), (i, t)
            assert np.array_equal(png[..., -1], want[..., 1].astype(np.uint16) * 257), (i, t)
