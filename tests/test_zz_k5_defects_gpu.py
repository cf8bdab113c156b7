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
    after K5."""
    cv2 = pytest.importorskip("cv2")
    from calcium_b200 import generate_simulation_batch, ensemble as E
    cfg = {"vignetting": True, "vignetting_strength": 0.6, "quantization": True, "bit_depth": 6}
    res = generate_simulation_batch(2, str(tmp_path), pouch_sizes=["xsmall"], sim_types=["Intercellular waves"],
                                    time_steps=[0, 9000], defect_configs=[cfg, {}], image_size="160x120",
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


@pytest.mark.gpu
def test_batch_path_stochastic_and_default_random_defects(cuda_device, tmp_path):
    """(1) A configuration with stochastic defects runs the host chain with the pouch's own generator: the stored
    frames equal ``apply_all_defects_host`` fed with the generator state the pouch's draws left behind, frame after
    frame.  (2) Without ``defect_configs`` every frame draws ``generate_random_defect_config()`` from the stdlib
    generator in the reference's order (main.py:216-217, 273-277); frames whose configuration raises in the
    reference are not written."""
    cv2 = pytest.importorskip("cv2")
    import random
    from calcium_b200 import generate_simulation_batch, ensemble as E, defects as D
    from calcium_b200.parameters import SimulationParameters as SP
    steps = [0, 6000, 12000]

    def clean_and_state(sim_type, i):
        spec = E.PouchSpec.draw(SP(sim_type).generate_random_params(seed=i), size="xsmall", sim_number=i)
        rs = np.random.RandomState(); rs.set_state(np.random.get_state())
        ens = E.Ensemble([spec], frame_steps=steps, device=cuda_device, output_size=(160, 120))
        ens.simulate()
        img = ens.rasterize(image=True, instance=False)["image"].cpu().numpy()[0]
        cm = E._u16_numpy(ens.label_maps()[1])[0].astype(np.int32)
        return img, cm, rs

    def stored(d, i, t):
        hit = [x for x in os.listdir(os.path.join(d, "images", "train")) if f"_{i}_t{t:05d}_" in x]
        if not hit:
            return None
        return cv2.imread(os.path.join(d, "images", "train", hit[0]), cv2.IMREAD_UNCHANGED)

    # (1) explicit stochastic configuration
    cfg = {"background_fluorescence": True, "background_intensity": 0.1, "poisson_noise": True, "poisson_scaling": 1.2,
           "gaussian_noise": True, "gaussian_sigma": 0.05, "quantization": True, "bit_depth": 6}
    d1 = str(tmp_path / "s")
    res = generate_simulation_batch(2, d1, pouch_sizes=["xsmall"], sim_types=["Fluttering"], time_steps=steps,
                                    defect_configs=[cfg, cfg], image_size="160x120", generate_prompts=False)
    assert [s["image_count"] for s in res["simulations"]] == [3, 3]
    for i in range(2):
        img, cm, rs = clean_and_state("Fluttering", i)
        for f, t in enumerate(steps):
            want = D.apply_all_defects_host(img[f], cm, cfg, rs)
            png = stored(d1, i, t)
            assert np.array_equal(png[..., 0], want[..., 0].astype(np.uint16) * 4), (i, t)
            assert np.array_equal(png[..., -1], want[..., 1].astype(np.uint16) * 257), (i, t)
    # (2) default: per-frame random configurations
    d2 = str(tmp_path / "r")
    random.seed(77)
    res = generate_simulation_batch(3, d2, pouch_sizes=["xsmall"], sim_types=["Fluttering", "Intercellular waves"],
                                    time_steps=steps, image_size="160x120", generate_prompts=False)
    random.seed(77)
    n_written = 0
    for i in range(3):
        random.choice(["xsmall"]); st = random.choice(["Fluttering", "Intercellular waves"])
        cfgs = [D.random_defect_config() for _ in steps]
        assert res["simulations"][i]["simulation_type"] == st
        img, cm, rs = clean_and_state(st, i)
        kept = 0
        for f, t in enumerate(steps):
            png = stored(d2, i, t)
            try:
                want = D.apply_all_defects_host(img[f], cm, cfgs[f], rs)
            except IndexError:
                assert png is None, (i, t, "the reference skips this frame")
                continue
            kept += 1
            assert png is not None and np.array_equal(png[..., 0], want[..., 0].astype(np.uint16) * 4), (i, t)
        assert res["simulations"][i]["image_count"] == kept
        n_written += kept
    assert n_written == len(os.listdir(os.path.join(d2, "images", "train")))
    # broken_defects="off": the two raising defects are switched off, every frame is written
    random.seed(77)
    res = generate_simulation_batch(3, str(tmp_path / "o"), pouch_sizes=["xsmall"], sim_types=["Fluttering", "Intercellular waves"],
                                    time_steps=steps, image_size="160x120", generate_prompts=False, broken_defects="off")
    assert [s["image_count"] for s in res["simulations"]] == [3, 3, 3]
