"""the edge_blur / with_border branches of Pouch.generate_image
(reference core/pouch.py:436-481).  CPU part: the oracle (oracle/edges_oracle.py) against the fixtures
recorded from the reference and against the real cv2 calls it restates.  GPU part: K0c / K2b
(cab200_edge_maps, cab200_raster_edges_batch) against the oracle, bit for bit."""
import os

import numpy as np
import pytest

from calcium_b200 import geometry as G
from oracle import edges_oracle as EO

# kernel sizes whose result cannot depend on float32 accumulation details (odd sizes -- the GUI offers
# 3, 5, 7, 9 -- and powers of two); 6 and 10 have exact .5 ties with an inexact reciprocal, where the real
# cv2 switches between FMA (SIMD body) and mul+add (row tail)
SIZES = [1, 2, 3, 4, 5, 7, 8, 9]


def _polys(size, output_size):
    g = G.get_geometry(size)
    offs, pts = g.scaled_polygons(output_size)
    return g, [pts[offs[i]:offs[i + 1]] for i in range(g.n_cells)]


# ---- CPU: the oracle is pinned ---------------------------------------------------------------------

def test_oracle_against_reference_fixtures(golden_dir):
    z = np.load(os.path.join(golden_dir, "raster_xsmall.npz"))
    _, polys = _polys("xsmall", (160, 120))
    img = z["image_160x120_10000"]
    got = EO.edge_effects_oracle(img, polys, (160, 120), False, True, 3, "mean")
    assert np.array_equal(got, z["edge_mean3_160x120_10000"])
    got = EO.edge_effects_oracle(img, polys, (160, 120), False, True, 5, "motion")
    assert np.array_equal(got, z["edge_motion5_160x120_10000"])
    got = EO.edge_effects_oracle(img, polys, (160, 120), True, True, 3, "mean")
    assert np.array_equal(got, z["border_blur_160x120_10000"])
    assert not np.array_equal(z["edge_mean3_160x120_10000"], img)        # the branch does something


def test_oracle_polylines_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for _ in range(400):
        W, H = rng.randint(8, 200), rng.randint(8, 200)
        n = rng.randint(2, 8)
        pts = np.stack([rng.randint(0, W, n), rng.randint(0, H, n)], 1).astype(np.int32)   # scaled vertices never leave the image
        want = np.zeros((H, W), np.uint8)
        cv2.polylines(want, [pts.reshape(-1, 1, 2)], True, 255, 1)
        assert np.array_equal(EO.polylines_oracle([pts], (W, H)), want)
    for size, osz in (("xsmall", (160, 120)), ("small", (256, 256))):
        _, polys = _polys(size, osz)
        want = np.zeros((osz[1], osz[0]), np.uint8)
        for p in polys:
            cv2.polylines(want, [p.reshape(-1, 1, 2).astype(np.int32)], True, 255, 1)
        assert np.array_equal(EO.polylines_oracle(polys, osz), want)


def test_oracle_filter_and_dilate_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(1)
    for it in range(40):
        W, H = rng.randint(3, 90), rng.randint(3, 90)
        src = rng.randint(0, 256, (H, W, 2) if it % 2 else (H, W)).astype(np.uint8)
        if it % 3 == 0:
            src = ((src > 128) * 255).astype(np.uint8)
        for s in SIZES:
            for t in ("mean", "motion"):
                k = EO.blur_kernel(s, t)
                assert np.array_equal(EO.filter2d_oracle(src, k), cv2.filter2D(src, -1, k)), (s, t, W, H)
        one = src if src.ndim == 2 else np.ascontiguousarray(src[..., 0])
        assert np.array_equal(EO.dilate3_oracle(one), cv2.dilate(one, np.ones((3, 3), np.uint8), iterations=1))


# ---- GPU: K0c / K2b against the oracle ---------------------------------------------------------------

@pytest.fixture(scope="module")
def env(cuda_device):
    from calcium_b200 import ensemble as E
    from oracle import pouch_oracle as O
    O.build()
    return {"E": E, "O": O, "dev": cuda_device}


@pytest.mark.gpu
@pytest.mark.parametrize("size,osz", [("xsmall", (160, 120)), ("small", (256, 256)), ("medium", (512, 512)),
                                      ("xsmall", (97, 64))])
def test_edge_maps_match_oracle(env, size, osz):
    E = env["E"]
    g, polys = _polys(size, osz)
    want_edges = EO.polylines_oracle(polys, osz)
    combos = [(3, "mean"), (5, "motion")] if size == "medium" else [(s, t) for s in SIZES for t in ("mean", "motion")]
    for ks, bt in combos:
        edges, dil, _ = E.build_edge_maps(g, osz, env["dev"], ks, bt)
        assert np.array_equal(edges.cpu().numpy(), want_edges), (ks, bt)
        want = EO.dilate3_oracle(EO.filter2d_oracle(want_edges, EO.blur_kernel(ks, bt)))
        assert np.array_equal(dil.cpu().numpy(), want), (ks, bt)


@pytest.mark.gpu
def test_raster_edges_batch_matches_oracle(env):
    """K2b over a small ensemble, every mode / kernel: image == oracle post-processing of K2's image."""
    E = env["E"]
    osz = (160, 120)
    specs = E.make_specs(3, size="xsmall", first_seed=2)
    T, steps = 3000, [0, 1500, 2999]
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"], output_size=osz)
    ens.simulate()
    base = ens.rasterize(image=True, instance=True)
    img0 = base["image"].cpu().numpy()
    ins0 = base["instance"].cpu().numpy()
    _, polys = _polys("xsmall", osz)
    cases = [(False, True, s, t) for s in SIZES for t in ("mean", "motion")] + [(True, True, 3, "mean"), (True, False, 3, "mean")]
    for wb, eb, ks, bt in cases:
        res = ens.rasterize(image=True, instance=True, with_border=wb, edge_blur=eb, blur_kernel_size=ks, blur_type=bt)
        got = res["image"].cpu().numpy()
        assert np.array_equal(res["instance"].cpu().numpy(), ins0)
        for i in range(len(specs)):
            for f in range(len(steps)):
                want = EO.edge_effects_oracle(img0[i, f], polys, osz, wb, eb, ks, bt)
                assert np.array_equal(got[i, f], want), (wb, eb, ks, bt, i, f)
    # a sub-range of pouches writes the same frames
    part = ens.rasterize(image=True, instance=False, pouches=(1, 3), edge_blur=True, blur_kernel_size=5, blur_type="motion")
    full = ens.rasterize(image=True, instance=False, edge_blur=True, blur_kernel_size=5, blur_type="motion")
    assert np.array_equal(part["image"].cpu().numpy(), full["image"].cpu().numpy()[1:3])


@pytest.mark.gpu
def test_raster_edges_odd_width_takes_the_generic_path(env):
    """A width that is not a multiple of 4 (and kernels wider than the fast path's 3/5/7/9) runs the
    one-pixel-per-thread kernel: same bytes as the oracle."""
    E = env["E"]
    osz = (75, 50)
    specs = E.make_specs(2, size="xsmall", first_seed=6)
    steps = [0, 999]
    ens = E.Ensemble(specs, T=1000, frame_steps=steps, device=env["dev"], output_size=osz)
    ens.simulate()
    img0 = ens.rasterize(image=True, instance=False)["image"].cpu().numpy()
    _, polys = _polys("xsmall", osz)
    for wb, eb, ks, bt in [(False, True, 3, "mean"), (False, True, 5, "motion"), (False, True, 10, "mean"),
                           (False, True, 4, "motion"), (True, True, 3, "mean"), (True, False, 3, "mean")]:
        got = ens.rasterize(image=True, instance=False, with_border=wb, edge_blur=eb, blur_kernel_size=ks,
                            blur_type=bt)["image"].cpu().numpy()
        for i in range(2):
            for f in range(2):
                assert np.array_equal(got[i, f], EO.edge_effects_oracle(img0[i, f], polys, osz, wb, eb, ks, bt)), (wb, eb, ks, bt)


@pytest.mark.gpu
def test_raster_edges_full_size_and_cv2(env):
    """512x512 medium frames (BASELINE config 3's image size): K2b == the reference's literal cv2 / numpy
    sequence (core/pouch.py:455-467) applied to K2's image."""
    cv2 = pytest.importorskip("cv2")
    E = env["E"]
    osz = (512, 512)
    specs = E.make_specs(2, size="medium", first_seed=3)          # a Fluttering and a spikes pouch
    steps = [0, 2000, 3999]
    ens = E.Ensemble(specs, T=4000, frame_steps=steps, device=env["dev"], output_size=osz)
    ens.simulate()
    img0 = ens.rasterize(image=True, instance=False)["image"].cpu().numpy()
    _, polys = _polys("medium", osz)
    edges = np.zeros((512, 512), np.uint8)
    for p in polys:
        cv2.polylines(edges, [p.reshape(-1, 1, 2).astype(np.int32)], True, 255, 1)
    for ks, bt in ((3, "mean"), (5, "motion"), (9, "mean")):
        got = ens.rasterize(image=True, instance=False, edge_blur=True, blur_kernel_size=ks,
                            blur_type=bt)["image"].cpu().numpy()
        k = EO.blur_kernel(ks, bt)
        m = cv2.dilate(cv2.filter2D(edges, -1, k), np.ones((3, 3), np.uint8), iterations=1) / 255.0
        m2 = np.stack([m, m], axis=-1)
        for i in range(2):
            for f in range(3):
                img = img0[i, f]
                want = (img * (1 - m2) + cv2.filter2D(img, -1, k) * m2).astype(np.uint8)
                assert np.array_equal(got[i, f], want), (ks, bt, i, f)
    assert (got != img0).any()


@pytest.mark.gpu
def test_edge_arguments_are_validated(env):
    E = env["E"]
    from calcium_b200 import _lib
    specs = E.make_specs(1, size="xsmall")
    ens = E.Ensemble(specs, T=200, frame_steps=[0, 199], device=env["dev"], output_size=(64, 64))
    ens.simulate()
    with pytest.raises(_lib.Cab200Error, match="kernel size"):
        ens.rasterize(image=True, instance=False, edge_blur=True, blur_kernel_size=11)
    with pytest.raises(_lib.Cab200Error, match="kernel size"):
        ens.rasterize(image=True, instance=False, edge_blur=True, blur_kernel_size=0)


@pytest.mark.gpu
def test_batch_path_applies_edge_blur(env, tmp_path):
    """generate_simulation_batch(edge_blur=True) (reference main.py:281-287): the stored PNGs hold the blurred
    frames -- 10-bit-in-16-bit gray = 4 x K2b's gray."""
    cv2 = pytest.importorskip("cv2")
    from calcium_b200 import generate_simulation_batch
    res = generate_simulation_batch(2, str(tmp_path), pouch_sizes=["xsmall"], sim_types=["Intercellular waves"],
                                    time_steps=[0, 9000], edge_blur=True, blur_kernel_size=5, blur_type="motion",
                                    image_size="160x120", generate_masks=True, generate_prompts=False,
                                    defect_configs=[{}] * 2)
    assert len(res["simulations"]) == 2 and all(s["image_count"] == 2 for s in res["simulations"])
    E = env["E"]
    specs = [E.PouchSpec.draw(res["simulations"][i]["parameters"], size="xsmall", sim_number=i) for i in range(2)]
    ens = E.Ensemble(specs, frame_steps=[0, 9000], device=env["dev"], output_size=(160, 120))
    ens.simulate()
    res_k2 = ens.rasterize(image=True, instance=True, edge_blur=True, blur_kernel_size=5, blur_type="motion")
    want = res_k2["image"].cpu().numpy()
    want_ins = res_k2["instance"].cpu().numpy().view(np.uint16)
    n_act = res_k2["n_active"].cpu().numpy()
    mdir = os.path.join(str(tmp_path), "masks", "train")
    mfiles = sorted(os.listdir(mdir))
    assert len(mfiles) == int((n_act > 0).sum()) == sum(s["mask_count"] for s in res["simulations"])
    for i in range(2):
        for f, t in enumerate((0, 9000)):
            hit = [x for x in mfiles if f"_{i}_t{t:05d}_" in x]
            assert len(hit) == int(n_act[i, f] > 0)
            if hit:            # the masks are not blurred: finished .npz files from K3's mask stream
                assert np.array_equal(np.load(os.path.join(mdir, hit[0]))["mask"], want_ins[i, f])
    files = sorted(os.listdir(os.path.join(str(tmp_path), "images", "train")))
    assert len(files) == 4
    for i in range(2):
        for f, t in enumerate((0, 9000)):
            name = [x for x in files if f"_{i}_t{t:05d}_" in x][0]
            png = cv2.imread(os.path.join(str(tmp_path), "images", "train", name), cv2.IMREAD_UNCHANGED)
            assert png.dtype == np.uint16 and np.array_equal(png[..., 0], want[i, f, :, :, 0].astype(np.uint16) * 4)
