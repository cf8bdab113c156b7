"""Parity tests proper: the CUDA path (through the C ABI, via calcium_b200.ensemble / Pouch) against
the oracle and against fixtures recorded from the reference.  Run with `-m gpu` on a B200.

Bars (north_star): trajectories within rtol 1e-9 of the reference's numba build (stated in RTOL
below); against the oracle -- the reference's arithmetic in source order, every operation rounded
once -- the FP64 mode is required to be BIT-IDENTICAL, which is the stronger statement the kernel is
designed for.  Images and masks: bit-exact.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TYPES = ["Single cell spikes", "Intercellular transients", "Intercellular waves", "Fluttering"]
RTOL = 1e-9          # tolerance of BASELINE.json's north_star for FP64 trajectories vs the reference


@pytest.fixture(scope="module")
def env(cuda_device, oracle):
    import torch
    from calcium_b200 import _lib, build
    build.build()
    from calcium_b200 import ensemble
    return dict(dev=cuda_device, O=oracle, lib=_lib.load(), L=_lib, E=ensemble, torch=torch)


def _oracle_frames(O, spec, T, steps, f32=False):
    g = spec.geometry
    c0 = np.zeros(g.n_cells) if spec.c0 is None else spec.c0
    p0 = np.zeros(g.n_cells) if spec.p0 is None else spec.p0
    return O.simulate_c(g.csr, O.pack_params(spec.params), c0, p0, spec.r0, spec.vplc, T, steps, f32=f32)


def _max_rel(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max())


# ---- the division K1 relies on --------------------------------------------------------------------

def test_branch_free_division_is_ieee(env):
    mm = ctypes.c_ulonglong(1)
    for er in (4, 40, 300):
        env["L"].check(env["lib"].cab200_selftest_division(1 << 28, er, 1234 + er, ctypes.byref(mm), None), "selftest")
        assert mm.value == 0, f"{mm.value} mismatches vs __ddiv_rn at exponent range +-{er}"


# ---- K1 against the oracle: bit-exact -------------------------------------------------------------

@pytest.mark.parametrize("size,count,T", [("xsmall", 8, 18000), ("small", 4, 18000), ("medium", 2, 18000),
                                          ("large", 1, 3000)])
def test_k1_bit_exact_vs_oracle(env, size, count, T):
    E, O = env["E"], env["O"]
    specs = E.make_specs(count, size=size, first_seed=17)       # types round-robin, randomised params
    steps = sorted(set(list(range(0, T, 200)) + [1, T - 1]))
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"])
    ens.simulate(); ens.synchronize()
    assert not ens.check_status().any()
    for i, s in enumerate(specs):
        want = _oracle_frames(O, s, T, ens.frame_steps)
        got = ens.frames_of(i).cpu().numpy()
        assert np.array_equal(want, got), f"{size}[{i}] max abs diff {np.abs(want - got).max():.3e}"


def test_k1_against_reference_fixtures(env, golden_dir):
    """GPU vs what the reference itself produced (numba fastmath build) -- rtol 1e-9."""
    E = env["E"]
    from calcium_b200.parameters import SimulationParameters as SP
    from calcium_b200.geometry import get_geometry
    z = np.load(os.path.join(golden_dir, "traj_xsmall_waves.npz"))
    spec = E.PouchSpec.draw(SP("Intercellular waves").get_params_dict(), "xsmall", 0)
    ens = E.Ensemble([spec], frame_steps=z["steps"].tolist(), device=env["dev"])
    ens.simulate()
    got = ens.frames_of(0).cpu().numpy()
    assert _max_rel(got, z["frames_jit"]) < RTOL and _max_rel(got, z["frames_pyfunc"]) < RTOL
    z = np.load(os.path.join(golden_dir, "traj_small_types.npz"))
    specs = [E.PouchSpec.draw(SP(st).generate_random_params(seed=i), "small", i) for i, st in enumerate(TYPES)]
    ens = E.Ensemble(specs, frame_steps=z["steps"].tolist(), device=env["dev"])
    ens.simulate()
    for i in range(4):
        assert _max_rel(ens.frames_of(i).cpu().numpy(), z[f"frames_{i}"]) < RTOL, TYPES[i]
    z = np.load(os.path.join(golden_dir, "traj_medium_flutter.npz"))
    spec = E.PouchSpec.draw(SP("Fluttering").generate_random_params(seed=3), "medium", 3)
    ens = E.Ensemble([spec], frame_steps=z["steps"].tolist(), device=env["dev"])
    ens.simulate()
    assert _max_rel(ens.frames_of(0).cpu().numpy(), z["frames"]) < RTOL


def test_layout_choices_never_change_results(env):
    """Slot plan on/off and every CTA shape give the same bits (layout is not arithmetic)."""
    E, lib, L = env["E"], env["lib"], env["L"]
    specs = E.make_specs(3, size="medium", first_seed=40)
    T, steps = 3000, [0, 1500, 2999]
    ref = None
    try:
        for shape in [(0, 0), (512, 3), (768, 2), (1024, 2)]:
            L.check(lib.cab200_sim_set_plan(100000, *shape), "set_plan")
            for use_plan in (False, True):
                ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"], use_plan=use_plan)
                f = ens.simulate().cpu().numpy().copy()
                assert not ens.check_status().any()
                if ref is None:
                    ref = f
                assert np.array_equal(ref, f), (shape, use_plan)
    finally:
        lib.cab200_sim_set_plan(0, 0, 0)
    want = _oracle_frames(env["O"], specs[1], T, steps)
    assert np.array_equal(want, ref.reshape(-1)[1500 * 4 * 3: 2 * 1500 * 4 * 3].reshape(3, 4, 1500))


@pytest.mark.parametrize("size,n_cta,count,T", [("xsmall", 2, 3, 4000), ("small", 2, 3, 4000), ("medium", 2, 3, 6000),
                                                  ("medium", 4, 5, 6000), ("large", 4, 2, 5000),
                                                  ("large", 3, 2, 3000), ("large", 8, 2, 3000), ("xlarge", 8, 2, 3000)])
def test_cluster_path_bit_exact(env, size, n_cta, count, T):
    """A pouch split over a thread-block cluster (halo entries pushed through distributed shared
    memory, cluster-wide step barrier) produces the same bits as the oracle."""
    E, O = env["E"], env["O"]
    specs = E.make_specs(count, size=size, first_seed=60)
    steps = [0, 1, 2, T // 2, T - 1]
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"], cluster=n_cta)
    assert ens.g_cluster.tolist() == [n_cta]
    ens.simulate(); ens.synchronize()
    assert not ens.check_status().any()
    for i, s in enumerate(specs):
        want = _oracle_frames(O, s, T, steps)
        got = ens.frames_of(i).cpu().numpy()
        assert np.array_equal(want, got), f"{size} x{n_cta} [{i}] max abs diff {np.abs(want - got).max():.3e}"


def test_xlarge_takes_a_cluster_of_eight(env):
    """`xlarge` (core/geometry_loader.py:36; 6500 synthetic cells) exceeds 4 x 1536 cells: cluster="auto" runs it
    on 8 CTAs, next to a large pouch on 4, in one call; a pouch beyond 8 x 1536 cells is refused with the limit."""
    E, O = env["E"], env["O"]
    specs = E.make_specs(1, size="xlarge", first_seed=3) + E.make_specs(1, size="large", first_seed=4)
    T, steps = 2500, [0, 1250, 2499]
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"])
    assert sorted(ens.g_cluster.tolist()) == [4, 8]
    ens.simulate()
    assert not ens.check_status().any()
    for i, s in enumerate(specs):
        assert np.array_equal(_oracle_frames(O, s, T, steps), ens.frames_of(i).cpu().numpy()), i
    lay = np.zeros(4, dtype=np.int64)
    lib = env["lib"]
    assert lib.cab200_cluster_layout(8 * 1536 + 1, 8, 64, env["L"].as_i64p(lay)) != 0
    assert b"1536" in lib.cab200_last_error()


def test_cluster_auto_for_large_and_mixed_ensembles(env):
    """cluster="auto": large pouches take 4-CTA clusters, a handful of medium pouches (too few to fill the
    SMs one CTA each) too, small ones stay on one CTA -- all in one call."""
    E, O = env["E"], env["O"]
    specs = E.make_specs(2, size="large", first_seed=5) + E.make_specs(3, size="medium", first_seed=9) + \
        E.make_specs(2, size="xsmall", first_seed=1)
    T, steps = 3000, [0, 1500, 2999]
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"])
    assert sorted(ens.g_cluster.tolist()) == [1, 4, 4]
    ens.simulate()
    assert not ens.check_status().any()
    for i, s in enumerate(specs):
        assert np.array_equal(_oracle_frames(O, s, T, steps), ens.frames_of(i).cpu().numpy()), i


def _weighted_geometry(n=300, seed=5):
    from calcium_b200.geometry import Geometry, make_synthetic_geometry
    g = make_synthetic_geometry("w", n_cells=n, seed=seed)
    rng = np.random.RandomState(seed)
    w = np.zeros((n, n))
    vals = rng.uniform(0.2, 1.7, len(g.edges))
    w[g.edges[:, 0], g.edges[:, 1]] = vals
    w[g.edges[:, 1], g.edges[:, 0]] = vals
    lap = w - np.diag(w.sum(axis=1))
    return Geometry(size="w", vertices=g.vertices, edges=g.edges, weights=lap)


def test_weighted_laplacian_path(env):
    """General CSR weights (not adjacency - degree) take the multiply-per-entry kernel; still bit-exact."""
    E, O = env["E"], env["O"]
    from calcium_b200.parameters import SimulationParameters as SP
    g = _weighted_geometry()
    specs = []
    for i in range(3):
        prm = SP(TYPES[i + 1]).generate_random_params(seed=i)
        v = E.draw_vplc(prm, g.n_cells, i)
        r0, _ = E.draw_initial_state(prm, g.n_cells, i, v)
        specs.append(E.PouchSpec(params=prm, geometry=g, sim_number=i, vplc=v.flatten(), r0=r0))
    ens = E.Ensemble(specs, T=2500, frame_steps=[0, 1200, 2499], device=env["dev"])
    assert ens.g_unit[0] == 0
    ens.simulate()
    assert not ens.check_status().any()
    for i, s in enumerate(specs):
        assert np.array_equal(_oracle_frames(O, s, 2500, [0, 1200, 2499]), ens.frames_of(i).cpu().numpy())


def test_false_unit_declaration_is_flagged(env):
    E = env["E"]
    g = _weighted_geometry(n=64, seed=9)
    from calcium_b200.parameters import SimulationParameters as SP
    prm = SP("Fluttering").get_params_dict()
    v = E.draw_vplc(prm, g.n_cells, 0)
    r0, _ = E.draw_initial_state(prm, g.n_cells, 0, v)
    ens = E.Ensemble([E.PouchSpec(params=prm, geometry=g, vplc=v.flatten(), r0=r0)], T=10, frame_steps=[9],
                     device=env["dev"])
    ens.g_unit[:] = 1                      # lie: declare unit off-diagonals (with a plan: without one the general
                                           # kernel runs whatever g_unit says)
    ens.simulate()
    with pytest.raises(env["L"].Cab200Error, match="rejected the geometry"):
        ens.check_status()


def test_tiny_and_degenerate_graphs(env):
    """n = 1 (no entries at all), a 2-cell path, an isolated cell, n not a multiple of 32."""
    E, O = env["E"], env["O"]
    from calcium_b200.geometry import Geometry
    from calcium_b200.parameters import SimulationParameters as SP
    prm = SP("Fluttering").get_params_dict()

    def geo(n, edges):
        verts = np.empty(n, dtype=object)
        for i in range(n):
            verts[i] = np.array([[i, 0.0], [i + 1, 0.0], [i + 1, 1.0], [i, 1.0]])
        return Geometry(size=f"g{n}", vertices=verts, edges=np.array(edges, dtype=np.int32).reshape(-1, 2))

    for g in (geo(1, []), geo(2, [(0, 1)]), geo(3, [(0, 1)]), geo(37, [(i, i + 1) for i in range(36)])):
        n = g.n_cells
        rng = np.random.RandomState(n)
        spec = E.PouchSpec(params=prm, geometry=g, vplc=rng.uniform(1.3, 1.5, n), r0=rng.uniform(.5, .7, n),
                           c0=rng.uniform(0, .2, n), p0=rng.uniform(0, .3, n))
        ens = E.Ensemble([spec], T=800, frame_steps=[0, 799], device=env["dev"])
        ens.simulate()
        assert not ens.check_status().any()
        assert np.array_equal(_oracle_frames(O, spec, 800, [0, 799]), ens.frames_of(0).cpu().numpy()), n


def test_horizon_and_capture_edge_cases(env):
    E, O = env["E"], env["O"]
    specs = E.make_specs(2, size="xsmall", first_seed=3)
    # T = 1: only the initial state exists
    ens = E.Ensemble(specs, T=1, frame_steps=[0], device=env["dev"])
    ens.simulate()
    f = ens.frames_of(1).cpu().numpy()
    assert np.all(f[0, 0] == 0) and np.array_equal(f[0, 3], specs[1].r0)
    assert np.all(f[0, 2] == specs[1].params["c_tot"] / specs[1].params["beta"])
    # no frames at all: runs, writes nothing
    ens = E.Ensemble(specs, T=50, frame_steps=[], device=env["dev"])
    ens.simulate(); ens.synchronize()
    assert ens.F == 0 and not ens.check_status().any()
    # every step captured (config 4's mode, short horizon)
    T = 300
    ens = E.Ensemble(specs, T=T, frame_steps=range(T), device=env["dev"])
    ens.simulate()
    assert np.array_equal(_oracle_frames(O, specs[0], T, list(range(T))), ens.frames_of(0).cpu().numpy())
    # out-of-range steps are dropped like the reference's `if time_step >= pouch.T: continue` (main.py:270)
    ens = E.Ensemble(specs, T=100, frame_steps=[0, 50, 100, 18000], device=env["dev"])
    assert ens.frame_steps.tolist() == [0, 50]


def test_non_finite_status(env):
    E = env["E"]
    spec = E.make_specs(1, size="xsmall")[0]
    spec.params = dict(spec.params, D_p=1e6)       # explicit Euler blows up
    ens = E.Ensemble([spec], T=400, frame_steps=[0, 399], device=env["dev"])
    ens.simulate()
    st = ens.check_status()
    assert st[0] & env["L"].ST_NONFINITE


def test_ragged_ensemble_more_pouches_than_sms(env):
    """Mixed sizes in one call, more CTAs than SMs; every 23rd pouch checked against the oracle."""
    E, O = env["E"], env["O"]
    sizes = ["xsmall", "small", "xsmall", "medium", "xsmall", "small"]
    specs = []
    for i in range(330):
        specs.append(E.make_specs(1, size=sizes[i % len(sizes)], first_seed=i, sim_types=[TYPES[i % 4]])[0])
    T, steps = 1200, [0, 600, 1199]
    ens = E.Ensemble(specs, T=T, frame_steps=steps, device=env["dev"])
    ens.simulate()
    assert not ens.check_status().any()
    for i in range(0, 330, 23):
        assert np.array_equal(_oracle_frames(O, specs[i], T, steps), ens.frames_of(i).cpu().numpy()), i


def test_fp32_mode_tracks_fp64_and_reports_error(env):
    """The optional FP32 mode (fast arithmetic, not strict): float32 frames that follow the FP64 trajectory to
    float precision over a short horizon, stay finite and physical over the full one, and whose deviation from
    FP64 -- O(1) in Ca once spike timing has drifted, by the nature of the system -- is a reported number, not a
    parity claim (profiles/, scripts/run_configs.py)."""
    E, L = env["E"], env["L"]
    specs = E.make_specs(4, size="small", first_seed=0)
    short = [0, 1, 10, 100]
    a32 = E.Ensemble(specs, T=101, frame_steps=short, device=env["dev"], precision=L.FP32)
    a64 = E.Ensemble(specs, T=101, frame_steps=short, device=env["dev"])
    a32.simulate(); a64.simulate()
    for i in range(len(specs)):
        got = a32.frames_of(i).cpu().numpy()
        assert got.dtype == np.float32
        want = a64.frames_of(i).cpu().numpy()
        assert np.allclose(got, want, rtol=2e-4, atol=1e-6), (i, np.abs(got - want).max())
    steps = [0, 6000, 17999]
    e32 = E.Ensemble(specs, frame_steps=steps, device=env["dev"], precision=L.FP32)
    e32.simulate()
    assert not (e32.check_status() & L.ST_NONFINITE).any()
    for i, s in enumerate(specs):
        got = e32.frames_of(i).cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all() and got[:, 0].min() >= 0 and got[:, 0].max() < 5.0
        # the ER store follows from c by the same identity as in FP64 (core/pouch.py:95), to float precision
        assert np.allclose(got[:, 2], (s.params["c_tot"] - got[:, 0]) / s.params["beta"], rtol=1e-5, atol=1e-6)
    # the FP32 mode has its own slot plans (8-byte entries: cab200_plan_slots_entry): which cell sits in which slot and
    # which copy an entry reads is not arithmetic -- the plan made for 16-byte entries gives the same float32 bits.
    # (Without any plan the rows take the legacy layout, whose diagonal term is a rounded product read from shared
    # memory instead of an FFMA from registers: last-bit differences in this non-strict mode, 6e-6 here.)
    from calcium_b200.ensemble import slot_plan
    import torch
    msp = E.make_specs(2, size="medium", first_seed=7)
    runs = []
    for variant in ("own", "fp64-plan", "none"):
        ens = E.Ensemble(msp, T=2500, frame_steps=[0, 1200, 2499], device=env["dev"], precision=L.FP32,
                         use_plan=variant != "none", cluster=1)
        if variant == "fp64-plan":
            ens.d_plan = torch.from_numpy(slot_plan(msp[0].geometry, True, L.FP64_STRICT).view(np.int16)).to(env["dev"])
        runs.append(ens.simulate().cpu().numpy().copy())
        assert not ens.check_status().any()
    assert np.array_equal(runs[0], runs[1])
    assert np.allclose(runs[0], runs[2], rtol=1e-3, atol=1e-4)


# ---- K0 / K2 ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("size,out", [("xsmall", (512, 512)), ("xsmall", (160, 120)), ("small", (640, 480)),
                                      ("medium", (512, 512)), ("large", (512, 512)), ("medium", (1024, 1024))])
def test_label_maps_bit_exact(env, size, out):
    E, O = env["E"], env["O"]
    from calcium_b200.geometry import get_geometry
    g = get_geometry(size)
    limg, lmsk = E.build_label_maps(g, out, env["dev"])
    assert np.array_equal(E._u16_numpy(lmsk), O.cells_mask_oracle(g.vertices, out).astype(np.uint16))
    assert np.array_equal(E._u16_numpy(limg), O.paint_label_map_cv2(g.vertices, out).astype(np.uint16))


def test_paint_label_kernel_matches_cv2_on_random_polygons(env):
    """K0b against the real cv2.fillPoly (paint order, later wins) on overlapping random polygons,
    convex and concave, including degenerate ones (repeated / collinear vertices, zero height)."""
    import cv2
    torch, L, lib = env["torch"], env["L"], env["lib"]
    rng = np.random.RandomState(3)
    H, W = 200, 260
    polys = []
    for i in range(500):
        k = rng.randint(3, 10)
        cx, cy = rng.randint(0, W), rng.randint(0, H)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k)); rad = rng.uniform(1, 30, k)
        p = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)
        p = np.clip(p, 0, [W - 1, H - 1]).astype(np.int32)
        if i % 37 == 0:
            p[1] = p[0]                      # repeated vertex
        if i % 41 == 0:
            p[:, 1] = p[0, 1]                # all on one row
        polys.append(p)
    offs = np.cumsum([0] + [len(p) for p in polys]).astype(np.int32)
    pts = np.concatenate(polys).astype(np.int32)
    want = np.zeros((H, W), dtype=np.int32)
    for i, p in enumerate(polys):
        cv2.fillPoly(want, [p.reshape(-1, 1, 2)], color=int(i + 1))
    d_offs, d_pts = torch.from_numpy(offs).to(env["dev"]), torch.from_numpy(pts).to(env["dev"])
    out = torch.empty((H, W), dtype=torch.int16, device=env["dev"])
    L.check(lib.cab200_paint_label_map(len(polys), d_offs.data_ptr(), d_pts.data_ptr(), H, W, out.data_ptr(), None),
            "cab200_paint_label_map")
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint16)
    assert np.array_equal(got, want.astype(np.uint16)), int((got != want).sum())


@pytest.mark.parametrize("size,out,quirk,thr", [("xsmall", (512, 512), True, 0.1), ("medium", (512, 512), True, 0.1),
                                                ("small", (160, 120), False, 0.1), ("medium", (640, 480), True, 0.35),
                                                ("xsmall", (250, 130), True, 0.1)])
def test_k2_bit_exact_vs_oracle(env, size, out, quirk, thr):
    E, O = env["E"], env["O"]
    specs = E.make_specs(4, size=size, first_seed=100)      # includes a Fluttering pouch: many active cells
    steps = list(range(0, 18000, 1500)) + [17999]
    ens = E.Ensemble(specs, frame_steps=steps, device=env["dev"], output_size=out)
    ens.simulate()
    res = ens.rasterize(image=True, instance=True, active=True, threshold=thr, quirk=quirk)
    ens.synchronize()
    limg, lmsk = [E._u16_numpy(t)[0] for t in ens.label_maps()]
    img = res["image"].cpu().numpy(); ins = res["instance"].cpu().numpy().view(np.uint16)
    act = res["active"].cpu().numpy(); nact = res["n_active"].cpu().numpy()
    seen_active = 0
    for i in range(len(specs)):
        fr = ens.frames_of(i).cpu().numpy()
        for f in range(ens.F):
            oi, oa, oin, on = O.raster_frame_c(fr[f, 0], limg, lmsk, thr, quirk)
            assert np.array_equal(oi, img[i, f]) and np.array_equal(oa, act[i, f])
            assert np.array_equal(oin, ins[i, f]) and on == nact[i, f]
            seen_active = max(seen_active, on)
    assert seen_active > 20, "the fixture ensemble should exercise frames with many active cells"
    # sub-range rasterisation writes the same bytes
    part = ens.rasterize(image=True, instance=True, pouches=(1, 3), threshold=thr, quirk=quirk)
    assert np.array_equal(part["image"].cpu().numpy(), img[1:3])
    assert np.array_equal(part["instance"].cpu().numpy().view(np.uint16), ins[1:3])


def test_k2_literal_reference_loops(env):
    """One frame against the reference's literal algorithm: a fillPoly per cell, a full-image pass
    per active cell, a relabel pass per active cell."""
    E, O = env["E"], env["O"]
    spec = E.make_specs(4, size="small", first_seed=0)[3]       # Fluttering
    ens = E.Ensemble([spec], frame_steps=[9000], device=env["dev"])
    ens.simulate()
    res = ens.rasterize(image=True, instance=True, active=True)
    ca = ens.frames_of(0).cpu().numpy()[0, 0]
    g = spec.geometry
    assert np.array_equal(O.generate_image_oracle(g.vertices, ca, (512, 512)), res["image"].cpu().numpy()[0, 0])
    active = O.active_cells_oracle(ca)
    cm = O.cells_mask_oracle(g.vertices, (512, 512))
    am = O.active_mask_oracle(cm, active)
    assert np.array_equal(am, res["active"].cpu().numpy()[0, 0])
    assert np.array_equal(O.instance_mask_oracle(am, active), res["instance"].cpu().numpy().view(np.uint16)[0, 0])
    assert len(active) == int(res["n_active"].cpu().numpy()[0, 0]) > 0


# ---- the Pouch drop-in against fixtures recorded from the reference's own methods ------------------

def test_pouch_api_against_reference_fixtures(env, golden_dir):
    from calcium_b200 import Pouch, SimulationParameters, create_instance_mask_from_cells
    z = np.load(os.path.join(golden_dir, "raster_xsmall.npz"))
    traj = np.load(os.path.join(golden_dir, "traj_xsmall_waves.npz"))
    p = SimulationParameters("Intercellular waves").get_params_dict()
    for tag in ("160x120", "512x512"):
        w, h = map(int, tag.split("x"))
        pouch = Pouch(params=p, size="xsmall", sim_number=0, output_size=(w, h))
        assert pouch.simulate() is True
        assert np.array_equal(pouch.V_PLC.flatten(), traj["vplc"])           # initiators applied in place
        assert pouch.cells_mask.dtype == np.int32
        assert np.array_equal(pouch.cells_mask.astype(np.uint16), z[f"cells_mask_{tag}"])
        for t in z[f"t_{tag}"].tolist():
            img = pouch.generate_image(t)
            assert img.shape == (h, w, 2) and img.dtype == np.uint8
            assert np.array_equal(img, z[f"image_{tag}_{t}"]), (tag, t)
            act = pouch.get_active_cells(t)
            assert act == z[f"active_{tag}_{t}"].tolist()
            am = pouch.get_cell_masks(active_only=True, time_step=t)
            assert am.dtype == np.int32 and np.array_equal(am.astype(np.uint16), z[f"activemask_{tag}_{t}"])
            inst = create_instance_mask_from_cells(am, act)
            assert np.array_equal(inst, z[f"instance_{tag}_{t}"])
            assert np.array_equal(pouch.get_instance_mask(t), z[f"instance_{tag}_{t}"])
        if tag == "160x120":
            assert np.array_equal(pouch.generate_image(10000, edge_blur=True), z["edge_mean3_160x120_10000"])
            assert np.array_equal(pouch.generate_image(10000, edge_blur=True, blur_kernel_size=5, blur_type="motion"),
                                  z["edge_motion5_160x120_10000"])
            assert np.array_equal(pouch.generate_image(10000, with_border=True, edge_blur=True),
                                  z["border_blur_160x120_10000"])
            assert pouch.generate_image(10000, with_border=True).shape == (h, w, 2)
    # disc_dynamics indexing like the reference's callers (labeling.py:166, pouch.py:418,563,584)
    steps = traj["steps"].tolist()
    dd = pouch.disc_dynamics
    assert dd.shape == (200, 4, 18000)
    got = np.transpose(dd[:, :, steps], (2, 1, 0))
    assert _max_rel(got, traj["frames_jit"]) < RTOL
    assert float(dd[7, 0, 10000]) == got[steps.index(10000), 0, 7]
    assert np.array_equal(dd[:, 0, 10000], got[steps.index(10000), 0])
    # a step outside the default sampling is produced on demand (deterministic re-integration)
    assert 4321 not in pouch._steps
    ca = dd[:, 0, 4321]
    O = env["O"]
    spec = env["E"].PouchSpec.draw(p, "xsmall", 0)
    assert np.array_equal(ca, _oracle_frames(O, spec, 18000, [4321])[0, 0])
    assert np.array_equal(dd[:, 0, 10000], got[steps.index(10000), 0])       # earlier frames unchanged
    lbl = pouch.generate_label_data(10000)
    assert lbl["n_active"] == len(lbl["active_cells"]) == len(lbl["calcium_values"])
    assert lbl["simulation_info"] == {"sim_number": 0, "size": "xsmall", "time_step": 10000, "total_cells": 200}
    with pytest.raises(RuntimeError, match="Error generating image"):
        Pouch(params=p, size="xsmall", output_size=(3, 3)).generate_image(0)   # odd pixel count unsupported


# ---- BASELINE-size properties (config 3's per-GPU shard) --------------------------------------------

def test_full_size_medium_wave_properties(env):
    """148 medium pouches (one wave on 148 SMs), full T, all 90 frames + masks: properties that
    hold at any size, plus bit-exact spot checks of three pouches against the oracle."""
    E, O, torch = env["E"], env["O"], env["torch"]
    specs = E.make_specs(148, size="medium", first_seed=0)
    ens = E.Ensemble(specs, device=env["dev"])
    ens.simulate(); ens.synchronize()
    assert not ens.check_status().any()
    frames = ens.frames.view(148, ens.F, 4, 1500)
    first = frames.clone()
    # conservation: s == (c_tot - c)/beta at every stored step, exactly (core/pouch.py:95)
    prm = torch.from_numpy(np.stack([E.pack_params(s.params) for s in specs])).to(env["dev"])
    c_tot, beta = prm[:, 8].view(-1, 1, 1), prm[:, 9].view(-1, 1, 1)
    assert torch.equal(frames[:, :, 2], (c_tot - frames[:, :, 0]) / beta)
    assert torch.isfinite(frames).all()
    # t = 0 frame is the initial state
    assert torch.count_nonzero(frames[:, 0, 0]) == 0 and torch.count_nonzero(frames[:, 0, 1]) == 0
    # determinism: a second run gives the same bits
    ens.simulate(); ens.synchronize()
    assert torch.equal(first, ens.frames.view(148, ens.F, 4, 1500))
    # sharding: two half-ensembles (what two ranks would run) equal the single run
    for lo, hi in ((0, 74), (74, 148)):
        half = E.Ensemble(specs[lo:hi], device=env["dev"])
        half.simulate()
        assert torch.equal(half.frames.view(hi - lo, ens.F, 4, 1500), first[lo:hi])
    for i in (5, 78, 147):
        assert np.array_equal(_oracle_frames(O, specs[i], 18000, ens.frame_steps), first[i].cpu().numpy()), i
    # rasterise in two chunks; image alpha is the static footprint, gray is 0 on background,
    # instance ids never exceed the active count, and a frame without active cells is all zero
    limg, lmsk = ens.label_maps()
    foot = (limg[0] != 0)
    for lo, hi in ((0, 74), (74, 148)):
        res = ens.rasterize(pouches=(lo, hi))
        img, ins, nact = res["image"], res["instance"], res["n_active"]
        assert torch.equal(img[..., 1] == 255, foot.expand_as(img[..., 1]))
        assert torch.count_nonzero(img[..., 0][~foot.expand_as(img[..., 0])]) == 0
        mx = ins.view(hi - lo, ens.F, -1).to(torch.int32).amax(dim=2)
        assert torch.all(mx <= nact)
        assert torch.count_nonzero(mx[nact == 0]) == 0
        active = (frames[lo:hi, :, 0] > 0.1).sum(dim=2).to(torch.int32)
        assert torch.equal(active, nact)
        del res
    # one full frame against the oracle raster
    fr = first[147, 60].cpu().numpy()
    res = ens.rasterize(pouches=(147, 148), active=True)
    oi, oa, oin, on = O.raster_frame_c(fr[0], E._u16_numpy(limg)[0], E._u16_numpy(lmsk)[0], 0.1, True)
    assert np.array_equal(oi, res["image"][0, 60].cpu().numpy())
    assert np.array_equal(oin, res["instance"][0, 60].cpu().numpy().view(np.uint16))
    assert np.array_equal(oa, res["active"][0, 60].cpu().numpy())


def test_end_to_end_host_path_matches_device_results(env):
    """Ensemble.run_to_host (pinned host in, pinned host out, chunked, copy stream) delivers exactly
    the bytes rasterize() produces, ragged tail chunk included."""
    E = env["E"]
    specs = E.make_specs(7, size="xsmall", first_seed=200)
    ens = E.Ensemble(specs, frame_steps=[0, 4000, 9000, 17999], device=env["dev"], output_size=(256, 192))
    got = {}

    def consumer(lo, hi, arrays):
        for i in range(lo, hi):
            got[i] = {k: v[i - lo].copy() for k, v in arrays.items()}

    moved = ens.run_to_host(chunk_pouches=3, consumer=consumer)
    ref = ens.rasterize(image=True, instance=True)
    ens.synchronize()
    assert sorted(got) == list(range(7))
    for i in range(7):
        assert np.array_equal(got[i]["image"], ref["image"][i].cpu().numpy())
        assert np.array_equal(got[i]["instance"], ref["instance"][i].cpu().numpy())
        assert np.array_equal(got[i]["n_active"], ref["n_active"][i].cpu().numpy())
    assert np.array_equal(ens._h_frames.numpy(), ens.frames.cpu().numpy())
    assert moved["d2h_bytes"] == ens.frames.numel() * 8 + 7 * 4 * 256 * 192 * 4 + 7 * 4 * 4
    assert moved["h2d_bytes"] == (7 * 16 + 4 * 7 * 200) * 8


# ---- batch entry points ------------------------------------------------------------------------------

def test_generate_simulation_batch_and_controller(env, tmp_path):
    import random
    from calcium_b200 import BatchGenerationController, generate_simulation_batch
    O = env["O"]
    random.seed(5)
    seen = []
    res = generate_simulation_batch(5, str(tmp_path / "b"), pouch_sizes=["xsmall", "small"],
                                    time_steps=[0, 5000, 9000, 17800, 99999],
                                    progress_callback=lambda a, b: seen.append((a, b)), keep_arrays=True,
                                    defect_configs=[{}] * 5)          # clean frames (default: random defects per frame)
    assert seen == [(i + 1, 5) for i in range(5)]
    assert res["num_simulations"] == 5 and len(res["simulations"]) == 5 and res["batch_index"] == 0
    sim = res["simulations"][2]
    assert set(sim) >= {"simulation_id", "simulation_name", "simulation_type", "pouch_size", "parameters",
                        "image_count", "image_dir", "mask_dir", "pouch"}
    assert sim["image_count"] == 4 and sim["pouch"] is None
    assert sim["simulation_name"] == f"{sim['simulation_type'].replace(' ', '_')}_2"
    imgs = os.listdir(tmp_path / "b" / "images" / "train")
    assert len(imgs) == 20 and os.path.exists(tmp_path / "b" / "simulation_results.json")
    # arrays equal the oracle for one simulation (parameters drawn with seed = sim index)
    from calcium_b200.parameters import SimulationParameters as SP
    from calcium_b200.ensemble import PouchSpec
    a = res["arrays"][2]
    spec = PouchSpec.draw(SP(sim["simulation_type"]).generate_random_params(seed=2), sim["pouch_size"], 2)
    assert sim["parameters"] == spec.params
    assert np.array_equal(_oracle_frames(O, spec, 18000, a["time_steps"]), a["frames"])
    masks = os.listdir(tmp_path / "b" / "masks" / "train")
    assert len(masks) == sum(int((s_["n_active"] > 0).sum()) for s_ in res["arrays"].values())
    if masks:
        m = np.load(tmp_path / "b" / "masks" / "train" / masks[0])["mask"]
        assert m.dtype == np.uint16 and m.shape == (512, 512)
        assert "image_mask_csv" in res

    msgs = []
    ctl = BatchGenerationController()
    out = ctl.run_batch_generation({"num_simulations": 3, "pouch_sizes": ["xsmall"], "time_steps": [0, 9000],
                                    "image_size": "160x120", "create_csv": True, "write_files": True,
                                    "defect_config": {"vignetting": False}},
                                   str(tmp_path / "c"), progress_callback=lambda *a: False,
                                   status_callback=msgs.append)
    assert out["success"] is True and len(out["simulations"]) == 3
    assert msgs[0] == "Initializing batch generation..." and msgs[-1].startswith("Batch complete: 3 simulations, 6 images")
    out = ctl.run_batch_generation({"enable_multi_batch": True, "num_batches": 2, "sims_per_batch": 2,
                                    "pouch_sizes": ["xsmall"], "time_steps": [100], "defect_config": {"vignetting": False}},
                                   str(tmp_path / "d"))
    assert out["success"] and len(out["simulations"]) == 4
    # two batches into one directory within the same second: no file of the first is overwritten by the second
    # (unique run ids), and the CSV covers both
    dfiles = os.listdir(tmp_path / "d" / "images" / "train")
    assert len(dfiles) == 4 and len({f.rsplit("_", 1)[-1] for f in dfiles}) == 2
    rows = open(out["image_mask_csv"]).read().strip().splitlines() if "image_mask_csv" in out else ["hdr"]
    assert len(rows) - 1 == len(os.listdir(tmp_path / "d" / "masks" / "train"))
    bad = ctl.run_batch_generation({"num_simulations": "x"}, str(tmp_path / "e"))
    assert bad["success"] is False and bad["error"].startswith("Batch generation error:")


# ---- coalesced capture (frames_tmp): every-step histories ---------------------------------------------

@pytest.mark.parametrize("size,count,cluster,prec", [("medium", 3, 1, 64), ("large", 2, 4, 64), ("small", 5, 2, 64),
                                                     ("medium", 2, 1, 32)])
def test_every_step_capture_through_slot_order_staging(env, size, count, cluster, prec):
    """Capturing every step (the reference's full disc_dynamics history, core/pouch.py:203) goes through
    coalesced slot-order stores + an un-permute kernel; the frames are the same bits as the direct path."""
    E = env["E"]
    T = 600
    specs = E.make_specs(count, size=size, first_seed=23)
    a = E.Ensemble(specs, T=T, frame_steps=range(T), device=env["dev"], cluster=cluster, precision=prec,
                   coalesced_capture=True)
    a.simulate(); a.synchronize()
    assert a._frames_tmp is not None
    b = E.Ensemble(specs, T=T, frame_steps=range(T), device=env["dev"], cluster=cluster, precision=prec,
                   coalesced_capture=False)
    b.simulate(); b.synchronize()
    assert b._frames_tmp is None
    assert not a.check_status().any()
    assert env["torch"].equal(a.frames, b.frames)
    if prec == 64:
        want = _oracle_frames(env["O"], specs[0], T, list(range(T)))
        assert np.array_equal(want, a.frames_of(0).cpu().numpy())


def test_config3_head_against_reference_fixture(env, golden_dir):
    """The bench workload's first 8 pouches (make_specs(8, "medium") = BASELINE config 3's enumeration) and a
    large pouch on the GPU against what the unmodified reference computed for them
    (tests/golden/traj_config3_head.npz): identical host draws, trajectories within the north-star tolerance,
    K2's active-cell count equal in every one of the 90 sampled frames."""
    E = env["E"]
    z = np.load(os.path.join(golden_dir, "traj_config3_head.npz"))
    specs = E.make_specs(8, size="medium")
    for i, s in enumerate(specs):
        assert np.array_equal(s.vplc, z[f"vplc_{i}"]) and np.array_equal(s.r0, z[f"r0_{i}"])
    sampled, steps = z["sampled"].tolist(), z["steps"].tolist()
    allsteps = sorted(set(sampled + steps))
    ens = E.Ensemble(specs, frame_steps=allsteps, device=env["dev"], output_size=(128, 128))
    ens.simulate()
    assert not ens.check_status().any()
    nact = ens.rasterize(image=False, instance=True)["n_active"].cpu().numpy()
    for i in range(8):
        fr = ens.frames_of(i).cpu().numpy()
        got = fr[[allsteps.index(t) for t in steps]]
        assert _max_rel(got, z[f"frames_{i}"]) < RTOL, (i, _max_rel(got, z[f"frames_{i}"]))
        assert np.array_equal(nact[i, [allsteps.index(t) for t in sampled]], z[f"n_active_{i}"]), i
    from calcium_b200.parameters import SimulationParameters as SP
    prm = SP("Intercellular waves").generate_random_params(seed=2)
    spec = E.PouchSpec.draw(prm, size="large", sim_number=2)
    assert np.array_equal(spec.vplc, z["large_vplc"]) and np.array_equal(spec.r0, z["large_r0"])
    ens = E.Ensemble([spec], frame_steps=z["large_steps"].tolist(), device=env["dev"])
    assert ens.g_cluster.tolist() == [4]                     # a large pouch runs as a 4-CTA cluster
    ens.simulate()
    assert _max_rel(ens.frames_of(0).cpu().numpy(), z["large_frames"]) < RTOL


@pytest.mark.parametrize("edge_blur", [False, True])
def test_batch_sharded_over_two_shards_of_one_gpu_equals_unsharded(env, tmp_path, edge_blur):
    """The ``devices=[...]`` drop-in sharding (LPT partition, one host thread per shard, results merged) on a
    one-GPU box: two shards on ``cuda:0`` run through exactly the code the multi-GPU call runs
    (``run_on_device`` per entry of ``devices``) -- same files, result dict and generator state as unsharded.  With
    ``edge_blur`` the images come from K3r, whose pinned landing buffer the two shards of the device share."""
    import random
    from calcium_b200 import generate_simulation_batch
    outs = []
    for k, devs in enumerate((None, [0, 0])):
        random.seed(13)
        d = tmp_path / f"run{k}"
        res = generate_simulation_batch(7, str(d), pouch_sizes=["xsmall", "small", "medium"], time_steps=[0, 6000, 12000],
                                        image_size="128x128", devices=devs, defect_configs=[{}] * 7, edge_blur=edge_blur)
        files = {}
        for sub in ("images", "masks", "prompts"):
            for name in sorted(os.listdir(d / sub / "train")):
                key = sub + "/" + "_".join(name.split("_")[:-1])          # drop the run id
                files[key] = open(d / sub / "train" / name, "rb").read()
        sims = [(s["simulation_id"], s["pouch_size"], s["image_count"], s["mask_count"], s["prompt_count"])
                for s in res["simulations"]]
        outs.append((files, sims, np.random.get_state()[1].copy(), res["stats"]))
    assert outs[0][1] == outs[1][1] and [s[0] for s in outs[1][1]] == list(range(7))
    assert outs[0][0].keys() == outs[1][0].keys() and len(outs[0][0]) > 0
    assert all(outs[0][0][k] == outs[1][0][k] for k in outs[0][0])
    assert np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]


def test_batch_sharded_over_two_gpus_equals_one(env, tmp_path):
    """generate_simulation_batch(devices=[0, 1]) (independent pouches per GPU, no
    collective): same files, same result dict and the same generator state as the one-GPU run."""
    import random
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from calcium_b200 import generate_simulation_batch
    outs = []
    for k, devs in enumerate((None, [0, 1])):
        random.seed(11)
        d = tmp_path / f"run{k}"
        res = generate_simulation_batch(7, str(d), pouch_sizes=["xsmall", "small", "medium"], time_steps=[0, 6000, 12000],
                                        image_size="128x128", devices=devs, defect_configs=[{}] * 7)
        files = {}
        for sub in ("images", "masks", "prompts"):
            for name in sorted(os.listdir(d / sub / "train")):
                key = sub + "/" + "_".join(name.split("_")[:-1])          # drop the run id
                files[key] = open(d / sub / "train" / name, "rb").read()
        sims = [(s["simulation_id"], s["pouch_size"], s["image_count"], s["mask_count"], s["prompt_count"])
                for s in res["simulations"]]
        outs.append((files, sims, np.random.get_state()[1].copy(), res["stats"]))
    assert outs[0][1] == outs[1][1] and [s[0] for s in outs[1][1]] == list(range(7))
    assert outs[0][0].keys() == outs[1][0].keys() and len(outs[0][0]) > 0
    assert all(outs[0][0][k] == outs[1][0][k] for k in outs[0][0])
    assert np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]


def test_segmented_run_continues_bit_exactly(env):
    """Ensemble.continue_from_last_frame: two segments of an every-step capture, the second started on the device from
    the first one's last frame, equal one long run bit for bit (how histories larger than HBM are produced, BASELINE
    config 4) -- mixed sizes, one of them a cluster."""
    E = env["E"]
    specs = E.make_specs(2, size="small", first_seed=11) + E.make_specs(1, size="large", first_seed=12) + \
        E.make_specs(1, size="small", first_seed=13)
    T1, T2 = 1201, 801
    whole = E.Ensemble(specs, T=T1 + T2 - 1, frame_steps=range(T1 + T2 - 1), device=env["dev"])
    whole.simulate(); whole.synchronize()
    seg = E.Ensemble(specs, T=T1, frame_steps=range(T1), device=env["dev"])
    seg.simulate()
    a = [seg.frames_of(i).cpu().numpy().copy() for i in range(len(specs))]
    seg.continue_from_last_frame()
    seg2 = E.Ensemble(specs, T=T2, frame_steps=range(T2), device=env["dev"])
    seg2.upload(); seg2._d_in.copy_(seg._d_in)
    seg2.simulate(upload=False); seg2.synchronize()
    assert not whole.check_status().any() and not seg2.check_status().any()
    for i in range(len(specs)):
        w = whole.frames_of(i).cpu().numpy()
        assert np.array_equal(w[:T1], a[i]), i
        assert np.array_equal(w[T1 - 1:], seg2.frames_of(i).cpu().numpy()), i


def test_k1_protocol_checks(env):
    """K1 compiled with -DCAB_K1_CHECK (the split-phase barrier's buffer-reuse invariants and the cluster halo
    protocol as run-time checks, integrate.cuh) on oversubscribed single-CTA and cluster ensembles: no step may see
    a buffer outside its time window, and the checked build produces the same bits.  Stands in for
    compute-sanitizer's racecheck, which this GPU pool does not offer."""
    import json, subprocess, sys
    from calcium_b200 import build
    lib = build.build_k1_check()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "k1_check_run.py")], capture_output=True, text=True,
                         env=dict(os.environ, CALCIUM_B200_LIB=lib), timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["lib"] == lib and len(out["cases"]) == 6
    for c in out["cases"]:
        assert c["status_or"] == 0 and not c["protocol_violation"] and c["bit_exact"], c
