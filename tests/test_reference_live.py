"""Oracle pinned to the LIVE reference (build container only: /root/reference is not on the GPU
box, where this module skips).  Short horizon so it runs in seconds."""
import numpy as np
import pytest

from oracle import reference_shims as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return R.load_reference(), R.synthetic_geometry_dir()


@pytest.mark.parametrize("sim_type,seed", [("Single cell spikes", 11), ("Fluttering", 12)])
def test_oracle_tracks_live_reference(ref, oracle, sim_type, seed):
    ns, gdir = ref
    from calcium_b200 import geometry as G
    prm = ns.SimulationParameters(sim_type).generate_random_params(seed=seed)
    pouch = ns.Pouch(params=prm, size="xsmall", sim_number=seed, geometry_dir=gdir)
    pouch.T = 1500                                   # horizon override
    pouch.disc_dynamics = np.zeros((pouch.n_cells, 4, pouch.T))
    pouch.simulate()
    g = G.get_geometry("xsmall")
    vplc, r0, _ = oracle.host_draws(prm, g.n_cells, seed)
    assert np.array_equal(vplc, pouch.V_PLC.flatten())
    steps = [0, 1, 700, 1499]
    z = np.zeros(g.n_cells)
    got = oracle.simulate_c(g.csr, oracle.pack_params(prm), z, z, r0, vplc, 1500, steps)
    want = np.transpose(pouch.disc_dynamics[:, :, steps], (2, 1, 0))
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    assert rel.max() < 1e-9
    # the reference's own raster methods against the oracle's gather formulation
    for t in (700, 1499):
        ca = pouch.disc_dynamics[:, 0, t]
        img = pouch.generate_image(t)
        li = oracle.paint_label_map_cv2(g.vertices, (512, 512))
        o_img, o_act, o_inst, _ = oracle.raster_frame_c(ca, li, pouch.cells_mask, 0.1, True)
        assert np.array_equal(img, o_img)
        act = pouch.get_active_cells(t)
        am = pouch.get_cell_masks(active_only=True, time_step=t)
        assert np.array_equal(am, o_act)
        assert np.array_equal(ns.create_instance_mask_from_cells(am, act), o_inst)


def test_product_parameters_match_live_reference(ref):
    ns, _ = ref
    from calcium_b200.parameters import SimulationParameters
    for st in SimulationParameters.get_available_sim_types():
        for seed in (0, 5, 99, 1023):
            a = SimulationParameters(st).generate_random_params(seed=seed)
            b = ns.SimulationParameters(st).generate_random_params(seed=seed)
            assert a == b and list(a) == list(b)
