"""INTEGRATION.md section B is executable: the ctypes stub a maintainer of the reference would add is extracted from
the document itself and run -- against a stand-in object carrying the attributes of the reference's ``Pouch`` that
the stub reads, and, where the unmodified reference travelled with the tree (``oracle/_ref``), against the real
``Pouch``: its own ``simulate()`` loop versus the stub's replacement of that loop (``core/pouch.py:338-366``)."""
import os
import re
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_stub():
    from calcium_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = [b for b in blocks if "def simulate_on_gpu" in b]
    assert len(code) == 1, "INTEGRATION.md must hold exactly one stub defining simulate_on_gpu"
    src = code[0].replace('ctypes.CDLL("libcalcium_b200.so")', f'ctypes.CDLL({_lib.LIB_PATH!r})')
    ns: dict = {}
    exec(compile(src, "INTEGRATION.md:stub", "exec"), ns)
    return ns


@pytest.mark.gpu
def test_integration_stub_on_a_stand_in_pouch(cuda_device, oracle):
    import scipy.sparse as sp
    from calcium_b200 import ensemble as E
    from calcium_b200.geometry import get_geometry
    from calcium_b200.parameters import SimulationParameters as SP
    ns = _load_stub()
    g = get_geometry("small")
    prm = SP("Intercellular waves").generate_random_params(seed=4)
    v = E.draw_vplc(prm, g.n_cells, 4)
    r0, _ = E.draw_initial_state(prm, g.n_cells, 4, v)
    T, steps = 600, [0, 1, 300, 599]
    rowptr, col, val = g.csr
    dd = np.zeros((g.n_cells, 4, T))
    dd[:, 3, 0] = r0
    fake = types.SimpleNamespace(                                   # what the stub reads of the reference's Pouch
        n_cells=g.n_cells, T=T, dt=0.2, D_c=prm["D_c_ratio"] * prm["D_p"], param_dict=prm, V_PLC=v,
        laplacian_sparse=sp.csr_matrix((val, col, rowptr), shape=(g.n_cells, g.n_cells)), disc_dynamics=dd)
    assert ns["simulate_on_gpu"](fake, steps, device=str(cuda_device)) is True
    z = np.zeros(g.n_cells)
    want = oracle.simulate_c(g.csr, oracle.pack_params(prm), z, z, r0, v.flatten(), T, steps)
    for i, t in enumerate(steps):
        assert np.array_equal(fake.disc_dynamics[:, :, t], want[i].T), t


@pytest.mark.gpu
def test_integration_stub_replaces_the_loop_of_the_real_reference(cuda_device):
    from oracle import reference_shims as R
    if not R.reference_available():
        pytest.skip("the reference did not travel with the tree (python -m oracle.make_ref in the build container)")
    ns = _load_stub()
    ref = R.load_reference()
    gdir = R.synthetic_geometry_dir()
    prm = ref.SimulationParameters("Fluttering").generate_random_params(seed=21)
    T, steps = 1200, [0, 1, 2, 600, 1199]
    pouches = []
    for _ in range(2):
        p = ref.Pouch(params=prm, size="xsmall", sim_number=21, geometry_dir=gdir)
        p.T = T
        p.disc_dynamics = np.zeros((p.n_cells, 4, T))
        pouches.append(p)
    a, b = pouches
    a.simulate()                                                     # the reference's own numba loop
    # b: the reference's preamble (core/pouch.py:317-334) then the stub instead of the loop
    np.random.seed(b.sim_number)
    b.disc_dynamics[:, 2, 0] = (b.c_tot - b.disc_dynamics[:, 0, 0]) / b.beta
    b.disc_dynamics[:, 3, 0] = np.random.uniform(0.5, 0.7, size=(b.n_cells, 1)).T
    n_init = max(1, int(b.frac * b.n_cells))
    init = np.random.choice(b.n_cells, n_init, replace=False)
    b.V_PLC[init, 0] = np.random.uniform(1.3, 1.5, len(init))
    assert np.array_equal(a.V_PLC, b.V_PLC) and np.array_equal(a.disc_dynamics[:, 3, 0], b.disc_dynamics[:, 3, 0])
    ns["simulate_on_gpu"](b, steps, device=str(cuda_device))
    for t in steps:
        x, y = a.disc_dynamics[:, :, t], b.disc_dynamics[:, :, t]
        assert np.abs(x - y).max() <= 1e-9 * max(1.0, np.abs(x).max()), t
