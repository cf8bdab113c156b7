"""GPU fuzz test of K1's row layout (runs last: the file name sorts after the other GPU test files)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def test_random_graphs_with_long_rows_on_either_side_of_the_diagonal(env):
    """K1 adds a row's diagonal term from registers at one point per warp; rows with more than 5 entries before, or
    after, their diagonal pin their warp's point (k1_layout.cuh).  Random graphs whose low / high cell ids collect many
    neighbours exercise both kinds at once, across several warps and CTA shapes; graphs too small to give each kind a
    warp of its own take the general kernel (``unit`` switched off by the planner's refusal).  Always the oracle's bits."""
    E, O = env["E"], env["O"]
    from calcium_b200.geometry import Geometry
    from calcium_b200.parameters import SimulationParameters as SP
    prm = SP("Intercellular waves").get_params_dict()
    seen_unit, seen_general, seen_special = 0, 0, 0
    for trial, n in enumerate((24, 40, 90, 150, 333, 700, 1300)):
        rng = np.random.RandomState(100 + trial)
        deg = np.zeros(n, dtype=int)
        edges = set()
        # hubs: a few low ids linked to many higher ids (many entries AFTER the diagonal), a few high ids linked to
        # many lower ids (many entries BEFORE it); then a sparse random background; rows stay within 9 neighbours
        def link(a, b):
            if a != b and (min(a, b), max(a, b)) not in edges and deg[a] < 9 and deg[b] < 9:
                edges.add((min(a, b), max(a, b))); deg[a] += 1; deg[b] += 1
        n_hub = max(1, n // (40 if trial % 3 == 2 else 150))    # every third graph: more long rows than the planner can seat
        for hub in rng.choice(max(2, n // 6), size=n_hub, replace=False):
            for j in rng.choice(np.arange(hub + 1, n), size=min(7, n - hub - 1), replace=False):
                link(int(hub), int(j))
        for hub in n - 1 - rng.choice(max(2, n // 6), size=n_hub, replace=False):
            for j in rng.choice(np.arange(0, hub), size=min(7, hub), replace=False):
                link(int(hub), int(j))
        for _ in range(2 * n):
            a, b = rng.randint(0, n, 2)
            link(int(a), int(b))
        verts = np.empty(n, dtype=object)
        for i in range(n):
            verts[i] = np.array([[i, 0.0], [i + 1, 0.0], [i + 1, 1.0], [i, 1.0]])
        g = Geometry(size=f"fuzz{n}", vertices=verts, edges=np.array(sorted(edges), dtype=np.int32).reshape(-1, 2))
        rowptr, col, _ = g.csr
        nb = np.array([(col[rowptr[i]:rowptr[i + 1]] < i).sum() for i in range(n)])
        na = np.array([(col[rowptr[i]:rowptr[i + 1]] > i).sum() for i in range(n)])
        seen_special += int(((nb > 5) | (na > 5)).sum())
        specs = [E.PouchSpec(params=prm, geometry=g, vplc=rng.uniform(0.4, 0.8, n), r0=rng.uniform(.5, .7, n),
                             c0=rng.uniform(0, .2, n), p0=rng.uniform(0, .3, n)) for _ in range(2)]
        ens = E.Ensemble(specs, T=1500, frame_steps=[0, 700, 1499], device=env["dev"])
        seen_unit += int(ens.g_unit[0]); seen_general += int(not ens.g_unit[0])
        ens.simulate()
        assert not ens.check_status().any(), n
        for i, sp in enumerate(specs):
            assert np.array_equal(_oracle_frames(O, sp, 1500, [0, 700, 1499]), ens.frames_of(i).cpu().numpy()), (n, i)
    assert seen_special > 20 and seen_unit >= 3              # the register-diagonal kernel met long rows of both kinds
