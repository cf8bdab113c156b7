"""Host logic and the C-ABI library without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from calcium_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "cab200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(cab200_[a-z0-9_]+)\s*\(", header)))
    from calcium_b200 import _lib
    assert declared == sorted(_lib.EXPORTS), "keep _lib.EXPORTS in sync with include/cab200.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libcalcium_b200.so"
    assert lib.cab200_version() == 100


def test_library_targets_sm100a_only():
    so = os.path.join(ROOT, "calcium_b200", "libcalcium_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_compute_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from calcium_b200.ensemble import Ensemble, make_specs
    with pytest.raises(RuntimeError, match="CUDA"):
        Ensemble(make_specs(1, size="xsmall"))
    info = np.zeros(4, dtype=np.int64)
    assert lib.cab200_device_info(0, info.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))) == -2
    assert b"cuda" in lib.cab200_last_error().lower()


def test_argument_validation_needs_no_gpu(lib):
    from calcium_b200 import _lib
    z32 = np.zeros(4, dtype=np.int32); z64 = np.zeros(4, dtype=np.int64)
    rc = lib.cab200_sim_batch(1, _lib.as_i32p(z32), 0, _lib.as_i32p(z32), _lib.as_i32p(z32),
                              _lib.as_i32p(z32), _lib.as_i64p(z64), _lib.as_i64p(z64), None, None, None,
                              None, None, None, _lib.as_i64p(z64), None, None, None, None, None, 10, 0,
                              _lib.as_i32p(z32), None, None, 64, None, 0, None, 0, None)
    assert rc == -1 and b"bad sizes" in lib.cab200_last_error()
    # K3's argument checks and host-side table builder need no GPU either
    assert lib.cab200_encode_tables(4096, 64, 1, None) == -1 and b"[1, 2048]" in lib.cab200_last_error()
    rc = lib.cab200_encode_batch(1, _lib.as_i32p(z32), 1, _lib.as_i32p(z32), _lib.as_i64p(z64), None, None, 64, 64, 3,
                                 None, 0.1, 1, 1, 1, None, None, _lib.as_i64p(z64), None, None, 0, None, None, None, 0,
                                 None, 0, None)
    assert rc == -1 and b"null pointer" in lib.cab200_last_error()
    assert lib.cab200_publish_to_host(None, None, 16, None) == -1
    assert lib.cab200_sim_frames_tmp_bytes(1500, 1, 1, 64, 2, 10) == 2 * 10 * 4 * 1536 * 8
    assert lib.cab200_sim_frames_tmp_bytes(3255, 4, 1, 64, 1, 10) == 10 * 4 * (4 * 832) * 8        # 4 CTAs of 416 x 2 slots
    assert lib.cab200_sim_set_plan(2000, 333, 2) == -1
    assert lib.cab200_sim_set_plan(0, 0, 0) == 0


def test_layout_and_planner(lib):
    from calcium_b200 import _lib
    from calcium_b200.geometry import get_geometry
    lay = np.zeros(4, dtype=np.int64)
    assert lib.cab200_sim_layout(1500, 1, 64, _lib.as_i64p(lay)) == 0
    threads, cpt, copies, smem = lay.tolist()
    assert threads * cpt >= 1500 and copies == 2 and smem <= 227 * 1024
    assert lib.cab200_sim_layout(3255, 1, 64, _lib.as_i64p(lay)) == 0 and lay[2] == 2 and lay[3] <= 227 * 1024
    assert lib.cab200_sim_layout(4000, 1, 64, _lib.as_i64p(lay)) == 0 and lay[2] == 1
    assert lib.cab200_sim_layout(5000, 1, 64, _lib.as_i64p(lay)) == -3
    g = get_geometry("small")
    rowptr, col, _ = g.csr
    n = g.n_cells
    for copies in (1, 2):
        plan = np.zeros(2 * n, dtype=np.uint16)
        cost = np.zeros(4, dtype=np.int64)
        rc = lib.cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1, 2, 20000, 7, 0,
                                   plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), _lib.as_i64p(cost))
        assert rc == 0
        assert sorted(plan[:n].tolist()) == list(range(n)), "slot_of must be a permutation"
        assert cost[1] == cost[2] <= cost[0]
        lens = np.diff(rowptr)
        assert np.all(plan[n:] < (1 << lens.max())) and (copies == 2 or not plan[n:].any())
        # cost[3] = warp-wide gather loads per step: never below the off-diagonal entries / 32 (the
        # diagonal is added from registers), and the shape-sorted start keeps the padding below 50 % (the warp that
        # holds the rows with 6 entries before their diagonal runs all five position pairs)
        offdiag = int(lens.sum()) - n
        assert offdiag / 32 <= cost[3] <= 1.5 * offdiag / 32 + 2, (cost[3], offdiag / 32)
    assert lib.cab200_k1_legacy_rows(n, _lib.as_i32p(rowptr), _lib.as_i32p(col)) <= 8


def test_planner_models_the_entry_size(lib):
    """cab200_plan_slots_entry: the FP32 mode loads 8-byte entries (LDS.64: 16 lanes and 16 bank groups per
    phase), FP64 16-byte ones (8 and 8).  Same output format, cab200_plan_slots == entry size 16, annealing lowers the
    cost under the model it was given, and a conflict-free phase costs 2 wavefronts per warp-wide load with 8-byte
    entries (4 with 16-byte ones)."""
    from calcium_b200 import _lib
    from calcium_b200.geometry import get_geometry
    g = get_geometry("small")
    rowptr, col, _ = g.csr
    n = g.n_cells
    args = (n, _lib.as_i32p(rowptr), _lib.as_i32p(col), 2, 1, 2)

    def run(entry, iters, use_input=0, start=None):
        plan = np.zeros(2 * n, dtype=np.uint16) if start is None else start.copy()
        cost = np.zeros(4, dtype=np.int64)
        assert lib.cab200_plan_slots_entry(*args, entry, iters, 7, use_input, plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                           _lib.as_i64p(cost)) == 0
        assert sorted(plan[:n].tolist()) == list(range(n))
        return plan, cost

    p16, c16 = run(16, 20000)
    ref = np.zeros(2 * n, dtype=np.uint16)
    cref = np.zeros(4, dtype=np.int64)
    assert lib.cab200_plan_slots(*args, 20000, 7, 0, ref.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), _lib.as_i64p(cref)) == 0
    assert np.array_equal(ref, p16) and np.array_equal(cref, c16)
    p8, c8 = run(8, 20000)
    assert c8[1] == c8[2] < c8[0] and c8[3] == c16[3]          # same warp ranges, another conflict model
    assert 2 * c8[3] <= c8[1] and 4 * c16[3] <= c16[1]         # never below the conflict-free count
    _, c8_of_16 = run(8, 0, use_input=1, start=p16)            # the FP64 plan judged by the FP32 model: worse than its own
    assert c8_of_16[1] > c8[1]
    assert lib.cab200_plan_slots_entry(*args, 12, 0, 7, 0, p8.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None) != 0


def test_planner_gives_every_warp_one_diagonal_point(lib):
    """K1 adds a row's diagonal term from registers after D gather positions, D one value per warp (k1_layout.cuh):
    the planner must put the rows with more than EP = 5 entries on one side of their diagonal into warps whose other
    rows fit the same D.  Restated here: per warp of the plan, some odd D with max_nb <= D <= 10 - max_na exists --
    for the single-CTA plans of every size and for every part of the cluster plans."""
    from calcium_b200 import _lib
    from calcium_b200.ensemble import cluster_layout, plan_layout
    from calcium_b200.geometry import get_geometry

    def check(cells, slot, rowptr, col, per, part=None):
        nb = {int(c): int((col[rowptr[c]:rowptr[c + 1]] < c).sum()) for c in cells}
        na = {int(c): int((col[rowptr[c]:rowptr[c + 1]] > c).sum()) for c in cells}
        special = 0
        for w in range((max(slot[c] for c in cells) // per) + 1):
            mine = [int(c) for c in cells if slot[c] // per == w]
            if not mine:
                continue
            mnb, mna = max(nb[c] for c in mine), max(na[c] for c in mine)
            assert any(mnb <= d <= 10 - mna for d in (1, 3, 5, 7, 9)), (w, mnb, mna)
            special += mnb > 5 or mna > 5
        return special

    for size in ("xsmall", "small", "medium", "large"):
        g = get_geometry(size)
        rowptr, col, _ = g.csr
        n = g.n_cells
        assert np.diff(rowptr).max() <= 10
        copies, cpt = plan_layout(n, True, 64)
        plan = np.zeros(2 * n, dtype=np.uint16)
        assert lib.cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1, cpt, 30000, 3, 0,
                                     plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None) == 0
        nspecial = lib.cab200_k1_legacy_rows(n, _lib.as_i32p(rowptr), _lib.as_i32p(col))
        warps = check(range(n), plan[:n].astype(int), rowptr, col, 32 * cpt)
        assert warps <= (2 if nspecial else 0)                   # one warp per kind of special row is enough here
    g = get_geometry("large")
    rowptr, col, _ = g.csr
    n = g.n_cells
    copies, cpt = cluster_layout(n, 4, 64)
    cap = 3 * n + 4 + 4 * n
    plan = np.zeros(cap, dtype=np.uint16)
    assert lib.cab200_plan_cluster(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), 4, copies, cpt, 20000, 3,
                                   plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), cap, None) > 0
    part, slot = plan[:n].astype(int), plan[n:2 * n].astype(int)
    for q in range(4):
        check(np.nonzero(part == q)[0], slot, rowptr, col, 32 * cpt)
    # a row with several diagonal entries fits no diagonal point: the planner says so
    rp = np.array([0, 3, 5], dtype=np.int32)
    cc = np.array([0, 0, 1, 0, 1], dtype=np.int32)
    assert lib.cab200_k1_legacy_rows(2, _lib.as_i32p(rp), _lib.as_i32p(cc)) == -1
    plan2 = np.zeros(4, dtype=np.uint16)
    assert lib.cab200_plan_slots(2, _lib.as_i32p(rp), _lib.as_i32p(cc), 1, 1, 1, 0, 1, 0,
                                 plan2.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None) == -1
    assert b"unit = 0" in lib.cab200_last_error()


def test_cluster_planner(lib):
    """cab200_plan_cluster: parts cover the cells evenly, slots are unique inside a part, halo lists
    are exactly the other-part columns of each part's rows (sorted), and boundaries are thin."""
    from calcium_b200 import _lib
    from calcium_b200.geometry import get_geometry
    lay = np.zeros(4, dtype=np.int64)
    for size, n_cta in (("medium", 2), ("medium", 4), ("large", 4)):
        g = get_geometry(size)
        rowptr, col, _ = g.csr
        n = g.n_cells
        lay = np.zeros(4, dtype=np.int64)
        assert lib.cab200_cluster_layout(n, n_cta, 64, _lib.as_i64p(lay)) == 0
        threads, cpt, copies, smem = lay.tolist()
        assert threads * cpt * n_cta >= n and smem <= 227 * 1024
        cap = 3 * n + n_cta + n_cta * n
        plan = np.zeros(cap, dtype=np.uint16)
        cost = np.zeros(4, dtype=np.int64)
        got = lib.cab200_plan_cluster(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), n_cta, copies, cpt, 5000, 3,
                                      plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), cap, _lib.as_i64p(cost))
        assert got > 3 * n + n_cta
        part, slot = plan[:n].astype(int), plan[n:2 * n].astype(int)
        counts = np.bincount(part, minlength=n_cta)
        assert counts.max() - counts.min() <= 1 and counts.sum() == n
        hcount = plan[3 * n:3 * n + n_cta].astype(int)
        assert got == 3 * n + n_cta + hcount.sum() and cost[2] == hcount.sum()
        hl = plan[3 * n + n_cta:got].astype(int)
        off = 0
        for q in range(n_cta):
            mine = np.nonzero(part == q)[0]
            assert sorted(slot[mine].tolist()) == list(range(len(mine)))       # a permutation inside the part
            want = sorted({int(c) for i in mine for c in col[rowptr[i]:rowptr[i + 1]] if part[c] != q})
            assert hl[off:off + hcount[q]].tolist() == want
            off += hcount[q]
            assert hcount[q] <= 384                                            # CAB_HMAX
        assert hcount.sum() < 0.25 * n                                         # thin boundaries
    assert lib.cab200_cluster_layout(1500, 5, 64, _lib.as_i64p(lay)) == -1        # sizes 2, 3, 4, 8
    assert lib.cab200_cluster_layout(3255, 3, 64, _lib.as_i64p(lay)) == 0 and lay[:2].tolist() == [512, 3]
    assert lib.cab200_cluster_layout(6500, 8, 64, _lib.as_i64p(lay)) == 0 and lay[:2].tolist() == [416, 2]
    assert lib.cab200_cluster_layout(3255, 4, 64, _lib.as_i64p(lay)) == 0 and lay[:2].tolist() == [416, 2]
    assert lib.cab200_cluster_layout(8000, 8, 64, _lib.as_i64p(lay)) == 0 and lay[:2].tolist() == [512, 2]
    assert lib.cab200_cluster_layout(3255, 2, 64, _lib.as_i64p(lay)) == -3      # 1628 cells per CTA > 1536
    assert lib.cab200_cluster_layout(20000, 4, 64, _lib.as_i64p(lay)) == -3


def test_parameters_api():
    from calcium_b200.parameters import SimulationParameters
    sp = SimulationParameters("Fluttering", custom_params={"D_p": 0.01})
    assert sp.params["lower"] == 1.4 and sp.params["D_p"] == 0.01 and len(sp.params) == 18
    assert SimulationParameters("nope").sim_type == "Intercellular waves"
    assert sp.validate()
    with pytest.raises(ValueError):
        SimulationParameters(custom_params={"beta": -1.0}).validate()
    with pytest.raises(ValueError):
        SimulationParameters(custom_params={"lower": 0.9, "upper": 0.5}).validate()
    assert SimulationParameters.get_available_sim_types()[0] == "Single cell spikes"


def test_parameters_roundtrip(tmp_path):
    from calcium_b200.parameters import SimulationParameters
    sp = SimulationParameters("Single cell spikes")
    f = tmp_path / "a" / "p.json"
    sp.save_to_file(str(f))
    sp2 = SimulationParameters.load_from_file(str(f))
    assert sp2.params == sp.params and sp2.sim_type == sp.sim_type


def test_geometry_loader_reference_format(tmp_path):
    from calcium_b200 import geometry as G
    G.write_reference_format(str(tmp_path), sizes=("xsmall",))
    ld = G.GeometryLoader(str(tmp_path))
    assert ld.get_available_sizes() == ["xsmall"]
    v, adj, lap = ld.load_geometry("xsmall")
    g = G.get_geometry("xsmall")
    assert len(v) == 200 and adj.shape == (200, 200) and np.array_equal(lap, g.laplacian)
    with pytest.raises(ValueError, match="not available"):
        ld.load_geometry("large")
    # file-backed geometry yields the same CSR as the synthetic object it was written from
    gf = ld.geometry("xsmall")
    for a, b in zip(gf.csr, g.csr):
        assert np.array_equal(a, b)
    assert ld.get_geometry_info("xsmall")["n_cells"] == 200


def test_host_draws_match_reference_sequence(golden_dir):
    from calcium_b200.ensemble import draw_initial_state, draw_vplc, pack_params
    from calcium_b200.parameters import SimulationParameters
    z = np.load(os.path.join(golden_dir, "traj_xsmall_waves.npz"))
    p = SimulationParameters("Intercellular waves").get_params_dict()
    v = draw_vplc(p, 200, 0)
    assert v.shape == (200, 1)
    r0, init = draw_initial_state(p, 200, 0, v)
    assert np.array_equal(v.flatten(), z["vplc"]) and np.array_equal(r0, z["r0"])
    assert len(init) == max(1, int(p["frac"] * 200))
    prm = pack_params(p)
    assert prm.shape == (16,) and prm[13] == p["D_c_ratio"] * p["D_p"] and prm[15] == 0.2


def test_pouch_validation_without_gpu():
    from calcium_b200.pouch import Pouch
    with pytest.raises(ValueError, match="Missing: D_p"):
        Pouch(params={k: 1.0 for k in Pouch.REQUIRED_PARAMS if k != "D_p"}, size="xsmall")
    with pytest.raises(ValueError, match="Unexpected: bogus"):
        Pouch(params={**{k: 1.0 for k in Pouch.REQUIRED_PARAMS}, "bogus": 1}, size="xsmall")
    p = Pouch(size="xsmall", sim_number=3)
    assert p.n_cells == 200 and p.T == 18000 and p.dt == 0.2 and p.V_PLC.shape == (200, 1)
    assert p.disc_dynamics.shape == (200, 4, 18000)
    assert p.D_c == p.D_c_ratio * p.D_p                      # reference defect D1 fixed
    assert p.get_active_cells(100) == []                     # all-zero state before simulate()
    assert p.adj_matrix.shape == (200, 200) and p.laplacian_sparse.nnz == len(p._geometry.csr[1])
    with pytest.raises(ValueError, match="time_step must be provided"):
        p.get_cell_masks(active_only=True)
    with pytest.raises(ValueError, match="Time step must be between"):
        p.generate_image(18000)
    bad = Pouch(params={**p.param_dict, "beta": 0.0}, size="xsmall")
    with pytest.raises(ValueError, match="must be positive"):
        bad.simulate()
    bad = Pouch(params={**p.param_dict, "frac": 1.5}, size="xsmall")
    with pytest.raises(ValueError, match="fraction of initiator"):
        bad.simulate()


def test_lpt_partition():
    from calcium_b200.parallel import lpt_partition, shard_indices
    parts = shard_indices([1500] * 1184, 18000, 8)
    assert [len(p) for p in parts] == [148] * 8
    assert sorted(i for p in parts for i in p) == list(range(1184))
    sizes = [200, 500, 1500, 3255] * 64
    parts = shard_indices(sizes, 18000, 8)
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 3255 and sorted(i for p in parts for i in p) == list(range(256))
    assert lpt_partition([], 4) == [[], [], [], []]
    assert lpt_partition([5.0], 1) == [[0]]


def test_auto_cluster_policy():
    """CTAs per pouch under cluster="auto": the shortest predicted time given how many clusters of each size are
    resident at once (B200: 74 / 45 / 33 / 15 clusters of 2 / 3 / 4 / 8)."""
    from types import SimpleNamespace as NS
    from calcium_b200.ensemble import _DeviceGeometry as DG
    caps = {2: 74, 3: 45, 4: 33, 8: 15}
    pick = lambda n, count, unit=True: DG.auto_cluster(NS(unit=unit, n=n, capacity=lambda nc: caps[nc]), count, 148)
    assert pick(3255, 1) == 4 and pick(3255, 33) == 4 and pick(3255, 148) == 4
    assert pick(3255, 40) == 3 and pick(3255, 45) == 3 and pick(3255, 90) == 3      # 33 resident clusters of 4 would need one more wave
    assert pick(1500, 1) == 4 and pick(1500, 33) == 4 and pick(1500, 34) == 2
    assert pick(1500, 74) == 2 and pick(1500, 75) == 1 and pick(1500, 148) == 1
    assert pick(500, 1) == 1 and pick(200, 3) == 1
    assert pick(6500, 1) == 8 and pick(6500, 100) == 8
    assert pick(3255, 1, unit=False) == 1          # general weights: single-CTA kernel only


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "calcium_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f
    code = "import sys; import calcium_b200, calcium_b200.ensemble, calcium_b200.pouch, calcium_b200.batch; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_io_writers(tmp_path):
    import cv2
    from calcium_b200 import io
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, (40, 56, 2)).astype(np.uint8)
    path = io.save_gray_alpha_png(img, str(tmp_path / "x" / "a.png"))
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint16 and back.shape == (40, 56, 4)
    assert np.array_equal(back[:, :, 0], img[:, :, 0].astype(np.uint16) * 4)      # 8 -> 10 bit
    assert np.array_equal(back[:, :, 3], img[:, :, 1].astype(np.uint16) * 257)
    m = rng.randint(0, 900, (40, 56)).astype(np.uint16)
    mp = io.save_instance_mask(m, str(tmp_path / "m.npz"))
    assert np.array_equal(np.load(mp)["mask"], m)
    cp = io.save_csv_mapping([("a.png", "a.npz")], str(tmp_path), "train")
    assert open(cp).read().splitlines() == ["ImageId,MaskId", "a.png,a.npz"]


def test_create_instance_mask_helper(oracle):
    from calcium_b200.masks import create_instance_mask_from_cells
    rng = np.random.RandomState(0)
    for _ in range(25):
        m = rng.randint(0, 40, (30, 30)).astype(np.int32)
        act = sorted(rng.choice(40, rng.randint(0, 15), replace=False).tolist())
        assert np.array_equal(create_instance_mask_from_cells(m, act), oracle.instance_mask_oracle(m, act))
    assert np.array_equal(create_instance_mask_from_cells(m), m.astype(np.uint16))


def test_batch_device_list_resolution():
    import torch
    from calcium_b200.batch import _resolve_devices
    assert _resolve_devices("cuda:1", None) == ["cuda:1"]
    assert _resolve_devices(None, [0, "cuda:3"]) == [torch.device("cuda", 0), torch.device("cuda", 3)]
    with pytest.raises(ValueError):
        _resolve_devices(None, [])
    with pytest.raises(ValueError):
        _resolve_devices(None, "both")


def test_k1_gather_order_emulation_on_random_graphs(lib, tmp_path):
    """CPU emulation of K1's gather order (tests/csrc/k1_layout_check.cpp, built on the layout header the kernel and the
    planner share): for the plans of every synthetic geometry and of random graphs with long rows on either side of the
    diagonal, every stored entry of every row is added exactly once, in CSR order, with the register diagonal at its
    place and inside the warp's position range -- before and after annealing."""
    import subprocess
    from calcium_b200 import _lib
    from calcium_b200.ensemble import plan_layout
    from calcium_b200.geometry import Geometry, get_geometry
    so = str(tmp_path / "k1_layout_check.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "csrc", "k1_layout_check.cpp"), "-o", so],
                   check=True)
    chk = ctypes.CDLL(so)
    chk.k1_layout_check.restype = ctypes.c_longlong

    def fuzz_graph(trial, n):
        rng = np.random.RandomState(100 + trial)
        deg = np.zeros(n, dtype=int)
        edges = set()

        def link(a, b):
            if a != b and (min(a, b), max(a, b)) not in edges and deg[a] < 9 and deg[b] < 9:
                edges.add((min(a, b), max(a, b))); deg[a] += 1; deg[b] += 1
        n_hub = max(1, n // (40 if trial % 3 == 2 else 150))
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
        return Geometry(size=f"fuzz{n}", vertices=verts, edges=np.array(sorted(edges), dtype=np.int32).reshape(-1, 2))

    geos = [get_geometry(s_) for s_ in ("xsmall", "small", "medium", "large")]
    geos += [fuzz_graph(t, n) for t, n in enumerate((24, 40, 90, 150, 333, 700, 1300))]
    checked = special_points = 0
    for g in geos:
        rowptr, col, _ = g.csr
        n = g.n_cells
        copies, cpt = plan_layout(n, True, 64)
        lay = np.zeros(4, dtype=np.int64)
        assert lib.cab200_sim_layout(n, 1, 64, _lib.as_i64p(lay)) == 0
        n_slots = int(lay[0] * lay[1])
        for iters in (0, 40000):
            plan = np.zeros(2 * n, dtype=np.uint16)
            rc = lib.cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1, cpt, iters, 5, 0,
                                       plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), None)
            if rc != 0:                                   # refused: this geometry runs the general kernel
                assert b"unit = 0" in lib.cab200_last_error()
                break
            wd = np.zeros(64, dtype=np.int32)
            bad = chk.k1_layout_check(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                      cpt, n_slots, _lib.as_i32p(wd))
            assert bad == 0, (g.size, iters, int(bad))
            checked += 1
            special_points += int(np.isin(wd, (1, 3, 7, 9)).sum())
    assert checked >= 12 and special_points >= 10         # warps with diagonal points 1, 3, 7, 9 were among them
