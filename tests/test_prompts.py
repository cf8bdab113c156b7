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


def test_prompt_file_text_is_json_dump_indent2():
    """save_prompts formats the file directly; the text must be what json.dump(..., indent=2) writes."""
    from calcium_b200.prompts import dumps_prompts
    rng = np.random.RandomState(0)
    for _ in range(100):
        n, m = rng.randint(0, 30), rng.randint(0, 6)
        pos = {int(i): {"point": [int(rng.randint(512)), int(rng.randint(512))], "label": 1, "area": int(rng.randint(10, 9999)),
                        "confidence": float(np.sqrt(np.float64(rng.randint(1, 400))))}
               for i in sorted(rng.choice(2000, n, replace=False))}
        neg = [{"point": [int(rng.randint(512)), int(rng.randint(512))], "label": 0,
                "confidence": float(np.sqrt(np.float64(rng.randint(25, 90000))))} for _ in range(m)]
        want = json.dumps({"positive_prompts": pos, "negative_prompts": neg, "num_cells": len(pos), "version": "1.0"}, indent=2)
        assert dumps_prompts(pos, neg) == want
        assert json.loads(dumps_prompts(pos, neg))["num_cells"] == n


def test_vector_randint_is_the_scalar_sequence():
    """frame_prompts draws a run of single-cell instances with one array call; the legacy generator must give
    the values of the reference's scalar calls and end in the same state."""
    rng = np.random.RandomState(5)
    for trial in range(50):
        cnts = rng.randint(1, 400, rng.randint(1, 1500))
        np.random.seed(trial)
        a = np.array([np.random.randint(c) for c in cnts])
        sa = np.random.get_state()
        rs = np.random.RandomState(trial)
        b = rs.randint(0, cnts)
        sb = rs.get_state()
        assert np.array_equal(a, b) and sa[2] == sb[2] and np.array_equal(sa[1], sb[1])


@pytest.mark.gpu
def test_gpu_prompts_match_the_oracle_on_medium_frames(cuda_device):
    """frame_prompts (static tables + K4 + array draws, private generator) against the oracle's restatement of
    the reference functions (scipy distance transforms on the instance mask, global generator) on frames of
    medium pouches of all four sim types -- including frames whose id mismatch leaves no labelled pixel at all
    and frames with the background instance."""
    from calcium_b200 import build
    build.build()
    from calcium_b200 import ensemble as E
    from calcium_b200.prompts import PromptTables, dumps_prompts, frame_prompts, instance_lut
    from oracle import prompts_oracle as PO
    specs = E.make_specs(4, size="medium")
    steps = [0, 4000, 9000, 17999]
    ens = E.Ensemble(specs, frame_steps=steps, output_size=(256, 256), device=cuda_device)
    ens.simulate()
    lmsk = ens.label_maps()[1][0]
    cm = lmsk.cpu().numpy().view(np.uint16)
    tables = PromptTables(lmsk, specs[0].geometry.n_cells)
    seen_empty = seen_bg = 0
    for i in range(4):
        ca_all = ens.frames_of(i)[:, 0, :].cpu().numpy()
        for f in range(len(steps)):
            for thr in (0.1, 0.0):
                if not (ca_all[f] > thr).any():
                    continue
                lut = instance_lut(ca_all[f], thr, True)
                inst = lut[cm]
                seen_empty += int(not (inst > 0).any())
                seen_bg += int(lut[0] != 0)
                np.random.seed(100 * i + f)
                want = (PO.point_prompts(inst), PO.negative_prompts(inst, 5))
                end = np.random.get_state()
                rs = np.random.RandomState(100 * i + f)
                got = frame_prompts(tables, ca_all[f], threshold=thr, rng=rs)
                assert got[0] == want[0] and list(got[0]) == list(want[0]), (i, f, thr)
                assert got[1] == want[1], (i, f, thr)
                st = rs.get_state()
                assert st[2] == end[2] and np.array_equal(st[1], end[1])       # generator left in the same state
                rt = np.random.RandomState(100 * i + f)
                text, n_pos = frame_prompts(tables, ca_all[f], threshold=thr, rng=rt, as_text=True)
                assert text == dumps_prompts(got[0], got[1]) and n_pos == len(got[0])
    assert seen_empty > 0 and seen_bg > 0


@pytest.mark.gpu
def test_batch_prompts_do_not_depend_on_the_thread_count(cuda_device, tmp_path):
    from calcium_b200 import generate_simulation_batch
    import random
    outs = []
    for k, threads in enumerate((1, 6)):
        random.seed(3)
        d = tmp_path / f"run{k}"
        res = generate_simulation_batch(5, str(d), pouch_sizes=["xsmall", "small"], time_steps=[0, 6000, 12000],
                                        image_size="128x128", prompt_threads=threads, defect_configs=[{}] * 5)
        files = {}
        pdir = d / "prompts" / "train"
        for name in sorted(os.listdir(pdir)):
            key = "_".join(name.split("_")[:-1])            # drop the run id
            files[key] = open(pdir / name).read()
        outs.append((files, [s["prompt_count"] for s in res["simulations"]], np.random.get_state()[1].copy()))
    assert outs[0][0] == outs[1][0] and len(outs[0][0]) > 0
    assert outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2])


def test_legacy_choice_is_numpys_choice():
    """cab200_legacy_choice (host code, no GPU): the indices AND the generator state numpy's own
    RandomState.choice(n, k, replace=False) produces, for many sizes, states (incl. a cached gaussian) and k."""
    from calcium_b200 import build
    build.build()
    from calcium_b200.prompts import legacy_choice
    sizes = [4096, 5000, 65535, 65536, 65537, 100000, 131071, 131072, 262143, 262144]
    for trial in range(40):
        n = sizes[trial % len(sizes)]
        k = [5, 1, 3, 17][trial % 4]
        a, b = np.random.RandomState(trial), np.random.RandomState(trial)
        pre = (trial * 37) % 700
        a.random_sample(pre); b.random_sample(pre)
        if trial % 5 == 0:
            a.normal(); b.normal()
        assert np.array_equal(a.choice(n, k, replace=False), legacy_choice(b, n, k)), (trial, n, k)
        sa, sb = a.get_state(), b.get_state()
        assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1]) and sa[3] == sb[3] and sa[4] == sb[4]
        assert np.array_equal(a.randint(0, 1000, 10), b.randint(0, 1000, 10))
    # the global generator (what the reference draws from) goes through the same path
    from calcium_b200.prompts import global_rng
    np.random.seed(9)
    want = np.random.choice(200000, 5, replace=False)
    w2 = np.random.randint(1000)
    np.random.seed(9)
    assert np.array_equal(legacy_choice(global_rng(), 200000, 5), want) and np.random.randint(1000) == w2


def test_sample_background_is_the_numpy_selection_and_draw():
    """cab200_sample_background (host code): candidate count, drawn pixels and generator state of
    flatnonzero(~labelled & (d2 >= 25)) + RandomState.choice(n, k, replace=False)."""
    from calcium_b200 import build
    build.build()
    from calcium_b200.prompts import sample_background
    R = np.random.RandomState(0)
    for trial in range(40):
        H, W = [(64, 80), (512, 512), (200, 333), (16, 16)][trial % 4]
        n = R.randint(5, 1500)
        labels = R.randint(0, n + 1, (H, W)).astype(np.uint16)
        lutpos = R.rand(n + 1) < [0.0, 0.02, 0.5, 1.0][trial % 4]
        if trial % 4 != 3:
            lutpos[0] = False
        d2 = R.randint(0, 60, (H, W)).astype(np.int32)
        k = [5, 1, 7, 5][trial % 4]
        a, b = np.random.RandomState(trial), np.random.RandomState(trial)
        a.random_sample(trial * 13 % 650); b.random_sample(trial * 13 % 650)
        flat = np.flatnonzero(~lutpos[labels] & (d2 >= 25))
        want = flat[a.choice(len(flat), min(k, len(flat)), replace=False)] if len(flat) else np.zeros(0, np.int64)
        got, cnt = sample_background(b, labels, lutpos, d2, 25, k)
        assert cnt == len(flat) and np.array_equal(got, want), trial
        sa, sb = a.get_state(), b.get_state()
        assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1]), trial


def _scipy_edt_squared(label_map, class_luts=None):
    """Stand-in for K4 on the CPU (tests only): exact squared distance to the nearest pixel of a different class."""
    import torch
    from scipy.ndimage import distance_transform_edt
    lab = label_map.cpu().numpy().view(np.uint16)
    maps = [lab] if class_luts is None else [class_luts.cpu().numpy().view(np.uint16)[i][lab] for i in range(class_luts.shape[0])]
    out = []
    for cls in maps:
        d2 = np.zeros(cls.shape, dtype=np.int64)
        for v in np.unique(cls):
            b = cls == v
            if b.all():
                d2[b] = 2 ** 30                      # no pixel of another class (K4 writes "far"; never selected)
                continue
            ys, xs = np.nonzero(b)
            y0, y1, x0, x1 = max(ys.min() - 1, 0), min(ys.max() + 2, cls.shape[0]), max(xs.min() - 1, 0), min(xs.max() + 2, cls.shape[1])
            sub = b[y0:y1, x0:x1]
            if sub.all():                            # the class touches every border of its box: use the full image
                sub, y0, x0 = b, 0, 0
            d = distance_transform_edt(sub)
            full = np.rint(d * d).astype(np.int64)
            d2[y0:y0 + sub.shape[0], x0:x0 + sub.shape[1]][sub] = full[sub]
        out.append(d2.astype(np.int32))
    return torch.from_numpy(np.stack(out))


@pytest.mark.parametrize("t,thr", CASES)
def test_frame_prompts_host_logic_on_cpu(gold, monkeypatch, t, thr):
    """The host side of the prompt path (static tables, array draws, the lock-free selection + draw, generator
    state) against the reference's recorded prompts, with scipy standing in for the K4 kernel."""
    import torch
    from calcium_b200 import build
    build.build()
    from calcium_b200 import prompts as P
    monkeypatch.setattr(P, "edt_squared", _scipy_edt_squared)

    def cpu_transforms(tables, lut, background=True, held=None):   # _frame_transforms without the device workspace
        rows = [(lut == 1), (lut > 0)] if background else [(lut > 0)]
        luts = torch.from_numpy(np.stack(rows).astype(np.uint16).view(np.int16))
        both = _scipy_edt_squared(tables.label_mask, luts).numpy()
        return (both[0], both[1]) if background else (None, both[0])
    monkeypatch.setattr(P, "_frame_transforms", cpu_transforms)
    z, j = gold
    g = j[f"{t}_{thr}"]
    lmsk = torch.from_numpy(z["cells_mask"].astype(np.uint16).view(np.int16))
    tables = P.PromptTables(lmsk, int(z["ca"].shape[1]))
    ca = z["ca"][list(z["t"]).index(t)]
    np.random.seed(g["seed"])
    pos, neg = P.frame_prompts(tables, ca, threshold=thr, quirk=True)
    assert pos == _norm(g["positive"])
    assert neg == g["negative"]
    rs = np.random.RandomState(g["seed"])
    pos2, neg2 = P.frame_prompts(tables, ca, threshold=thr, quirk=True, rng=rs)
    assert pos2 == pos and neg2 == neg
    # the text route of the batch path: the file save_prompts / json.dump(indent=2) would write, byte for byte
    rt = np.random.RandomState(g["seed"])
    text, n_pos = P.frame_prompts(tables, ca, threshold=thr, quirk=True, rng=rt, as_text=True)
    want = json.dumps({"positive_prompts": pos, "negative_prompts": neg, "num_cells": len(pos), "version": "1.0"}, indent=2)
    assert text == want and n_pos == len(pos)
    assert rt.get_state()[2] == rs.get_state()[2] and np.array_equal(rt.get_state()[1], rs.get_state()[1])
    end = np.random.get_state()
    assert rs.get_state()[2] == end[2] and np.array_equal(rs.get_state()[1], end[1])


def test_native_percentile_and_randint_are_numpys():
    """The numpy restatements inside cab200_prompts_pouch (host code, no GPU): np.percentile(sqrt(d2), 80) -- method
    "linear", two-sided lerp -- and RandomState.randint(0, high) element by element (masked rejection on 32-bit
    outputs, no draw for a range of one): same values, same generator state."""
    import ctypes
    from calcium_b200 import _lib, build
    build.build()
    lib = _lib.load()
    R = np.random.RandomState(5)
    for trial in range(300):
        m = [1, 2, 3, 4, 5, 6, 7, 11, 100, 1001, 4096, 99999][trial % 12] + (trial // 12) * (trial % 5)
        d2 = R.randint(1, [3, 50, 5000, 131072][trial % 4], m).astype(np.int32)
        want = np.percentile(np.sqrt(d2.astype(np.float64)), 80)
        got = ctypes.c_double(0)
        assert lib.cab200_np_percentile80_sqrt(d2.ctypes.data, m, ctypes.byref(got)) == 0
        assert got.value == want, (trial, m, got.value, want)
    for trial in range(60):
        a, b = np.random.RandomState(trial), np.random.RandomState(trial)
        a.random_sample(trial * 11 % 640); b.random_sample(trial * 11 % 640)
        k = [1, 5, 300, 1400][trial % 4]
        high = R.randint(1, [2, 9, 70, 100000][trial % 4], k).astype(np.int64)
        want = a.randint(0, high)
        name, key, pos, hg, cg = b.get_state()
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        cpos = ctypes.c_int32(int(pos))
        out = np.zeros(k, dtype=np.int64)
        assert lib.cab200_legacy_randint(key.ctypes.data, ctypes.addressof(cpos), high.ctypes.data, k, out.ctypes.data) == 0
        assert np.array_equal(out, want), trial
        sa = a.get_state()
        assert sa[2] == cpos.value and np.array_equal(sa[1], key), trial
        # ... and the scalar call the background instance's draw makes
        h = int(high[0])
        w1 = a.randint(h)
        assert lib.cab200_legacy_randint(key.ctypes.data, ctypes.addressof(cpos), high[:1].ctypes.data, 1, out.ctypes.data) == 0
        assert out[0] == w1 and a.get_state()[2] == cpos.value


@pytest.mark.gpu
def test_native_pouch_prompts_equal_the_frame_by_frame_path(cuda_device):
    """cab200_prompts_pouch (all frames of a pouch in one native call) against frame_prompts frame by frame: the same
    file text byte for byte, the same entry counts and the same generator state afterwards -- medium pouches of all
    four sim types (frames with and without the background instance, quiet frames, frames without any labelled
    pixel) and a small pouch at another output size."""
    import torch
    from calcium_b200 import build
    build.build()
    from calcium_b200 import prompts as P
    from calcium_b200.ensemble import Ensemble, make_specs, build_label_maps
    for size, osz, count in (("medium", (512, 512), 4), ("small", (200, 160), 4)):
        specs = make_specs(count, size=size, first_seed=3)
        steps = list(range(0, 18000, 1500))
        ens = Ensemble(specs, device=cuda_device, frame_steps=steps, output_size=osz)
        ens.simulate(); ens.synchronize()
        tables = P.prompt_tables(build_label_maps(ens.geometries[0], osz, cuda_device)[1], ens.geometries[0].n_cells)
        seen_bg = seen_quiet = 0
        for i in range(count):
            ca = ens.frames_of(i)[:, 0, :].cpu().numpy()
            a, b = np.random.RandomState(100 + i), np.random.RandomState(100 + i)
            want = []
            for f in range(len(steps)):
                if not (ca[f] > 0.1).any():
                    want.append(None); seen_quiet += 1
                    continue
                seen_bg += bool(ca[f][0] > 0.1)
                want.append(P.frame_prompts(tables, ca[f], rng=a, as_text=True))
            nent, texts = P.pouch_prompts(tables, ca, b, return_text=True)
            for f in range(len(steps)):
                if want[f] is None:
                    assert nent[f] == -1 and texts[f] is None
                else:
                    assert texts[f] == want[f][0], (size, i, f)
                    assert nent[f] == want[f][1]
            sa, sb = a.get_state(), b.get_state()
            assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1]), (size, i)
        assert seen_bg > 0


def test_legacy_choice_generic_path_without_avx512():
    """The block compaction of cab200_legacy_choice has an AVX-512 and a generic implementation, chosen at run time;
    the numpy-equality tests above run once more in a process where the wide one is switched off."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_prompts.py"), "-q", "-x", "-k",
                          "test_legacy_choice_is_numpys_choice or test_sample_background_is_the_numpy"],
                         capture_output=True, text=True, env=dict(os.environ, CAB200_NO_AVX512="1"), cwd=root, timeout=600)
    assert res.returncode == 0 and "2 passed" in res.stdout, res.stdout[-1500:] + res.stderr[-500:]
