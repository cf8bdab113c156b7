"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz|json``.

Run in the build container (``python -m oracle.make_golden``), where ``/root/reference`` is
mounted: imports the *unmodified* reference with the shims of ``oracle/reference_shims.py`` and
records what the reference itself computes, so that the oracle (and through it the CUDA path) is
pinned to the reference rather than to a restatement.  The fixtures are committed; the GPU box
never sees ``/root/reference``.

Fixtures
  params.json          generate_random_params(seed) for 4 sim types x seeds 0..7  (core/parameters.py:222-282)
  step_kat.npz         _simulate_time_step on hand-made 1/2/5-cell inputs: numba-jit and py_func outputs (core/pouch.py:31-104)
  traj_xsmall_waves.npz   BASELINE config 1: xsmall, default "Intercellular waves", sim_number 0, full T;
                          V_PLC / r0 as drawn by the reference, frames at 12 steps, jit and py_func
  traj_small_types.npz    small pouch, the four sim types, generate_random_params(seed=i), sim_number=i (jit)
  traj_medium_flutter.npz medium pouch, "Fluttering" seed 3 (worst conditioning seen in SURVEY.md section 7), 3 frames (jit)
  raster_xsmall.npz       cells_mask, generate_image(t), get_active_cells, get_cell_masks(active_only), instance masks
                          at 160x120 and 512x512, from the reference's own methods on the xsmall run
  geometry_xsmall.npz     the synthetic xsmall geometry itself (drift check for the generator)
  prompts_small.npz/json  generate_point_prompts_from_mask + generate_negative_prompts (utils/sam2_data_generator.py)
                          on instance masks of a small "Fluttering" pouch (reference pipeline, 256x256), with
                          np.random.seed(k) in front of each frame; the calcium frames are stored as inputs
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_shims as R  # noqa: E402
from calcium_b200 import geometry as G   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TYPES = ["Single cell spikes", "Intercellular transients", "Intercellular waves", "Fluttering"]


def geometry_digest(g) -> str:
    h = hashlib.sha256()
    h.update(np.concatenate(list(g.vertices)).astype(np.float64).tobytes())
    h.update(np.asarray(g.edges, dtype=np.int32).tobytes())
    return h.hexdigest()


def frames_of(pouch, steps):
    return np.ascontiguousarray(np.transpose(pouch.disc_dynamics[:, :, steps], (2, 1, 0)))


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    ns = R.load_reference()
    gdir = R.synthetic_geometry_dir()

    # ---- parameters -------------------------------------------------------------------
    pj = {}
    for st in TYPES:
        pj[st] = {"default": ns.SimulationParameters(st).get_params_dict(),
                  "seeds": {str(s): {k: float(v) for k, v in
                                     ns.SimulationParameters(st).generate_random_params(seed=s).items()}
                            for s in range(8)}}
    with open(os.path.join(OUT, "params.json"), "w") as fh:
        json.dump(pj, fh, indent=1, sort_keys=True)

    # ---- step known-answer tests --------------------------------------------------------
    rng = np.random.RandomState(7)
    kat = {}
    scal = dict(dt=0.2, k_1=1.11, k_a=0.08, k_p=0.13, k_2=0.0203, V_SERCA=0.9, K_SERCA=0.1,
                K_PLC=0.2, K_5=0.66, c_tot=2.0, beta=0.185, k_tau=1.5, tau_max=800.0, k_i=0.4)
    for n in (1, 2, 5, 64):
        ca = rng.uniform(0, 1.3, (n, 1)); ipt = rng.uniform(0, 2.0, (n, 1))
        s = (2.0 - ca) / 0.185; r = rng.uniform(0.5, 0.9, (n, 1))
        lc = rng.uniform(-1e-3, 1e-3, (n, 1)); lp = rng.uniform(-1e-2, 1e-2, (n, 1))
        v = rng.uniform(0.1, 1.5, n)
        args = (ca, ipt, s, r, lc, lp, v)
        jit = ns.step_jit(*args, *scal.values())
        pyf = ns.step_jit.py_func(*args, *scal.values())
        kat[f"in_{n}"] = np.concatenate([a.reshape(n, -1) for a in args], axis=1)
        kat[f"jit_{n}"] = np.concatenate(jit, axis=1)
        kat[f"pyf_{n}"] = np.concatenate(pyf, axis=1)
    kat["scalars"] = np.array(list(scal.values()))
    np.savez(os.path.join(OUT, "step_kat.npz"), **kat)

    # ---- config 1: xsmall default waves ------------------------------------------------------
    steps = [0, 1, 2, 200, 1000, 2000, 4000, 8000, 10000, 12000, 16000, 17999]
    p = ns.SimulationParameters("Intercellular waves").get_params_dict()
    pouch = ns.Pouch(params=p, size="xsmall", sim_number=0, geometry_dir=gdir)
    pouch.simulate()
    with R.use_py_func():
        pouch_py = ns.Pouch(params=p, size="xsmall", sim_number=0, geometry_dir=gdir)
        pouch_py.simulate()
    g = G.get_geometry("xsmall")
    np.savez(os.path.join(OUT, "traj_xsmall_waves.npz"),
             steps=np.array(steps), vplc=pouch.V_PLC.flatten(), r0=pouch.disc_dynamics[:, 3, 0],
             frames_jit=frames_of(pouch, steps), frames_pyfunc=frames_of(pouch_py, steps),
             geometry_sha256=np.array(geometry_digest(g)))
    offs = np.cumsum([0] + [len(v) for v in g.vertices])
    np.savez(os.path.join(OUT, "geometry_xsmall.npz"), offs=offs,
             flat=np.concatenate(list(g.vertices)), edges=g.edges)

    # ---- raster fixtures from the same run -----------------------------------------------
    rz = {}
    for (w, h) in ((160, 120), (512, 512)):
        pr = ns.Pouch(params=p, size="xsmall", sim_number=0, geometry_dir=gdir, output_size=(w, h))
        pr.disc_dynamics = pouch.disc_dynamics        # reuse the trajectory
        tag = f"{w}x{h}"
        rz[f"cells_mask_{tag}"] = pr.cells_mask.astype(np.uint16)
        ts = [0, 1000, 4000, 10000, 17999]
        rz[f"t_{tag}"] = np.array(ts)
        for t in ts:
            rz[f"image_{tag}_{t}"] = pr.generate_image(t)
            act = pr.get_active_cells(t)
            rz[f"active_{tag}_{t}"] = np.array(act, dtype=np.int32)
            am = pr.get_cell_masks(active_only=True, time_step=t)
            rz[f"activemask_{tag}_{t}"] = am.astype(np.uint16)
            rz[f"instance_{tag}_{t}"] = ns.create_instance_mask_from_cells(am, act)
    pr = ns.Pouch(params=p, size="xsmall", sim_number=0, geometry_dir=gdir, output_size=(160, 120))
    pr.disc_dynamics = pouch.disc_dynamics
    rz["edge_mean3_160x120_10000"] = pr.generate_image(10000, edge_blur=True, blur_kernel_size=3, blur_type="mean")
    rz["edge_motion5_160x120_10000"] = pr.generate_image(10000, edge_blur=True, blur_kernel_size=5, blur_type="motion")
    # with_border=True alone raises in the reference under OpenCV 4.13 (cv2.polylines on the
    # non-contiguous view img_data[:, :, 0], core/pouch.py:481) -> no fixture for that branch.
    rz["border_blur_160x120_10000"] = pr.generate_image(10000, with_border=True, edge_blur=True)
    rz["ca"] = np.ascontiguousarray(pouch.disc_dynamics[:, 0, [0, 1000, 4000, 10000, 17999]].T)
    np.savez_compressed(os.path.join(OUT, "raster_xsmall.npz"), **rz)

    # ---- small pouches, four types, randomised -------------------------------------------
    steps_s = [0, 3000, 9000, 17999]
    sm = {"steps": np.array(steps_s)}
    gs = G.get_geometry("small")
    sm["geometry_sha256"] = np.array(geometry_digest(gs))
    for i, st in enumerate(TYPES):
        prm = ns.SimulationParameters(st).generate_random_params(seed=i)
        po = ns.Pouch(params=prm, size="small", sim_number=i, geometry_dir=gdir)
        po.simulate()
        sm[f"vplc_{i}"] = po.V_PLC.flatten()
        sm[f"r0_{i}"] = po.disc_dynamics[:, 3, 0].copy()
        sm[f"frames_{i}"] = frames_of(po, steps_s)
        del po
    np.savez(os.path.join(OUT, "traj_small_types.npz"), **sm)

    # ---- medium fluttering ----------------------------------------------------------------
    steps_m = [6000, 12000, 17999]
    prm = ns.SimulationParameters("Fluttering").generate_random_params(seed=3)
    po = ns.Pouch(params=prm, size="medium", sim_number=3, geometry_dir=gdir)
    po.simulate()
    gm = G.get_geometry("medium")
    np.savez(os.path.join(OUT, "traj_medium_flutter.npz"), steps=np.array(steps_m),
             vplc=po.V_PLC.flatten(), r0=po.disc_dynamics[:, 3, 0].copy(),
             frames=frames_of(po, steps_m), geometry_sha256=np.array(geometry_digest(gm)))
    # ---- prompts: the reference's own functions on its own instance masks ------------------
    from utils.sam2_data_generator import generate_point_prompts_from_mask, generate_negative_prompts
    prm = ns.SimulationParameters("Fluttering").generate_random_params(seed=3)
    pf = ns.Pouch(params=prm, size="small", sim_number=3, geometry_dir=gdir, output_size=(256, 256))
    pf.simulate()
    ts = [2000, 9000, 16000]
    pj2, pz = {}, {"t": np.array(ts), "ca": np.ascontiguousarray(pf.disc_dynamics[:, 0, ts].T),
                   "cells_mask": pf.cells_mask.astype(np.uint16)}
    for k, t in enumerate(ts):
        # threshold 0.1 is the default; 0.0 makes (almost) every cell active, cell 0 included: the id
        # mismatch then labels the whole background as instance 1
        for thr in (0.1, 0.0):
            act = pf.get_active_cells(t, threshold=thr)
            am = pf.get_cell_masks(active_only=True, time_step=t, threshold=thr)
            inst = ns.create_instance_mask_from_cells(am, act)
            np.random.seed(40 + k)
            pos = generate_point_prompts_from_mask(inst, strategy="distance_based")
            neg = generate_negative_prompts(inst, num_negative=5)
            pj2[f"{t}_{thr}"] = {"positive": {str(i): v for i, v in pos.items()}, "negative": neg,
                                  "n_active": len(act), "seed": 40 + k}
            pz[f"instance_{t}_{thr}"] = inst
    np.savez_compressed(os.path.join(OUT, "prompts_small.npz"), **pz)
    with open(os.path.join(OUT, "prompts_small.json"), "w") as fh:
        json.dump(pj2, fh, indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
