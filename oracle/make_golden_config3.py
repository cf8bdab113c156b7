"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/traj_config3_head.npz``.

Run in the build container (``python -m oracle.make_golden_config3``), where ``/root/reference`` is mounted:
the first pouches of BASELINE config 3 / bench.py's workload exactly as ``calcium_b200.ensemble.make_specs``
enumerates them (medium, pouch i: sim type ``TYPES[i % 4]``, ``generate_random_params(seed=i)``,
``sim_number=i``), integrated by the *unmodified* reference (numba ``fastmath`` step, scipy CSR products,
``core/pouch.py:289-372``) for the full T = 18000 steps.  Recorded per pouch: the reference's own V_PLC and
r0 draws and the state (c, p, s, r) at four steps as float64, plus -- for all 90 sampled frames of
``main.py:167`` -- the active-cell counts ``len(get_active_cells(t))``.  Also one large pouch (3255 cells).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_shims as R  # noqa: E402
from oracle.make_golden import TYPES, frames_of, geometry_digest  # noqa: E402
from calcium_b200 import geometry as G   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_POUCH = 8
STEPS = [0, 6000, 12000, 17999]


def main() -> None:
    ns = R.load_reference()
    gdir = R.synthetic_geometry_dir(sizes=("xsmall", "small", "medium", "large"))
    z = {"steps": np.array(STEPS), "geometry_sha256": np.array(geometry_digest(G.get_geometry("medium"))),
         "sampled": np.arange(0, 18000, 200)}
    for i in range(N_POUCH):
        st = TYPES[i % 4]
        prm = ns.SimulationParameters(st).generate_random_params(seed=i)
        po = ns.Pouch(params=prm, size="medium", sim_number=i, geometry_dir=gdir)
        po.simulate()
        z[f"vplc_{i}"] = po.V_PLC.flatten()
        z[f"r0_{i}"] = po.disc_dynamics[:, 3, 0].copy()
        z[f"frames_{i}"] = frames_of(po, STEPS)
        z[f"n_active_{i}"] = np.array([len(po.get_active_cells(int(t))) for t in z["sampled"]], dtype=np.int32)
        print("medium", i, st, "max c", float(po.disc_dynamics[:, 0, :].max()), flush=True)
        del po
    prm = ns.SimulationParameters("Intercellular waves").generate_random_params(seed=2)
    po = ns.Pouch(params=prm, size="large", sim_number=2, geometry_dir=gdir)
    po.simulate()
    z["large_geometry_sha256"] = np.array(geometry_digest(G.get_geometry("large")))
    z["large_vplc"] = po.V_PLC.flatten()
    z["large_r0"] = po.disc_dynamics[:, 3, 0].copy()
    z["large_frames"] = frames_of(po, [9000, 17999])
    z["large_steps"] = np.array([9000, 17999])
    np.savez_compressed(os.path.join(OUT, "traj_config3_head.npz"), **z)
    print("wrote", os.path.join(OUT, "traj_config3_head.npz"), os.path.getsize(os.path.join(OUT, "traj_config3_head.npz")))


if __name__ == "__main__":
    main()
