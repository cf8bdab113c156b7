"""Offline: long annealing runs of the K1 slot plan (cab200_plan_slots) from several seeds; keeps the best.
usage: python scripts/anneal_long.py SIZE ITERS SEED [polish]   -> prints costs, writes /tmp/plan_SIZE_SEED.npy"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import _lib
from calcium_b200.ensemble import plan_layout, slot_plan
from calcium_b200.geometry import get_geometry
size, iters, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
polish = len(sys.argv) > 4
g = get_geometry(size); n = g.n_cells
rowptr, col, _ = g.csr
copies, cpt = plan_layout(n, True, 64)
lib = _lib.load()
plan = slot_plan(g, True, 64).copy() if polish else np.zeros(2 * n, dtype=np.uint16)
cost = np.zeros(4, dtype=np.int64)
t = time.time()
_lib.check(lib.cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1, cpt, iters, seed, 1 if polish else 0,
                                 plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), _lib.as_i64p(cost)), "plan")
out = f"/tmp/plan_{size}_{seed}{'_p' if polish else ''}.npy"
np.save(out, plan)
print(size, "seed", seed, "iters", iters, "polish", polish, "cost0", cost[0], "final", cost[1], cost[2], "gather instr", cost[3],
      f"{time.time()-t:.0f}s", out, flush=True)
