"""Offline: anneal K1 slot plans for the synthetic geometries and store them under
calcium_b200/plans/ (tracked in git; tiny).  usage: python scripts/make_plans.py SIZE ITERS"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import _lib
from calcium_b200.ensemble import plan_path, plan_layout
from calcium_b200.geometry import get_geometry
size, iters = sys.argv[1], int(sys.argv[2])
lib = _lib.load()
g = get_geometry(size); rowptr, col, _ = g.csr; n = g.n_cells
copies, cpt = plan_layout(n, True, 64)
plan = np.zeros(2 * n, dtype=np.uint16); cost = np.zeros(4, dtype=np.int64)
t = time.time()
_lib.check(lib.cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, cpt, iters, 1, 0,
                                 plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), _lib.as_i64p(cost)), "plan")
path = plan_path(g, copies, cpt, tracked=True)
np.save(path, plan)
print(size, "copies", copies, "cpt", cpt, "iters", iters, "cost", cost.tolist(), f"{time.time()-t:.0f}s ->", path)
