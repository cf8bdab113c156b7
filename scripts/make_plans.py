"""Offline: anneal K1 layout plans for the synthetic geometries and store them under
calcium_b200/plans/ (tracked in git; tiny).  usage: python scripts/make_plans.py SIZE ITERS [N_CTA]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import cluster_layout, plan_layout, plan_path, slot_plan
from calcium_b200.geometry import get_geometry
size, iters = sys.argv[1], int(sys.argv[2])
n_cta = int(sys.argv[3]) if len(sys.argv) > 3 else 1
g = get_geometry(size); n = g.n_cells
copies, cpt = plan_layout(n, True, 64) if n_cta == 1 else cluster_layout(n, n_cta, 64)
t = time.time()
plan = slot_plan(g, True, 64, iters=iters, n_cta=n_cta)
path = plan_path(g, copies, cpt, True, n_cta)
np.save(path, plan)
print(size, "n_cta", n_cta, "copies", copies, "cpt", cpt, "iters", iters, "len", len(plan), f"{time.time()-t:.0f}s ->", path)
