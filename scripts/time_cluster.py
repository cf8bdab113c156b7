"""Throughput / latency of the cluster path vs one CTA per pouch (GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
dev = torch.device("cuda:0")
def run(size, P, cluster, T=18000):
    specs = make_specs(P, size=size)
    ens = Ensemble(specs, device=dev, T=T, cluster=cluster)
    ens.simulate(); ens.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.simulate(upload=False); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{size:7s} P={P:4d} cluster={cluster}: {best:8.2f} ms  {ens.cell_timesteps/best*1e3:.3e} cell-steps/s  status={int(ens.check_status().max())}", flush=True)
from calcium_b200 import _lib
from calcium_b200.geometry import get_geometry
lib = _lib.load()
for size in ("medium", "large", "xlarge"):
    g = get_geometry(size); mr = int(np.diff(g.csr[0]).max())
    print(size, g.n_cells, "cells: resident clusters", {c: lib.cab200_cluster_capacity(g.n_cells, mr, c, 64) for c in (2, 3, 4, 8)}, flush=True)
for cl in (1, 3, 4):
    run("large", 148, cl)
for P, cl in ((32, 4), (34, 4), (36, 4), (37, 4), (48, 3), (49, 3), (144, 3), (96, 3), (16, 8), (18, 8)):
    run("large", P, cl)
run("large", 1, 3)
run("xlarge", 1, 8); run("xlarge", 18, 8); run("xlarge", 74, 8)
run("large", 1, 1); run("large", 1, 4)
for cl in (1, 2, 4):
    run("medium", 1, cl)
run("medium", 74, 2); run("medium", 148, 1)
