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
for cl in (1, 4):
    run("large", 148, cl)
run("large", 37, 4)
run("large", 1, 1); run("large", 1, 4)
for cl in (1, 2, 4):
    run("medium", 1, cl)
run("medium", 74, 2); run("medium", 148, 1)
