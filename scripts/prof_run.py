"""Short K1+K2 run for ncu: medium ensemble, one launch each (plus one warm-up)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
P = int(os.environ.get("PROF_P", "148")); T = int(os.environ.get("PROF_T", "18000"))
size = os.environ.get("PROF_SIZE", "medium")
dev = torch.device("cuda:0")
ens = Ensemble(make_specs(P, size=size), device=dev, T=T)
for _ in range(2):
    ens.simulate()
    out = ens.rasterize()
ens.synchronize()
print("ok", ens.check_status().max())
