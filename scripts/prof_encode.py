"""Short K3 run for ncu: one chunk of medium pouches (one warm-up launch, one profiled)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
P = int(os.environ.get("PROF_P", "37"))
dev = torch.device("cuda:0")
ens = Ensemble(make_specs(P, size="medium"), device=dev)
ens.simulate()
enc = {}
for _ in range(2):
    ens.encode(out=enc)
ens.synchronize()
print("ok", int(enc["cursor"].item()))
