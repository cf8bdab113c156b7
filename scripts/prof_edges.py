"""Short K2b run for ncu: one 37-pouch chunk of medium 512x512 frames, edge_blur mean 3x3 (one warm-up + one launch)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
dev = torch.device("cuda:0")
ens = Ensemble(make_specs(37, size="medium"), device=dev)
ens.simulate()
out = {}
for _ in range(2):
    ens.rasterize(image=True, instance=False, out=out, edge_blur=True, blur_kernel_size=3, blur_type="mean")
ens.synchronize()
print("ok")
