"""K3 timing on the benchmark workload: chunks of medium pouches, 90 frames each (GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 37
ens = Ensemble(make_specs(P, size="medium"), device=dev)
ens.simulate(); ens.synchronize()
out, enc = {}, {}
for name, fn in (("K2 raw", lambda: ens.rasterize(image=True, instance=True, out=out)),
                 ("K3 img+mask", lambda: ens.encode(out=enc)),
                 ("K3 per-pixel", lambda: ens.encode(out=enc, parse="pixels")),
                 ("K3 img only", lambda: ens.encode(mask=False, out=enc)),
                 ("K3 mask only", lambda: ens.encode(image=False, out=enc))):
    fn(); ens.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    extra = ""
    if name.startswith("K3"):
        arena, off, size, flags = ens.encoded_files(enc)
        extra = (f" used={int(enc['cursor'].item())/1e6:.1f} MB  mean png={size[...,0][size[...,0]>0].mean() if (size[...,0]>0).any() else 0:.0f}"
                 f" mean npz={size[...,1][size[...,1]>0].mean() if (size[...,1]>0).any() else 0:.0f} max={size.max()} flags={np.unique(flags).tolist()}")
    print(f"{name:14s} P={P} frames={P*ens.F}: {best:8.3f} ms  ({best/(P*ens.F)*1e3:.2f} us/frame){extra}", flush=True)
