"""BASELINE config 4: large pouches, 10x T, every step captured -- direct vs coalesced capture (GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
dev = torch.device("cuda:0")
T = int(os.environ.get("T", "180000"))
for P, coal, cl in ((1, False, 4), (1, True, 4), (1, True, 1), (3, True, 4)):
    specs = make_specs(P, size="large", sim_types=["Intercellular waves"])
    ens = Ensemble(specs, device=dev, T=T, frame_steps=range(T), coalesced_capture=coal, cluster=cl)
    ens.simulate(); ens.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.simulate(upload=False); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    gb = ens.frames.numel() * 8 / 1e9
    print(f"large P={P} T={T} every step, cluster={cl} coalesced={coal}: {best:8.1f} ms  {ens.cell_timesteps/best*1e3:.3e} cell-steps/s  "
          f"frames {gb:.1f} GB -> {gb/best*1e3:.1f} GB/s written  status={int(ens.check_status().max())}", flush=True)
    del ens; torch.cuda.empty_cache()
