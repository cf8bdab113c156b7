"""Short K5 run for ncu: the full deterministic defect chain over one 37-pouch chunk of 512x512 frames."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200.ensemble import Ensemble, make_specs
from calcium_b200.defects import apply_defects
dev = torch.device("cuda:0")
ens = Ensemble(make_specs(37, size="medium"), device=dev)
ens.simulate()
out = ens.rasterize(image=True, instance=False)
cfg = {"background_fluorescence": True, "non_uniform_background": False, "background_intensity": 0.1, "vignetting": True,
       "vignetting_strength": 0.5, "dynamic_range_compression": True, "gamma": 0.8, "quantization": True, "bit_depth": 6,
       "adjust_brightness_contrast": True, "brightness": 0.02, "contrast": 1.1}
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
apply_defects(out["image"], cfg, (512, 512))
s.record()
apply_defects(out["image"], cfg, (512, 512))
e.record(); torch.cuda.synchronize()
print(f"K5 chain (max + 5 passes) over {37 * 90} frames: {s.elapsed_time(e):.3f} ms", flush=True)
