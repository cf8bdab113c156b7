"""Times K2b (cab200_raster_edges_batch) against K2 on one 37-pouch chunk of the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from calcium_b200.ensemble import Ensemble, make_specs

dev = torch.device("cuda:0")
P = 37
ens = Ensemble(make_specs(P, size="medium"), device=dev)
ens.simulate()
out = {}
def timed(label, **kw):
    for _ in range(2):
        ens.rasterize(image=True, instance=False, out=out, **kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ens.rasterize(image=True, instance=False, out=out, **kw)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    frames = P * ens.F
    print(f"{label:28s} {ms:8.3f} ms per {frames} frames  ({ms * 1e3 / frames:.2f} us/frame, "
          f"{frames * 512 * 512 * 2 / ms / 1e6:.0f} GB/s of image written)")
timed("K2 image only")
for ks, bt in ((3, "mean"), (5, "motion"), (5, "mean"), (9, "mean")):
    timed(f"K2b edge_blur {bt}{ks}", edge_blur=True, blur_kernel_size=ks, blur_type=bt)
timed("K2b blur+border", edge_blur=True, with_border=True)
timed("K2b border", with_border=True)
