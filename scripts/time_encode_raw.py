"""K3r alone: device time of cab200_encode_raw_batch for one 37-pouch chunk of edge-blurred 512 x 512 frames, bytes
out, and the copy to pinned host memory.  python scripts/time_encode_raw.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs, encode_raw_frames, _stream_ptr
dev = torch.device("cuda:0")
ens = Ensemble(make_specs(37, size="medium"), device=dev)
ens.simulate()
for name, kw in (("clean (flat-shaded)", {}), ("edge_blur mean 3", dict(edge_blur=True, blur_kernel_size=3, blur_type="mean"))):
    img = ens.rasterize(image=True, instance=False, **kw)["image"].clone()
    lib = _lib.load()
    n, h, w = img.shape[0] * img.shape[1], img.shape[2], img.shape[3]
    n_slots = 2 * torch.cuda.get_device_properties(dev).multi_processor_count
    nb = int(lib.cab200_encode_raw_scratch_bytes(h, w, n_slots))
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev); records = torch.empty((n, 2), dtype=torch.int64, device=dev)
    cursor = torch.zeros(2, dtype=torch.int64, device=dev); cap = n * h * w * 2
    arena = torch.empty(cap, dtype=torch.uint8, device=dev)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.cab200_encode_raw_batch(img.data_ptr(), n, h, w, arena.data_ptr(), cap, cursor.data_ptr(), records.data_ptr(),
                                               scratch.data_ptr(), nb, n_slots, _stream_ptr(dev)), "raw")
        e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    used = int(cursor[0].item())
    t0 = time.perf_counter(); a, off, size, fl = encode_raw_frames(img, keep_pinned=True); t1 = time.perf_counter()
    print(f"{name:22s}: {n} frames, K3r {best:7.3f} ms ({best / n * 1e3:.2f} us per frame), {used / 1e6:.0f} MB = {used / n / 1e3:.0f} KB per frame "
          f"({used / (n * h * w):.2f} B per pixel); encode_raw_frames incl. copy-out {1e3 * (t1 - t0):.1f} ms", flush=True)
