"""Shared-memory pipe micro-benchmarks (cab200_peak_smem): the measured denominator of K1's roofline and the
answer to "why no warp shuffles for the neighbour sums".  python scripts/microbench_smem.py > profiles/..."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calcium_b200 import _lib
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device="cuda")
names = ["LDS.128 conflict-free", "LDS.128 all lanes one address", "LDS.128 one address per 8-lane phase",
         "LDS.128 half the lanes on one padding address", "LDS.64 conflict-free", "SHFL.IDX 32-bit", "STS.128 conflict-free",
         "LDS.128 2-way bank-group conflict"]
res = {}
for mode, name in enumerate(names):
    out = np.zeros(3)
    _lib.check(lib.cab200_peak_smem(mode, 20000, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None), "cab200_peak_smem")
    res[name] = {"bytes_per_clk_per_sm": round(out[0], 2), "device_GBps": round(out[1], 1), "clk_per_warp_access": round(out[2], 3)}
    print(f"{name:48s} {out[0]:7.2f} B/clk/SM  {out[1]:9.1f} GB/s  {out[2]:6.3f} clk per warp-wide access", flush=True)
print(json.dumps(res))
