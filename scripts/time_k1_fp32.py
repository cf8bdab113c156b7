"""K1 in the FP32 mode on the benchmark wave (148 medium pouches, full T) beside FP64: ms per launch (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs
dev = torch.device("cuda:0")
specs = make_specs(148, size="medium")
for name, prec in (("fp64", _lib.FP64_STRICT), ("fp32", _lib.FP32)):
    if os.environ.get("ONLY") and os.environ["ONLY"] != name: continue
    ens = Ensemble(specs, device=dev, precision=prec)
    ens.upload()
    for _ in range(2): ens.simulate(upload=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ens.simulate(upload=False); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    print(f"{name}: K1 {min(ts):7.3f} ms", flush=True)
