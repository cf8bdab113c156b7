"""K1 alone on the benchmark wave (148 medium pouches, full T): ms per launch with the tracked plan, a quick plan
and no plan.  Used by scripts/k1_variants.py with CALCIUM_B200_LIB pointing at experimental builds."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs, plan_layout
P = int(os.environ.get("K1_P", "148")); size = os.environ.get("K1_SIZE", "medium"); T = int(os.environ.get("K1_T", "18000"))
dev = torch.device("cuda:0")
specs = make_specs(P, size=size)
def run(ens, tag):
    ens.upload()
    for _ in range(2): ens.simulate(upload=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ens.simulate(upload=False); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    print(f"{os.path.basename(_lib.LIB_PATH):28s} {tag:22s} K1 {min(ts):7.3f} ms  (status {int(ens.check_status().max())})", flush=True)
cl = os.environ.get("K1_CLUSTER", "auto")
ens = Ensemble(specs, T=T, device=dev, cluster=cl if cl == "auto" else int(cl))
run(ens, "tracked plan")
if os.environ.get("K1_QUICK", "1") == "1" and int(ens.g_cluster[0]) == 1:
    g = specs[0].geometry; rowptr, col, _ = g.csr; n = g.n_cells
    copies, cpt = plan_layout(n, True, 64)
    for iters in (300000, 0):
        plan = np.zeros(2 * n, dtype=np.uint16); cost = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.load().cab200_plan_slots(n, _lib.as_i32p(rowptr), _lib.as_i32p(col), copies, 1, cpt, iters, 1, 0,
                   plan.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), _lib.as_i64p(cost)), "plan")
        ens.d_plan = torch.from_numpy(plan.view(np.int16)).to(dev)
        run(ens, f"plan {iters} it ({int(cost[2])} wf)")
