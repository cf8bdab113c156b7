"""Workload for the -DCAB_K1_CHECK build of K1 (calcium_b200/build.py: build_k1_check): single-CTA and cluster
ensembles, results compared bit for bit with the oracle and the CAB200_ST_PROTOCOL status bit reported.  Run with
CALCIUM_B200_LIB=calcium_b200/_variants/libk1check.so; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs
from oracle import pouch_oracle as O
O.build()
dev = torch.device("cuda:0")
out = {"lib": _lib.LIB_PATH, "cases": []}
for size, cluster, count, T in (("medium", 1, 150, 3000), ("small", 1, 40, 3000), ("medium", 2, 5, 3000), ("medium", 4, 5, 3000),
                                ("large", 4, 40, 2500), ("xlarge", 8, 20, 2000)):
    specs = make_specs(count, size=size, first_seed=7)
    steps = [0, T // 2, T - 1]
    ens = Ensemble(specs, T=T, frame_steps=steps, device=dev, cluster=cluster)
    ens.simulate(); ens.synchronize()
    st = ens.status.cpu().numpy()
    s = specs[count // 2]; g = s.geometry; z = np.zeros(g.n_cells)
    want = O.simulate_c(g.csr, O.pack_params(s.params), z, z, s.r0, s.vplc, T, steps)
    out["cases"].append({"size": size, "cluster": cluster, "pouches": count, "T": T, "status_or": int(np.bitwise_or.reduce(st)),
                         "protocol_violation": bool((st & _lib.ST_PROTOCOL).any()),
                         "bit_exact": bool(np.array_equal(want, ens.frames_of(count // 2).cpu().numpy()))})
print(json.dumps(out))
