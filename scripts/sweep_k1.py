"""K1 shape sweep + bit-parity spot check + division self-test (GPU)."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs
from oracle import pouch_oracle as O
dev = torch.device("cuda:0"); lib = _lib.load()
mm = ctypes.c_ulonglong()
for er in (8, 60, 300, 900):
    _lib.check(lib.cab200_selftest_division(1 << 30, er, 42 + er, ctypes.byref(mm), None), "selftest")
    print(f"division self-test exp_range={er}: mismatches {mm.value} of {1<<30}")
size = os.environ.get("SWEEP_SIZE", "medium"); P = int(os.environ.get("SWEEP_P", "148")); T = int(os.environ.get("SWEEP_T", "18000"))
specs = make_specs(P, size=size)
shapes = [tuple(map(int, s.split("x"))) for s in os.environ.get("SWEEP_SHAPES", "512x3,768x2,1024x2").split(",")]
ref = None
import itertools
for (th, cpt), use_plan in itertools.product(shapes, (False, True)):
    _lib.check(lib.cab200_sim_set_plan(100000, th, cpt), "plan")
    ens = Ensemble(specs, device=dev, T=T, use_plan=use_plan)
    ens.simulate(); ens.synchronize()
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ens.simulate(upload=False); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    s = specs[3]; z = np.zeros(s.geometry.n_cells)
    if ref is None:
        ref = O.simulate_c(s.geometry.csr, O.pack_params(s.params), z, z, s.r0, s.vplc, T, ens.frame_steps)
    eq = np.array_equal(ref, ens.frames_of(3).cpu().numpy())
    print(f"K1 {size} P={P} T={T} shape {th}x{cpt} plan={use_plan}: {ms:.2f} ms  {ens.cell_timesteps/ms*1e3:.3e} cell-steps/s  bit-exact[3]={eq} status={int(ens.check_status().max())}")
    del ens
