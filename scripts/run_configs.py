"""BASELINE.json configs 2, 4, 5 on one GPU: parity spot checks + timings + the FP32 error report.
Writes gpurun_out/configs_r1.json (copied to profiles/ by hand)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs
from calcium_b200.parallel import shard_indices
from oracle import pouch_oracle as O
dev = torch.device("cuda:0")
out = {}

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

def oracle_frames(s, T, steps, f32=False):
    z = np.zeros(s.geometry.n_cells)
    return O.simulate_c(s.geometry.csr, O.pack_params(s.params), z, z, s.r0, s.vplc, T, steps, f32=f32)

# ---- config 2: 64 small pouches, all four sim types, FP64, one GPU -------------------------------
specs = make_specs(64, size="small")
ens = Ensemble(specs, device=dev)
ms = timed(lambda: ens.simulate(upload=False) if ens._d_in is not None else ens.simulate())
bad = sum(not np.array_equal(oracle_frames(specs[i], 18000, ens.frame_steps), ens.frames_of(i).cpu().numpy()) for i in range(0, 64, 7))
out["config2_64_small"] = {"k1_ms": ms, "cell_steps_per_s": ens.cell_timesteps / ms * 1e3, "bit_exact_checked": 10, "mismatches": bad}
print(out["config2_64_small"], flush=True)

# ---- config 4: one large pouch, 10x T, every step captured ------------------------------------------
T = 180000
spec = make_specs(1, size="large", sim_types=["Intercellular waves"])[0]
ens = Ensemble([spec], device=dev, T=T, frame_steps=range(T))
ms = timed(lambda: ens.simulate(upload=False) if ens._d_in is not None else ens.simulate(), reps=2)
fr = ens.frames_of(0)
gb = fr.numel() * 8 / 1e9
# parity on the first 4000 steps (the oracle needs ~10 s per 4000 large steps); the rest by the conservation property
want = oracle_frames(spec, 4000, list(range(0, 4000, 97)))
got = fr[0:4000:97].cpu().numpy()
cons = bool(torch.equal(fr[:, 2], (torch.tensor(spec.params["c_tot"], dtype=torch.float64, device=dev) - fr[:, 0]) / torch.tensor(spec.params["beta"], dtype=torch.float64, device=dev)))  # tensor divisor: torch turns division by a Python scalar into a multiply by 1/b
out["config4_large_10xT_every_step"] = {"k1_ms": ms, "cell_steps_per_s": ens.cell_timesteps / ms * 1e3, "frames_GB": gb,
                                        "hbm_write_GBps": gb / (ms * 1e-3), "bit_exact_first_4000_steps": bool(np.array_equal(want, got)),
                                        "conservation_all_steps": cons, "finite": bool(torch.isfinite(fr).all())}
print(out["config4_large_10xT_every_step"], flush=True)
del ens, fr
torch.cuda.empty_cache()

# ---- config 5: 256 pouches, sizes xsmall->large round-robin, FP32 vs FP64 -----------------------------
sizes = ["xsmall", "small", "medium", "large"]
specs = [make_specs(1, size=sizes[i % 4], first_seed=i, sim_types=[["Single cell spikes", "Intercellular transients", "Intercellular waves", "Fluttering"][(i // 4) % 4]])[0] for i in range(256)]
res = {}
for prec, name in ((_lib.FP64_STRICT, "fp64"), (_lib.FP32, "fp32")):
    ens = Ensemble(specs, device=dev, precision=prec)
    ms = timed(lambda: ens.simulate(upload=False) if ens._d_in is not None else ens.simulate())
    res[name] = (ens, ms)
e64, e32 = res["fp64"][0], res["fp32"][0]
errs = {}
for i, s in enumerate(specs):
    a = e64.frames_of(i).cpu().numpy(); b = e32.frames_of(i).cpu().numpy().astype(np.float64)
    d = np.abs(a - b)
    key = s.geometry.size
    e = errs.setdefault(key, {"max_abs_ca": 0.0, "rms_ca": [], "active_set_flips": 0, "frames": 0})
    e["max_abs_ca"] = max(e["max_abs_ca"], float(d[:, 0].max())); e["rms_ca"].append(float(np.sqrt((d[:, 0] ** 2).mean())))
    e["active_set_flips"] += int(((a[:, 0] > 0.1) != (b[:, 0] > 0.1)).sum()); e["frames"] += a.shape[0]
for e in errs.values():
    e["rms_ca"] = float(np.mean(e["rms_ca"]))
parts = shard_indices([s.geometry.n_cells for s in specs], 18000, 8)
loads = [sum(specs[i].geometry.n_cells for i in p) for p in parts]
out["config5_mixed_256"] = {"fp64_k1_ms": res["fp64"][1], "fp32_k1_ms": res["fp32"][1],
                            "fp64_cell_steps_per_s": e64.cell_timesteps / res["fp64"][1] * 1e3,
                            "fp32_cell_steps_per_s": e32.cell_timesteps / res["fp32"][1] * 1e3,
                            "fp32_vs_fp64_error_by_size": errs,
                            "lpt_8_rank_cell_loads": loads, "lpt_imbalance": max(loads) / (sum(loads) / 8)}
print(json.dumps(out["config5_mixed_256"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs_r1.json", "w"), indent=1)
