"""BASELINE config 4 at full occupancy: large pouches as 4-CTA clusters, as many as are resident at once (33 on a
148-SM B200: a cluster lives inside one GPC; 37 = 148 // 4 would run as two waves), the
10x horizon (180 000 states) with EVERY step of all four states captured.  One pouch's history is 18.7 GB, so the
horizon runs in segments whose frames fit HBM next to the slot-ordered staging copy; each segment continues from the
last captured state (Ensemble.continue_from_last_frame, bit-identical to one long run) and its frames are consumed
(here: overwritten by the next segment) -- the "ring" of SURVEY section 7.  Reports the frame-write rate against the
measured HBM peak, and checks pouch 0 bit for bit against the oracle on the first and on the LAST `CHECK` steps of
the horizon (the oracle is re-seeded from the captured state at step T_total-1-CHECK).
python scripts/config4_full.py > gpurun_out/r2_config4_full.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calcium_b200.ensemble import Ensemble, make_specs
from oracle import pouch_oracle as O
O.build()
dev = torch.device("cuda:0")
from calcium_b200 import _lib
from calcium_b200.geometry import get_geometry
_g = get_geometry("large")
CAPACITY = int(_lib.load().cab200_cluster_capacity(_g.n_cells, int(np.diff(_g.csr[0]).max()), 4, 64))   # resident clusters of 4: 33 on B200
P = int(os.environ.get("P", str(CAPACITY)))
T_TOTAL = int(os.environ.get("T_TOTAL", "180000"))
SEG = int(os.environ.get("SEG", "9000"))            # Euler steps per segment (SEG + 1 captured states)
CHECK = int(os.environ.get("CHECK", "4000"))
peak = 6540.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
specs = make_specs(P, size="large", sim_types=["Intercellular waves"])
n = specs[0].geometry.n_cells
steps_left = T_TOTAL - 1
segs = []
while steps_left > 0:
    k = min(SEG, steps_left); segs.append(k); steps_left -= k
ens = None
seg_ms, first, last_states, done = [], None, {}, 0
seed_state = None
for si, k in enumerate(segs):
    T = k + 1
    if ens is None or ens.T != T:
        prev = ens
        ens = Ensemble(specs, device=dev, T=T, frame_steps=range(T), cluster=4, coalesced_capture=True)
        if prev is not None:                       # the shorter last segment: carry the device state over
            ens.upload(); ens._d_in.copy_(prev._d_in); del prev; torch.cuda.empty_cache()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ens.simulate(upload=(si == 0)); e1.record(); e1.synchronize()
    seg_ms.append(e0.elapsed_time(e1))
    assert ens._frames_tmp is not None, "coalesced capture staging did not fit"
    fr = ens.frames_of(0)                          # [T, 4, n] of pouch 0, global steps done .. done + k
    if si == 0:
        first = fr[: CHECK + 1].cpu().numpy().copy()
    lo = T_TOTAL - 1 - CHECK                       # global step the oracle is re-seeded from
    a, b = max(lo, done), done + k
    if b >= lo:
        last_states[(a, b)] = fr[a - done: b - done + 1].cpu().numpy().copy()
    status = int(ens.check_status().max())
    done += k
    if si + 1 < len(segs):
        ens.continue_from_last_frame()
torch.cuda.synchronize()
# ---- parity, pouch 0 --------------------------------------------------------------------------------
sp = specs[0]; g = sp.geometry; z = np.zeros(n)
want_first = O.simulate_c(g.csr, O.pack_params(sp.params), z, z, sp.r0, sp.vplc, CHECK + 1, list(range(CHECK + 1)))
ok_first = bool(np.array_equal(want_first, first))
tail = np.concatenate([v if i == 0 else v[1:] for i, (kk, v) in enumerate(sorted(last_states.items()))])
seed = tail[0]                                     # captured state at global step T_TOTAL-1-CHECK
want_last = O.simulate_c(g.csr, O.pack_params(sp.params), seed[0], seed[1], seed[3], sp.vplc, CHECK + 1, list(range(CHECK + 1)))
ok_last = bool(np.array_equal(want_last, tail)) and tail.shape[0] == CHECK + 1
total_ms = float(sum(seg_ms))
frame_bytes = P * (T_TOTAL - 1 + len(segs)) * 4 * n * 8
out = {"resident_clusters_of_4": CAPACITY,
       "config": f"{P} large pouches (3255 cells) as 4-CTA clusters = {4 * P} CTAs, {T_TOTAL} states (10 x T), every step of c, p, s, r captured; "
                 f"{len(segs)} segments of <= {SEG} steps continued on the device",
       "ms_total": total_ms, "ms_per_segment": seg_ms, "cell_timesteps_per_s": P * n * (T_TOTAL - 1) / total_ms * 1e3,
       "us_per_step": total_ms * 1e3 / (T_TOTAL - 1),
       "frames_GB_written_algorithmic": frame_bytes / 1e9, "frame_write_GBps": frame_bytes / 1e9 / (total_ms * 1e-3),
       "hbm_peak_GBps": peak, "frac_of_hbm_peak_algorithmic": frame_bytes / 1e9 / (total_ms * 1e-3) / peak,
       "hbm_traffic_note": "the slot-ordered staging copy makes the real traffic 2 writes + 1 read per captured byte (staging store, unpermute read, cell-order store): multiply by 3 for DRAM bytes moved",
       "frac_of_hbm_peak_traffic": 3 * frame_bytes / 1e9 / (total_ms * 1e-3) / peak,
       "bit_exact_first_steps": ok_first, "bit_exact_last_steps": ok_last, "checked_steps_each": CHECK,
       "status_nonzero": status, "frames_tensor_GB": ens.frames.numel() * 8 / 1e9, "staging_GB": ens._frames_tmp.numel() / 1e9}
print(json.dumps(out, indent=1))
