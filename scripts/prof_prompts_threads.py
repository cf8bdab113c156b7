"""Where the prompt stage's wall time goes with many threads: per-section busy time summed over threads."""
import os, sys, time, threading, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, concurrent.futures as cf
from calcium_b200 import prompts as P
from calcium_b200.ensemble import Ensemble, make_specs, build_label_maps
dev = torch.device("cuda:0")
specs = make_specs(37, size="medium")
ens = Ensemble(specs, device=dev)
ens.simulate()
tables = P.prompt_tables(build_label_maps(ens.geometries[0], (512, 512), dev)[1], 1500)
tables.candidate_entries()
ca_host = [ens.frames_of(i)[:, 0, :].cpu().numpy() for i in range(37)]
acc = collections.defaultdict(float); lock = threading.Lock()
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        try: return fn(*a, **k)
        finally:
            dt = time.perf_counter() - t
            with lock: acc[name] += dt
    return w
P._frame_transforms = timed("transforms", P._frame_transforms)
P.sample_background = timed("sample_background", P.sample_background)
P.legacy_choice = timed("legacy_choice", P.legacy_choice)
P._assemble = timed("assemble", P._assemble)
P.instance_lut = timed("instance_lut", P.instance_lut)
def pouch(i):
    rs = np.random.RandomState(i); n = 0
    with torch.cuda.device(dev):
        for f in range(90):
            if not (ca_host[i][f] > 0.1).any(): continue
            t = time.perf_counter()
            text, _ = P.frame_prompts(tables, ca_host[i][f], rng=rs, as_text=True)
            with lock: acc["frame_prompts_total"] += time.perf_counter() - t
            n += 1
    return n
for th in (1, 8, 32):
    acc.clear()
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(th) as ex: n = sum(ex.map(pouch, range(37)))
    wall = time.perf_counter() - t0
    print(f"threads={th:2d} wall {wall:.2f} s frames {n}  " + "  ".join(f"{k}={v:.2f}" for k, v in sorted(acc.items())), flush=True)
