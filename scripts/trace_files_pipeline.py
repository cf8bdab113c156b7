"""Timeline of Ensemble.run_files_to_host on the benchmark workload: where do the gaps come from? (GPU)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import ensemble as E
dev = torch.device("cuda:0")
P, chunk, steps = 148, int(os.environ.get("CHUNK", "37")), 4
ens = E.Ensemble(E.make_specs(P, size="medium"), device=dev)
ens.run_files_to_host(chunk_pouches=chunk); torch.cuda.synchronize()
marks = []
orig_sim, orig_enc = ens.simulate, ens.encode
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def sim(*a, **k):
    b = ev(); r = orig_sim(*a, **k); marks.append(("K1", b, ev(), time.perf_counter())); return r
def enc(*a, **k):
    b = ev(); r = orig_enc(*a, **k); marks.append(("K3", b, ev(), time.perf_counter())); return r
ens.simulate, ens.encode = sim, enc
torch.cuda.synchronize()
t0 = ev(); h0 = time.perf_counter()
for _ in range(steps):
    ens.run_files_to_host(chunk_pouches=chunk, wait=False)
ens.flush_files()
t1 = ev(); torch.cuda.synchronize()
print(f"total {t0.elapsed_time(t1)/steps:.2f} ms/step, host loop {1e3*(time.perf_counter()-h0)/steps:.2f} ms/step")
for name, b, e, h in marks:
    print(f"{name}  start {t0.elapsed_time(b):8.2f}  end {t0.elapsed_time(e):8.2f}  dur {b.elapsed_time(e):6.2f}   host-queued at {1e3*(h-h0):8.2f}")
