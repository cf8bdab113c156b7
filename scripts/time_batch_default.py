"""Wall clock of the public batch call with the REFERENCE'S DEFAULTS: no defect_configs, so every frame draws
generate_random_defect_config() (main.py:273-277) and takes the host chain (numpy's generator, OpenCV); prompts on.
Also with broken_defects="off" (the two defects that raise in the reference switched off: every frame is written).
python scripts/time_batch_default.py [N]"""
import os, random, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calcium_b200 import generate_simulation_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 37
root = tempfile.mkdtemp(prefix="cab200_tbd_", dir="/dev/shm")
out = sys.stdout; sys.stdout = open(os.devnull, "w")
try:
    random.seed(1); np.random.seed(1)
    generate_simulation_batch(2, os.path.join(root, "warm"), pouch_sizes=["medium"])
    for name, kw in (("defaults", {}), ("defaults, broken_defects=off", {"broken_defects": "off"})):
        for rep in range(2):
            d = os.path.join(root, f"run{rep}")
            random.seed(1); np.random.seed(1)
            t0 = time.perf_counter()
            res = generate_simulation_batch(N, d, pouch_sizes=["medium"], **kw)
            dt = time.perf_counter() - t0
            print(f"N={N} {name:30s} rep={rep}: {dt:.2f} s  images {res['stats']['total_images']} masks {res['stats']['total_masks']}", file=out, flush=True)
            shutil.rmtree(d, ignore_errors=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
