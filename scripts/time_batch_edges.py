"""Wall clock of the public batch call on the routes whose frames are not flat-shaded: edge_blur=True, and a
deterministic defect configuration (K5) -- images through K3r (device PNG encoder for arbitrary frames), masks
through K3's mask stream.  python scripts/time_batch_edges.py [N]"""
import os, random, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import generate_simulation_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 37
root = tempfile.mkdtemp(prefix="cab200_tbe_", dir="/dev/shm")
out = sys.stdout; sys.stdout = open(os.devnull, "w")
det = {"background_fluorescence": True, "non_uniform_background": False, "vignetting": True, "dynamic_range_compression": True,
       "quantization": True, "brightness_contrast": True}
try:
    random.seed(1); generate_simulation_batch(4, os.path.join(root, "warm"), pouch_sizes=["medium"], defect_configs=[{}] * 4, edge_blur=True)
    for name, kw in (("clean", dict(defect_configs=[{}] * N)), ("edge_blur", dict(defect_configs=[{}] * N, edge_blur=True)),
                     ("deterministic defects", dict(defect_configs=[det] * N))):
        for rep in range(2):
            d = os.path.join(root, f"run{rep}")
            random.seed(1)
            t0 = time.perf_counter()
            res = generate_simulation_batch(N, d, pouch_sizes=["medium"], generate_prompts=False, **kw)
            dt = time.perf_counter() - t0
            nbytes = sum(os.path.getsize(os.path.join(r_, f)) for r_, _, fs in os.walk(d) for f in fs)
            print(f"N={N} {name:22s} rep={rep}: {dt:.3f} s  images {res['stats']['total_images']} masks {res['stats']['total_masks']}  {nbytes/1e6:.0f} MB", file=out, flush=True)
            shutil.rmtree(d, ignore_errors=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
