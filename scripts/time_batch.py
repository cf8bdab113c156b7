"""Wall-clock of the public batch entry point (generate_simulation_batch) for N medium pouches, with and
without prompts / file writing: where the host time goes once the GPU part takes milliseconds."""
import os, sys, time, tempfile, shutil, io, contextlib, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from calcium_b200 import generate_simulation_batch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 37
def run(label, **kw):
    d = tempfile.mkdtemp(prefix="cab_batch_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        random.seed(1)                         # the batch draws sizes / sim types from the stdlib generator
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = generate_simulation_batch(N, d, pouch_sizes=["medium"], **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = res.get("stats", {})
        print(f"{label:44s} {dt:7.2f} s  ({dt / N * 1e3:7.1f} ms per pouch)  images {st.get('total_images')} masks {st.get('total_masks')} "
              f"prompts {sum(s.get('prompt_count', 0) for s in res['simulations'])}", flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
run("warm-up (2 pouches)") if False else None
generate_simulation_batch.__module__
run("compute only (no files, no prompts)", write_files=False, generate_prompts=False)
run("compute only (no files, no prompts) again", write_files=False, generate_prompts=False)
run("files, no prompts", write_files=True, generate_prompts=False)
run("files + prompts (default)", write_files=True, generate_prompts=True)
run("edge_blur motion5: files, no prompts", write_files=True, generate_prompts=False, edge_blur=True, blur_kernel_size=5, blur_type="motion")
for th in (4, 16, 32):
    run(f"files + prompts, prompt_threads={th}", write_files=True, generate_prompts=True, prompt_threads=th)
