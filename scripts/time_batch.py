"""Wall-clock of the public batch entry point (generate_simulation_batch) for N medium pouches, clean frames, with and
without prompts, with a cProfile breakdown of the host side.  python scripts/time_batch.py [N] [profile]"""
import cProfile, os, pstats, random, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import generate_simulation_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
root = tempfile.mkdtemp(prefix="cab200_tb_", dir="/dev/shm")
out = sys.stdout; sys.stdout = open(os.devnull, "w")
try:
    random.seed(1); generate_simulation_batch(8, os.path.join(root, "warm"), pouch_sizes=["medium"], defect_configs=[{}] * 8)
    for prompts in (False, True):
        for rep in range(2):
            d = os.path.join(root, f"run{int(prompts)}{rep}")
            random.seed(1)
            pr = cProfile.Profile() if (len(sys.argv) > 2 and rep == 1) else None
            t0 = time.perf_counter()
            if pr: pr.enable()
            res = generate_simulation_batch(N, d, pouch_sizes=["medium"], defect_configs=[{}] * N, generate_prompts=prompts,
                                            prompt_threads=int(os.environ["PROMPT_THREADS"]) if os.environ.get("PROMPT_THREADS") else None)
            if pr: pr.disable()
            dt = time.perf_counter() - t0
            print(f"N={N} prompts={prompts} rep={rep}: {dt:.3f} s  images {res['stats']['total_images']} masks {res['stats']['total_masks']}", file=out, flush=True)
            if pr:
                pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(18)
            shutil.rmtree(d, ignore_errors=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
