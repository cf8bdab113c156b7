"""cProfile of the batch path's prompt generation (host side of row f2) for a few medium pouches."""
import os, sys, io, cProfile, pstats, tempfile, shutil, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import generate_simulation_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = tempfile.mkdtemp(prefix="cab_prof_", dir="/dev/shm")
with contextlib.redirect_stdout(io.StringIO()):
    generate_simulation_batch(2, d, pouch_sizes=["medium"], write_files=True)      # warm
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    generate_simulation_batch(N, d, pouch_sizes=["medium"], write_files=True, prompt_threads=int(sys.argv[2]) if len(sys.argv) > 2 else 8)
pr.disable()
shutil.rmtree(d, ignore_errors=True)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
