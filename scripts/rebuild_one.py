"""Development helper: recompile only the named csrc files and relink both libraries (the regular build and the
-DCAB_K1_CHECK variant) -- `python scripts/rebuild_one.py encode.cu`.  `python -m calcium_b200.build` stays the
authoritative full build (__graft_entry__.build())."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import build as b
for src in sys.argv[1:]:
    b._compile(src, verbose=False)
    if src in b.K1_CHECK_SOURCES:
        b._compile_check(src)
objs = [os.path.join(b.OBJDIR, os.path.splitext(s)[0] + ".o") for s in b.SOURCES]
subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", b.LIB] + objs, check=True)
b._link_check([os.path.join(b.K1_CHECK_OBJDIR, os.path.splitext(s)[0] + ".o") for s in b.K1_CHECK_SOURCES])
print(b.LIB)
