"""Builds experimental variants of the library (extra -D flags) HERE, then on the GPU box: python scripts/k1_variants.py run"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANTS = {n: d.split() for n, d in (l.split(":") for l in os.environ.get("K1_VARIANTS", "base:").split(";") if l)}
if len(sys.argv) > 1 and sys.argv[1] == "run":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in VARIANTS:
        env = dict(os.environ, CALCIUM_B200_LIB=os.path.join(here, "calcium_b200", "_variants", f"lib{name}.so"))
        subprocess.run([sys.executable, os.path.join(here, "scripts", "time_k1.py")], env=env)
else:
    from calcium_b200 import build
    for name, defs in VARIANTS.items():
        print(name, defs, build.build_variant(name, defs))
