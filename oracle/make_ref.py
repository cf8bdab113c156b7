"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference is pure Python, so there is nothing to compile; ``/root/reference`` exists only in the build
container.  ``python -m oracle.make_ref`` (called by ``__graft_entry__.build()`` when the reference is mounted)
copies the files of the hot path, byte for byte, into ``oracle/_ref/`` -- git-ignored, never part of the
history, but shipped with the tree by ``gpurun`` like the built ``.so`` files -- so that ``bench.py`` can time the
real ``Pouch.simulate()`` / ``generate_image`` / mask functions on the GPU box's own host cores
(``cpu_baseline.kind = "reference"``) and ``tests/test_reference_live.py`` can run there too.  The three shims
of ``oracle/reference_shims.py`` (D_p monkeypatch, skimage stub, synthetic geometry dir) are applied at import
time, never by editing the copied files.  Only ``tests/``, ``bench.py``'s CPU legs and ``smoke()`` may use it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

SRC = os.environ.get("CALCIUM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["core/__init__.py", "core/pouch.py", "core/parameters.py", "core/geometry_loader.py",
         "utils/__init__.py", "utils/sam2_data_generator.py", "utils/image_processing.py",
         "artifacts/__init__.py", "artifacts/background.py", "artifacts/optical.py", "artifacts/sensor.py",
         "LICENSE"]


def make_ref(verbose: bool = False) -> str:
    if not os.path.isfile(os.path.join(SRC, "core", "pouch.py")):
        raise RuntimeError(f"reference not found under {SRC}")
    manifest = []
    for rel in FILES:
        s = os.path.join(SRC, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied unmodified from {SRC}")
    return DST


if __name__ == "__main__":
    make_ref(verbose=True)
    sys.exit(0)
