"""Builds ``calcium_b200/libcalcium_b200.so`` (sm_100a only) with nvcc, in-tree.

``python -m calcium_b200.build`` or ``build()``; the ``.so`` is git-ignored but travels with the
tree to the GPU box.  No JIT, no fallback: importing the product without this library fails.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcalcium_b200.so")
OBJDIR = os.path.join(HERE, "_obj")
SOURCES = ["integrate_f64.cu", "integrate_cl64.cu", "integrate_f32.cu", "integrate_cl32.cu", "integrate.cu",
           "api.cu", "microbench.cu", "relay.cu", "filewriter.cu", "raster.cu", "edges.cu", "legacy_rng.cu", "prompts_host.cu", "defects.cu", "encode.cu", "encode_raw.cu", "edt.cu", "plan.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libcalcium_b200.so cannot be built")
    return exe


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "cab200.h"))
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build_variant(name: str, defines: List[str]) -> str:
    """An experimental build of the whole library with extra -D flags -> calcium_b200/_variants/lib<name>.so
    (selected at run time with CALCIUM_B200_LIB; see scripts/k1_variants.py)."""
    vdir = os.path.join(HERE, "_variants")
    odir = os.path.join(vdir, "_obj_" + name)
    os.makedirs(odir, exist_ok=True)
    out = os.path.join(vdir, f"lib{name}.so")

    def comp(src: str) -> str:
        obj = os.path.join(odir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(comp, SOURCES))
    res = subprocess.run([_nvcc()] + ARCH + ["-shared", "-o", out] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


K1_CHECK_LIB = os.path.join(HERE, "_variants", "libk1check.so")
K1_CHECK_OBJDIR = os.path.join(HERE, "_variants", "_obj_k1check")
K1_CHECK_SOURCES = ["integrate_f64.cu", "integrate_cl64.cu"]


def _compile_check(src: str) -> str:
    obj = os.path.join(K1_CHECK_OBJDIR, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DCAB_K1_CHECK", "-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} (-DCAB_K1_CHECK):\n{res.stdout}\n{res.stderr}")
    return obj


def _link_check(check_objs: List[str]) -> str:
    objs = check_objs + [os.path.join(OBJDIR, os.path.splitext(s_)[0] + ".o") for s_ in SOURCES if s_ not in K1_CHECK_SOURCES]
    res = subprocess.run([_nvcc()] + ARCH + ["-shared", "-o", K1_CHECK_LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return K1_CHECK_LIB


def build_k1_check(force: bool = False) -> str:
    """``_variants/libk1check.so``: the library with the FP64 integrator compiled with ``-DCAB_K1_CHECK`` (the step
    loop's buffer-reuse invariants as run-time checks, see integrate.cuh); every other object is the regular build's.
    Loaded through ``CALCIUM_B200_LIB`` by ``tests/test_gpu_parity.py::test_k1_protocol_checks``.  ``build()`` makes it
    in the same compiler pool as the regular library; this only rebuilds it when it is missing or stale."""
    build()
    if not force and os.path.exists(K1_CHECK_LIB) and os.path.getmtime(K1_CHECK_LIB) >= os.path.getmtime(LIB):
        return K1_CHECK_LIB
    os.makedirs(K1_CHECK_OBJDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=2) as ex:
        objs = list(ex.map(_compile_check, K1_CHECK_SOURCES))
    return _link_check(objs)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(K1_CHECK_OBJDIR, exist_ok=True)
    # one pool for the regular objects and the two -DCAB_K1_CHECK ones (the integrator's translation units are the
    # long poles; the check variant would otherwise add their compile time once more)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES) + len(K1_CHECK_SOURCES)) as ex:
        check_futs = [ex.submit(_compile_check, s_) for s_ in K1_CHECK_SOURCES]
        objs: List[str] = list(ex.map(lambda s_: _compile(s_, verbose), SOURCES))
        check_objs = [f.result() for f in check_futs]
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    _link_check(check_objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
