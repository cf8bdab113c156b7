"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Build-container helper: import the *real* reference.

``/root/reference`` exists only in the build container (never on the GPU box), so this module is
used by ``oracle/make_golden.py`` and by the container-only tests that pin the oracle to the
reference itself.  Nothing here edits the reference; the three shims are applied by monkeypatch
(SURVEY.md section 8(c)):

  1. ``Pouch._unpack_parameters`` never sets ``self.D_p`` although ``core/pouch.py:194`` reads it
     (defect D1) -> wrapped to add it.
  2. ``skimage`` is not installed -> a stub package exposing ``draw.polygon`` (the restatement in
     ``oracle/pouch_oracle.py``) and empty ``measure`` / ``transform`` modules
     (imported at ``core/pouch.py:24``, ``utils/labeling.py:15``, ``utils/image_processing.py:6``).
  3. geometry files are absent -> ``geometry_dir=`` points at synthetic files written in the
     reference's own format by ``calcium_b200.geometry.write_reference_format``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

def _find_reference_root() -> str:
    """The mounted reference in the build container; on the GPU box the byte-for-byte copy that
    ``oracle/make_ref.py`` leaves under ``oracle/_ref`` (git-ignored, shipped with the tree)."""
    env = os.environ.get("CALCIUM_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/core/pouch.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "pouch.py"))


def _install_skimage_stub() -> None:
    if "skimage" in sys.modules:
        return
    from oracle.pouch_oracle import polygon_skimage
    pkg = types.ModuleType("skimage")
    pkg.__path__ = []  # mark as package
    draw = types.ModuleType("skimage.draw")
    draw.polygon = polygon_skimage
    measure = types.ModuleType("skimage.measure")
    transform = types.ModuleType("skimage.transform")
    pkg.draw, pkg.measure, pkg.transform = draw, measure, transform
    sys.modules.update({"skimage": pkg, "skimage.draw": draw, "skimage.measure": measure,
                        "skimage.transform": transform})


_loaded = None


def load_reference():
    """Returns a namespace with the reference's ``Pouch``, ``SimulationParameters``,
    ``_simulate_time_step`` (numba dispatcher), ``create_instance_mask_from_cells`` and the
    ``core.pouch`` module itself."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_ref"))
    try:
        import skimage  # noqa: F401
    except ImportError:
        _install_skimage_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import core.pouch as ref_pouch
    from core.parameters import SimulationParameters
    from utils.sam2_data_generator import create_instance_mask_from_cells

    Pouch = ref_pouch.Pouch
    if not getattr(Pouch, "_d1_patched", False):
        orig = Pouch._unpack_parameters

        def _unpack_with_dp(self):
            orig(self)
            self.D_p = self.param_dict["D_p"]

        Pouch._unpack_parameters = _unpack_with_dp
        Pouch._d1_patched = True

    ns = types.SimpleNamespace(
        Pouch=Pouch, SimulationParameters=SimulationParameters, module=ref_pouch,
        step_jit=ref_pouch._simulate_time_step,
        create_instance_mask_from_cells=create_instance_mask_from_cells)
    _loaded = ns
    return ns


def synthetic_geometry_dir(sizes=("xsmall", "small", "medium")) -> str:
    """Directory with the synthetic geometry in the reference's three-file format."""
    from calcium_b200.geometry import write_reference_format, VERTICES_FILE
    d = os.path.join(tempfile.gettempdir(), "calcium_ref_geometry_" + "_".join(sizes))
    if not os.path.exists(os.path.join(d, VERTICES_FILE)):
        write_reference_format(d, sizes)
    return d


class use_py_func:
    """Context manager: run the reference with the un-jitted step function
    (``_simulate_time_step.py_func``) -- strict numpy evaluation, ISA independent."""

    def __enter__(self):
        self.ns = load_reference()
        self.saved = self.ns.module._simulate_time_step
        self.ns.module._simulate_time_step = self.saved.py_func
        return self

    def __exit__(self, *exc):
        self.ns.module._simulate_time_step = self.saved
        return False
