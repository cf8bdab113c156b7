"""calcium_b200 -- B200-native implementation of the simulation hot path of ygong2501/calcium.

Public surface (same names as the reference's Python API for this path):
  ``Pouch``, ``SimulationParameters``, ``GeometryLoader``, ``generate_simulation_batch``,
  ``BatchGenerationController``, ``create_instance_mask_from_cells``, ``apply_all_defects`` (deterministic
  defects only); plus the batched
  ``Ensemble`` / ``PouchSpec`` / ``make_specs`` driver the batch path is built on.

Importing the package does not need a GPU; every compute call does (there is no CPU fallback) and
needs ``libcalcium_b200.so`` (``python -m calcium_b200.build``).
"""
from .geometry import GeometryLoader, get_geometry  # noqa: F401
from .parameters import SimulationParameters  # noqa: F401

__all__ = ["Pouch", "SimulationParameters", "GeometryLoader", "generate_simulation_batch",
           "BatchGenerationController", "Ensemble", "PouchSpec", "make_specs",
           "create_instance_mask_from_cells", "get_geometry", "apply_all_defects"]


def __getattr__(name):
    # torch / the CUDA library are only imported when a compute entry point is touched
    if name in ("Pouch",):
        from .pouch import Pouch
        return Pouch
    if name in ("Ensemble", "PouchSpec", "make_specs"):
        from . import ensemble
        return getattr(ensemble, name)
    if name in ("generate_simulation_batch", "BatchGenerationController"):
        from . import batch
        return getattr(batch, name)
    if name == "create_instance_mask_from_cells":
        from .masks import create_instance_mask_from_cells
        return create_instance_mask_from_cells
    if name == "apply_all_defects":
        from .defects import apply_all_defects
        return apply_all_defects
    raise AttributeError(f"module 'calcium_b200' has no attribute {name!r}")
