"""Instance-mask helpers with the reference's names (``utils/sam2_data_generator.py``).

``create_instance_mask_from_cells(cell_masks, active_cells)`` exists for callers that already hold
the two host arrays; it is a pure relabelling through a look-up table (no per-cell full-image
passes like the reference's loop at ``:34-35``) and reproduces that loop's result bit for bit,
including its 0-based-vs-1-based id mismatch.  The batch path never calls it: K2
produces the same mask on the GPU straight from the frames (``Pouch.get_instance_mask``).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def create_instance_mask_from_cells(cell_masks: np.ndarray,
                                    active_cells: Optional[List[int]] = None) -> np.ndarray:
    if active_cells is None:
        return cell_masks.astype(np.uint16)
    top = int(cell_masks.max()) if cell_masks.size else 0
    lut = np.zeros(max(top, max(active_cells, default=0)) + 2, dtype=np.uint16)
    # later entries of the loop overwrite earlier ones only for equal ids; ids are unique
    for new_id, cell_id in enumerate(active_cells, start=1):
        if 0 <= cell_id < len(lut):
            lut[cell_id] = new_id
    return lut[np.clip(cell_masks, 0, len(lut) - 1)] * (cell_masks >= 0)
