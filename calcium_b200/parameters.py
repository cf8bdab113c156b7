"""Simulation parameter sets (host side).

API-compatible with the reference's ``SimulationParameters``
(reference ``core/parameters.py:19-308``): same 18 keys, same per-sim-type ``lower``/``upper``
ranges (``:69-90``) and -- important for parity -- the same *sequence of legacy
``np.random`` draws* in ``generate_random_params`` (``:222-282``), so that
``generate_random_params(seed=i)`` yields bit-identical dictionaries.

The randomisation is expressed as a small table (which keys are jittered, by what fraction of
``variation``, with optional clipping) walked in order; the draws themselves are
``np.random.uniform(1 - f, 1 + f)`` like the reference's.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

#: key order also fixes the order of ``params`` dictionaries (callers json-dump them)
BASE_PARAMETERS: Dict[str, float] = {
    "K_PLC": 0.2, "K_5": 0.66,
    "k_1": 1.11, "k_a": 0.08, "k_p": 0.13, "k_2": 0.0203,
    "V_SERCA": 0.9, "K_SERCA": 0.1,
    "c_tot": 2.0, "beta": 0.185,
    "k_i": 0.4, "tau_max": 800.0, "k_tau": 1.5,
    "D_p": 0.005, "D_c_ratio": 0.1,
    "frac": 0.007680491551459293,   # = 25 / 3255
}

SIM_TYPES: Dict[str, Dict[str, object]] = {
    "Single cell spikes": {"lower": 0.1, "upper": 0.5,
                           "description": "Isolated calcium spikes in individual cells"},
    "Intercellular transients": {"lower": 0.25, "upper": 0.6,
                                 "description": "Local calcium waves between adjacent cells"},
    "Intercellular waves": {"lower": 0.4, "upper": 0.8,
                            "description": "Propagating calcium waves across tissue"},
    "Fluttering": {"lower": 1.4, "upper": 1.5,
                   "description": "Synchronized oscillatory calcium dynamics"},
}
_FALLBACK_TYPE = "Intercellular waves"

# (key, divisor of `variation`, clip range or None), in draw order
_JITTER_PLAN = {
    "Single cell spikes": [("K_PLC", 1.0, None), ("k_1", 2.0, None)],
    "Intercellular transients": [("D_p", 1.0, None), ("D_c_ratio", 2.0, None)],
    "Intercellular waves": [("frac", 1.0, (0.001, 0.03)), ("D_p", 2.0, None)],
    "Fluttering": [("k_tau", 2.0, None), ("tau_max", 4.0, None)],
}
_POSITIVE_KEYS = ["K_PLC", "K_5", "k_1", "k_a", "k_p", "k_2", "V_SERCA", "K_SERCA", "c_tot",
                  "beta", "k_i", "tau_max", "k_tau", "D_p", "D_c_ratio"]


class SimulationParameters:
    """Parameter container; see module docstring for the reference mapping."""

    DEFAULT_PARAMS = BASE_PARAMETERS
    SIM_TYPE_PARAMS = SIM_TYPES

    def __init__(self, sim_type: Optional[str] = None,
                 custom_params: Optional[Dict[str, float]] = None):
        if sim_type is None:
            sim_type = _FALLBACK_TYPE
        if sim_type not in SIM_TYPES:
            print(f"Warning: Unknown simulation type '{sim_type}'. Using '{_FALLBACK_TYPE}'.")
            sim_type = _FALLBACK_TYPE
        self.sim_type = sim_type
        self.params = dict(BASE_PARAMETERS)
        self.params["lower"] = SIM_TYPES[sim_type]["lower"]
        self.params["upper"] = SIM_TYPES[sim_type]["upper"]
        if custom_params:
            self.params.update(custom_params)

    # -- validation / io ------------------------------------------------------------------
    def validate(self) -> bool:
        p = self.params
        for key in _POSITIVE_KEYS:
            if key in p and p[key] <= 0:
                raise ValueError(f"Parameter '{key}' must be positive, got {p[key]}")
        if "frac" in p and not 0 <= p["frac"] <= 1:
            raise ValueError(f"Parameter 'frac' must be in [0, 1], got {p['frac']}")
        if "lower" in p and "upper" in p and p["lower"] >= p["upper"]:
            raise ValueError(f"'lower' ({p['lower']}) must be less than 'upper' ({p['upper']})")
        return True

    def save_to_file(self, filepath: str) -> None:
        path = Path(filepath)
        path.parent.mkdir(parents=True, exist_ok=True)
        payload = {"simulation_type": self.sim_type, "parameters": self.params,
                   "description": SIM_TYPES.get(self.sim_type, {}).get("description", "")}
        with open(path, "w") as fh:
            json.dump(payload, fh, indent=4)

    @classmethod
    def load_from_file(cls, filepath: str) -> "SimulationParameters":
        with open(filepath, "r") as fh:
            payload = json.load(fh)
        inst = cls(sim_type=payload.get("simulation_type", _FALLBACK_TYPE))
        inst.params.update(payload.get("parameters", {}))
        inst.validate()
        return inst

    def get_params_dict(self) -> Dict[str, float]:
        return dict(self.params)

    # -- randomisation --------------------------------------------------------------------
    def generate_random_params(self, seed: Optional[int] = None,
                               variation: float = 0.2) -> Dict[str, float]:
        """Randomised copy of the parameters.  Draw order (legacy global ``np.random``):
        the two type-specific factors, then ``lower``, then ``upper``."""
        if seed is not None:
            np.random.seed(seed)
        out = dict(self.params)
        for key, div, clip in _JITTER_PLAN.get(self.sim_type, []):
            f = variation / div
            if clip is None:
                out[key] *= np.random.uniform(1 - f, 1 + f)
            else:
                out[key] = np.clip(out[key] * np.random.uniform(1 - f, 1 + f), clip[0], clip[1])
        for key in ("lower", "upper"):
            if key in out:
                out[key] *= np.random.uniform(1 - variation / 4, 1 + variation / 4)
        if "lower" in out and "upper" in out and out["lower"] >= out["upper"]:
            out["lower"], out["upper"] = out["upper"] * 0.9, out["lower"] * 1.1
        return out

    @classmethod
    def get_available_sim_types(cls) -> List[str]:
        return list(SIM_TYPES)

    def __repr__(self) -> str:
        return f"SimulationParameters(sim_type='{self.sim_type}', n_params={len(self.params)})"

    def __str__(self) -> str:
        desc = SIM_TYPES.get(self.sim_type, {}).get("description", "")
        rows = [f"Simulation Type: {self.sim_type}", f"Description: {desc}",
                f"Parameters ({len(self.params)}):"]
        rows += [f"  {k:15s} = {v}" for k, v in sorted(self.params.items())]
        return "\n".join(rows)
