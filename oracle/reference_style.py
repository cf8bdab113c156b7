"""ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's *execution style*, restated.

``oracle/csrc/oracle.c`` is a tight C port and therefore a much faster CPU baseline than the
reference really is.  This module restates the reference's loop the way it actually runs
(reference ``core/pouch.py:338-366``): a dense ``(n, 4, T)`` float64 history array indexed
``[:, k, step]`` (time is the fastest axis, so every per-step view is strided by 4*T*8 bytes), two
``scipy.sparse.csr_matrix.dot`` calls per step, and the step function compiled by numba with
``nopython, fastmath`` (``:31``) when numba is importable (plain numpy otherwise).  It exists to
report, next to the C port, what one host core achieves in the reference's own style; its results
are compared with the C oracle in tests (rtol, since fastmath may contract/reassociate).
"""
from __future__ import annotations

import numpy as np

try:
    from numba import jit
    HAVE_NUMBA = True
except Exception:  # noqa: BLE001
    HAVE_NUMBA = False

    def jit(*a, **k):
        return lambda f: f


@jit(nopython=True, cache=False, fastmath=True)
def _step(ca, ipt, s, r, ca_lap, ipt_lap, V_PLC, dt, k_1, k_a, k_p, k_2, V_SERCA, K_SERCA, K_PLC, K_5,
          c_tot, beta, k_tau, tau_max, k_i):
    term = (r * ca * ipt) / ((k_a + ca) * (k_p + ipt))
    J_channel = (k_1 * term ** 3 + k_2) * (s - ca)
    ca_sq = ca * ca
    J_SERCA = V_SERCA * ca_sq / (ca_sq + K_SERCA * K_SERCA)
    ca_next = ca + dt * (ca_lap + J_channel - J_SERCA)
    J_PLC = V_PLC.reshape(-1, 1) * ca_sq / (ca_sq + K_PLC * K_PLC)
    ipt_next = ipt + dt * (ipt_lap + J_PLC - K_5 * ipt)
    s_next = (c_tot - ca_next) / beta
    ca_4 = ca_sq * ca_sq
    k_tau_4 = k_tau * k_tau * k_tau * k_tau
    recovery = (k_tau_4 + ca_4) / (tau_max * k_tau_4)
    inact = 1.0 - r * (k_i + ca) / k_i
    r_next = r + dt * recovery * inact
    return ca_next, ipt_next, s_next, r_next


def simulate_reference_style(csr, p, r0, vplc, T=18000):
    """Returns the dense (n, 4, T) history, like ``Pouch.disc_dynamics`` after ``simulate()``."""
    from scipy.sparse import csr_matrix
    rowptr, col, val = csr
    n = len(rowptr) - 1
    L = csr_matrix((val, col, rowptr), shape=(n, n))
    D_p = p["D_p"]
    D_c = p["D_c_ratio"] * D_p
    dd = np.zeros((n, 4, T))
    dd[:, 2, 0] = (p["c_tot"] - dd[:, 0, 0]) / p["beta"]
    dd[:, 3, 0] = r0
    V = np.ascontiguousarray(vplc, dtype=np.float64)
    for step in range(1, T):
        ca = dd[:, 0, step - 1].reshape(-1, 1)
        ipt = dd[:, 1, step - 1].reshape(-1, 1)
        s = dd[:, 2, step - 1].reshape(-1, 1)
        r = dd[:, 3, step - 1].reshape(-1, 1)
        ca_lap = D_c * L.dot(ca)
        ipt_lap = D_p * L.dot(ipt)
        a, b, c, d = _step(ca, ipt, s, r, ca_lap, ipt_lap, V, 0.2, p["k_1"], p["k_a"], p["k_p"], p["k_2"],
                           p["V_SERCA"], p["K_SERCA"], p["K_PLC"], p["K_5"], p["c_tot"], p["beta"],
                           p["k_tau"], p["tau_max"], p["k_i"])
        dd[:, 0, step] = a.T
        dd[:, 1, step] = b.T
        dd[:, 2, step] = c.T
        dd[:, 3, step] = d.T
    return dd
