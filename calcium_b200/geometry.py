"""Wing-disc cell geometry: loading, synthetic generation and device-ready export.

Mirrors the interface of the reference's ``GeometryLoader``
(reference ``core/geometry_loader.py:16-194``): ``GeometryLoader(geometry_dir)
.load_geometry(size) -> (vertices, adjacency, laplacian)`` with the same three pickled-dict
``.npy`` files (``disc_vertices.npy``, ``disc_sizes_adj.npy``, ``disc_sizes_laplacian.npy``,
reference ``core/geometry_loader.py:31-33``) when a directory holding them is given.

The MSELab geometry files are not shipped with the reference (its ``.gitignore:21-23``), so this
module also contains a deterministic synthetic generator (Lloyd-relaxed Voronoi cells inside an
ellipse) used whenever the files are absent.  Cell counts: xsmall 200, small 500, medium 1500,
large 3255 (``frac = 25/3255`` in reference ``core/parameters.py:64`` pins the last one; the others
are assumed).

Differences from the reference, on purpose:
  * geometry is cached per (directory, size) instead of re-reading every size from disk for every
    ``Pouch`` (reference ``core/geometry_loader.py:123-125``);
  * ``Geometry.csr`` exposes the Laplacian as CSR with ascending column order and the diagonal in
    place -- exactly what ``scipy.sparse.csr_matrix(dense)`` yields at reference
    ``core/pouch.py:183`` -- which is the form the CUDA integrator consumes.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

VERTICES_FILE = "disc_vertices.npy"
ADJACENCY_FILE = "disc_sizes_adj.npy"
LAPLACIAN_FILE = "disc_sizes_laplacian.npy"

#: synthetic cell counts per size name (see module docstring)
SYNTHETIC_CELL_COUNTS: Dict[str, int] = {
    "xsmall": 200,
    "small": 500,
    "medium": 1500,
    "large": 3255,
    "xlarge": 6500,      # the reference lists the name (core/geometry_loader.py:36) without a size; K1 runs it as a cluster of 8
}
_SYNTHETIC_SEEDS = {"xsmall": 1000, "small": 1001, "medium": 1002, "large": 1003, "xlarge": 1004}
DEFAULT_SIZES = ["xsmall", "small", "medium", "large", "xlarge"]


@dataclass
class Geometry:
    """One pouch geometry, host side."""

    size: str
    vertices: np.ndarray            # object array, n x (k_i, 2) float64
    edges: np.ndarray               # (m, 2) int32, i < j, unique
    weights: Optional[np.ndarray] = None   # dense laplacian if it came from files
    _adj: Optional[np.ndarray] = field(default=None, repr=False)
    _lap: Optional[np.ndarray] = field(default=None, repr=False)
    _csr: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(default=None, repr=False)

    @property
    def n_cells(self) -> int:
        return len(self.vertices)

    @property
    def adjacency(self) -> np.ndarray:
        if self._adj is None:
            n = self.n_cells
            a = np.zeros((n, n), dtype=np.float64)
            a[self.edges[:, 0], self.edges[:, 1]] = 1.0
            a[self.edges[:, 1], self.edges[:, 0]] = 1.0
            self._adj = a
        return self._adj

    @property
    def laplacian(self) -> np.ndarray:
        """Dense ``adj - diag(rowsum)``: negative diagonal, because the reference *adds*
        ``D * L @ c`` (reference ``core/pouch.py:346-347,87``)."""
        if self._lap is None:
            if self.weights is not None:
                self._lap = self.weights
            else:
                a = self.adjacency
                self._lap = a - np.diag(a.sum(axis=1))
        return self._lap

    @property
    def csr(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(rowptr int32 [n+1], colidx int32 [nnz], vals float64 [nnz]); columns ascending,
        explicit zeros dropped -- the structure of ``csr_matrix(dense)``."""
        if self._csr is None:
            if self.weights is not None:
                lap = self.weights
                rows, cols = np.nonzero(lap)           # row-major => ascending cols per row
                vals = lap[rows, cols].astype(np.float64)
                n = lap.shape[0]
                rowptr = np.zeros(n + 1, dtype=np.int64)
                np.add.at(rowptr, rows + 1, 1)
                rowptr = np.cumsum(rowptr)
            else:
                n = self.n_cells
                e = self.edges
                deg = np.bincount(e.ravel(), minlength=n)
                r = np.concatenate([e[:, 0], e[:, 1], np.arange(n)[deg > 0]])
                c = np.concatenate([e[:, 1], e[:, 0], np.arange(n)[deg > 0]])
                v = np.concatenate([np.ones(2 * len(e)), -deg[deg > 0].astype(np.float64)])
                order = np.lexsort((c, r))
                rows, cols, vals = r[order], c[order], v[order]
                rowptr = np.zeros(n + 1, dtype=np.int64)
                np.add.at(rowptr, rows + 1, 1)
                rowptr = np.cumsum(rowptr)
            self._csr = (rowptr.astype(np.int32), np.asarray(cols, dtype=np.int32),
                         np.ascontiguousarray(vals, dtype=np.float64))
        return self._csr

    def bbox(self) -> Tuple[float, float, float, float]:
        allv = np.concatenate([np.asarray(v, dtype=np.float64) for v in self.vertices])
        return (float(allv[:, 0].min()), float(allv[:, 0].max()),
                float(allv[:, 1].min()), float(allv[:, 1].max()))

    def scaled_polygons(self, output_size: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
        """Integer pixel polygons, flattened.

        Same arithmetic as reference ``core/pouch.py:272-281`` and ``:430-446``:
        ``((x - x_min) / (x_max - x_min) * (W - 1)).astype(int)`` (truncation toward zero).
        Returns (poly_off int32 [n+1], pts int32 [sum k, 2] as (x, y)).
        """
        width, height = output_size
        x_min, x_max, y_min, y_max = self.bbox()
        offs = [0]
        chunks = []
        for v in self.vertices:
            v = np.asarray(v, dtype=np.float64)
            sx = ((v[:, 0] - x_min) / (x_max - x_min) * (width - 1)).astype(int)
            sy = ((v[:, 1] - y_min) / (y_max - y_min) * (height - 1)).astype(int)
            chunks.append(np.stack([sx, sy], axis=1))
            offs.append(offs[-1] + len(v))
        return (np.asarray(offs, dtype=np.int32),
                np.ascontiguousarray(np.concatenate(chunks), dtype=np.int32))


# ----------------------------------------------------------------------------------------------
# synthetic generator
# ----------------------------------------------------------------------------------------------

def _polygon_centroid(p: np.ndarray) -> np.ndarray:
    x, y = p[:, 0], p[:, 1]
    xn, yn = np.roll(x, -1), np.roll(y, -1)
    cross = x * yn - xn * y
    area = cross.sum() / 2.0
    if abs(area) < 1e-14:
        return p.mean(axis=0)
    cx = ((x + xn) * cross).sum() / (6.0 * area)
    cy = ((y + yn) * cross).sum() / (6.0 * area)
    return np.array([cx, cy])


def make_synthetic_geometry(size: str, n_cells: Optional[int] = None,
                            seed: Optional[int] = None, lloyd_iters: int = 15) -> Geometry:
    """Deterministic synthetic wing-disc pouch: ``n_cells`` Voronoi cells in an ellipse.

    Seeds in a padded ellipse (a=1, b=0.75) are Lloyd-relaxed; the ``n_cells`` innermost finite
    Voronoi regions are kept, their vertices ordered counter-clockwise and scaled by 100.
    Adjacency = shared Voronoi ridges between kept cells.
    """
    from scipy.spatial import Voronoi

    if n_cells is None:
        n_cells = SYNTHETIC_CELL_COUNTS[size]
    if seed is None:
        seed = _SYNTHETIC_SEEDS.get(size, 1234)
    rng = np.random.RandomState(seed)
    a, b, pad = 1.0, 0.75, 1.25
    n_seed = int(np.ceil(n_cells * pad * pad * 1.15)) + 32
    # uniform in the padded ellipse by rejection
    pts = np.empty((0, 2))
    while len(pts) < n_seed:
        cand = rng.uniform(-1.0, 1.0, size=(4 * n_seed, 2)) * np.array([a * pad, b * pad])
        keep = (cand[:, 0] / (a * pad)) ** 2 + (cand[:, 1] / (b * pad)) ** 2 <= 1.0
        pts = np.concatenate([pts, cand[keep]])
    pts = pts[:n_seed]

    def ell_r(q):
        return np.sqrt((q[:, 0] / a) ** 2 + (q[:, 1] / b) ** 2)

    for _ in range(lloyd_iters):
        vor = Voronoi(pts)
        new = pts.copy()
        for i, ridx in enumerate(vor.point_region):
            reg = vor.regions[ridx]
            if len(reg) < 3 or -1 in reg:
                continue
            poly = vor.vertices[reg]
            if ell_r(poly).max() > pad * 1.3:
                continue
            ctr = poly.mean(axis=0)
            ang = np.arctan2(poly[:, 1] - ctr[1], poly[:, 0] - ctr[0])
            new[i] = _polygon_centroid(poly[np.argsort(ang)])
        pts = new

    vor = Voronoi(pts)
    finite = np.array([(-1 not in vor.regions[r]) and len(vor.regions[r]) >= 3
                       for r in vor.point_region])
    radius = ell_r(pts)
    radius[~finite] = np.inf
    order = np.argsort(radius, kind="stable")
    kept = order[:n_cells]
    if not np.all(np.isfinite(radius[kept])):
        raise RuntimeError("synthetic geometry: not enough finite Voronoi cells")
    # stable renumbering: by angle-sweep rows (y then x) so ids have some spatial coherence
    kept = kept[np.lexsort((pts[kept, 0], np.round(pts[kept, 1] * 12.0)))]
    new_id = -np.ones(len(pts), dtype=np.int64)
    new_id[kept] = np.arange(n_cells)

    verts = np.empty(n_cells, dtype=object)
    for k, i in enumerate(kept):
        poly = vor.vertices[vor.regions[vor.point_region[i]]]
        ctr = poly.mean(axis=0)
        ang = np.arctan2(poly[:, 1] - ctr[1], poly[:, 0] - ctr[0])
        verts[k] = np.ascontiguousarray(poly[np.argsort(ang)] * 100.0)

    edges = set()
    for p, q in vor.ridge_points:
        i, j = new_id[p], new_id[q]
        if i >= 0 and j >= 0 and i != j:
            edges.add((min(i, j), max(i, j)))
    edges = np.array(sorted(edges), dtype=np.int32).reshape(-1, 2)
    return Geometry(size=size, vertices=verts, edges=edges)


# ----------------------------------------------------------------------------------------------
# loader (reference-compatible surface)
# ----------------------------------------------------------------------------------------------

_CACHE: Dict[Tuple[str, str], Geometry] = {}
_CACHE_LOCK = threading.Lock()


def _synthetic_cache_path(size: str) -> str:
    root = os.environ.get("CALCIUM_B200_GEOMETRY_CACHE",
                          os.path.join(os.path.dirname(os.path.abspath(__file__)), "_geometry_cache"))
    return os.path.join(root, f"synthetic_{size}.npz")


def get_geometry(size: str, geometry_dir: Optional[str] = None) -> Geometry:
    """Cached geometry for ``size``; from ``geometry_dir``'s reference-format files when they
    exist, else the synthetic generator (result cached on disk as a compact npz)."""
    key = (geometry_dir or "", size)
    with _CACHE_LOCK:
        geo = _CACHE.get(key)
        if geo is not None:
            return geo
        if geometry_dir is not None and os.path.exists(os.path.join(geometry_dir, VERTICES_FILE)):
            geo = _load_reference_files(geometry_dir, size)
        else:
            if size not in SYNTHETIC_CELL_COUNTS:
                raise ValueError(
                    f"Geometry size '{size}' not available. Available sizes: "
                    f"{', '.join(SYNTHETIC_CELL_COUNTS)}.")
            path = _synthetic_cache_path(size)
            geo = None
            if os.path.exists(path):
                try:
                    z = np.load(path)
                    offs, flat, edges = z["offs"], z["flat"], z["edges"]
                    verts = np.empty(len(offs) - 1, dtype=object)
                    for i in range(len(offs) - 1):
                        verts[i] = flat[offs[i]:offs[i + 1]].copy()
                    geo = Geometry(size=size, vertices=verts, edges=edges)
                except Exception:
                    geo = None
            if geo is None:
                geo = make_synthetic_geometry(size)
                try:
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    offs = np.cumsum([0] + [len(v) for v in geo.vertices]).astype(np.int64)
                    flat = np.concatenate(list(geo.vertices))
                    tmp = path + f".{os.getpid()}.tmp.npz"
                    np.savez(tmp, offs=offs, flat=flat, edges=geo.edges)
                    os.replace(tmp, path)
                except OSError:
                    pass
        _CACHE[key] = geo
        return geo


def _load_reference_files(geometry_dir: str, size: str) -> Geometry:
    paths = [os.path.join(geometry_dir, f) for f in (VERTICES_FILE, ADJACENCY_FILE, LAPLACIAN_FILE)]
    for p in paths:
        if not os.path.exists(p):
            raise FileNotFoundError(
                f"Geometry file not found: {p}. "
                f"Please download geometry files from the original repository.")
    try:
        vdict = np.load(paths[0], allow_pickle=True).item()
        adict = np.load(paths[1], allow_pickle=True).item()
        ldict = np.load(paths[2], allow_pickle=True).item()
        vertices, adj, lap = vdict[size], np.asarray(adict[size]), np.asarray(ldict[size])
    except KeyError:
        raise RuntimeError(f"Geometry size '{size}' exists in file list but could not be loaded. "
                           f"Geometry files may be corrupted.")
    except Exception as e:  # noqa: BLE001 - same wrapping as the reference loader
        raise RuntimeError(f"Error loading geometry files: {e}")
    n = len(vertices)
    if adj.shape != (n, n):
        raise RuntimeError(f"Adjacency matrix shape {adj.shape} does not match number of cells ({n})")
    if lap.shape != (n, n):
        raise RuntimeError(f"Laplacian matrix shape {lap.shape} does not match number of cells ({n})")
    iu, ju = np.nonzero(np.triu(adj, 1))
    edges = np.stack([iu, ju], axis=1).astype(np.int32)
    verts = np.empty(n, dtype=object)
    for i in range(n):
        verts[i] = np.asarray(vertices[i], dtype=np.float64)
    geo = Geometry(size=size, vertices=verts, edges=edges,
                   weights=np.ascontiguousarray(lap, dtype=np.float64))
    geo._adj = np.asarray(adj, dtype=np.float64)
    return geo


class GeometryLoader:
    """Drop-in for the reference's ``GeometryLoader`` (``core/geometry_loader.py:16``)."""

    VERTICES_FILE = VERTICES_FILE
    ADJACENCY_FILE = ADJACENCY_FILE
    LAPLACIAN_FILE = LAPLACIAN_FILE
    DEFAULT_SIZES = DEFAULT_SIZES

    def __init__(self, geometry_dir: Optional[str] = None):
        self.geometry_dir = geometry_dir
        self.available_sizes = self._discover_available_sizes()

    def _has_files(self) -> bool:
        return self.geometry_dir is not None and os.path.exists(
            os.path.join(self.geometry_dir, VERTICES_FILE))

    def _discover_available_sizes(self) -> List[str]:
        if not self._has_files():
            return list(SYNTHETIC_CELL_COUNTS)
        try:
            d = np.load(os.path.join(self.geometry_dir, VERTICES_FILE), allow_pickle=True).item()
            return sorted(d.keys())
        except Exception as e:  # noqa: BLE001
            print(f"Warning: Could not load geometry sizes from {self.geometry_dir}: {e}")
            return list(DEFAULT_SIZES)

    def geometry(self, size: str = "medium") -> Geometry:
        if size not in self.available_sizes:
            avail = ", ".join(self.available_sizes) if self.available_sizes else "none"
            raise ValueError(f"Geometry size '{size}' not available. Available sizes: {avail}. "
                             f"Please download geometry files from the original repository.")
        return get_geometry(size, self.geometry_dir if self._has_files() else None)

    def load_geometry(self, size: str = "medium") -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        g = self.geometry(size)
        return g.vertices, g.adjacency, g.laplacian

    def get_available_sizes(self) -> List[str]:
        return list(self.available_sizes)

    def get_geometry_info(self, size: str) -> dict:
        if size not in self.available_sizes:
            raise ValueError(f"Size '{size}' not available")
        try:
            return {"size": size, "n_cells": self.geometry(size).n_cells,
                    "geometry_dir": self.geometry_dir}
        except Exception as e:  # noqa: BLE001
            return {"size": size, "error": str(e)}


def write_reference_format(geometry_dir: str, sizes=("xsmall", "small", "medium", "large")) -> None:
    """Write the three pickled-dict ``.npy`` files the reference's loader reads
    (reference ``core/geometry_loader.py:123-130``) from the synthetic geometries."""
    os.makedirs(geometry_dir, exist_ok=True)
    vd, ad, ld = {}, {}, {}
    for s in sizes:
        g = get_geometry(s)
        vd[s], ad[s], ld[s] = g.vertices, g.adjacency, g.laplacian
    np.save(os.path.join(geometry_dir, VERTICES_FILE), vd, allow_pickle=True)
    np.save(os.path.join(geometry_dir, ADJACENCY_FILE), ad, allow_pickle=True)
    np.save(os.path.join(geometry_dir, LAPLACIAN_FILE), ld, allow_pickle=True)
