"""CPU restatement of the ``edge_blur`` / ``with_border`` branches of ``Pouch.generate_image``
(reference ``core/pouch.py:436-481``, blur kernels ``:494-513``) -- SURVEY.md section 8 row f1.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, never by the product (``calcium_b200`` computes
these on the GPU: ``cab200_edge_maps`` / ``cab200_raster_edges_batch``).

The reference delegates the arithmetic to OpenCV (``opencv-python >= 4.5``, ``requirements.txt``; 4.13.0
in this image), which is not vendored under ``/root/reference``.  The functions below restate the
published algorithms of the four calls the branch makes, in plain numpy / Python loops, and are pinned
(``tests/test_oracle_golden.py``) against (i) the reference's own ``generate_image`` outputs recorded in
``tests/golden/raster_xsmall.npz`` (``edge_mean3``, ``edge_motion5``, ``border_blur``) and (ii) the real
``cv2`` functions on the same inputs:

* ``cv2.polylines(mask, [pts], True, 255, 1)`` -> ``polylines_oracle``: thickness 1, shift 0, ``LINE_8``
  draws every segment with the Bresenham ``cv::LineIterator`` (left to right, 8-connected) -- the same
  iterator ``fillPoly`` outlines its polygons with;
* ``cv2.filter2D(src_u8, -1, kernel_f32)`` -> ``filter2d_oracle``: correlation, anchor at
  ``ksize // 2``, ``BORDER_REFLECT_101``, float32 accumulation over the kernel's non-zero taps in
  row-major order with one fused multiply-add per tap (OpenCV's AVX2 path), result rounded half-to-even
  and saturated to uint8.  For odd kernel sizes (the GUI offers 3/5/7/9) and powers of two no sum can come
  within rounding error of a .5 tie, so the accumulation details cannot show; sizes 6 and 10 can tie, and
  there the real cv2 mixes FMA (SIMD body) and mul+add (row tail) -- not pinned;
* ``cv2.dilate(src, ones((3, 3)))`` -> ``dilate3_oracle``: 3x3 maximum, pixels outside the image ignored;
* the blend ``(img * (1 - m) + blurred * m).astype(uint8)`` with ``m = dilated / 255.0`` in float64.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

def line_pixels(p1: Tuple[int, int], p2: Tuple[int, int]) -> List[Tuple[int, int]]:
    """Pixels of ``cv::LineIterator(img, p1, p2, 8, leftToRight=true)`` -- what ``cv::Line`` (and through
    it ``polylines`` with thickness 1, shift 0) writes: Bresenham from the left end point, error term
    ``major - 2*minor``, ``major + 1`` pixels.  End points inside the image only (``cv::Line`` clips the
    segment first; the reference's scaled vertices lie in ``[0, W-1] x [0, H-1]`` by construction)."""
    x1, y1, x2, y2 = int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])
    dx, dy = x2 - x1, y2 - y1
    if dx < 0:
        x1, y1, dx, dy = x2, y2, -dx, -dy
    sy = -1 if dy < 0 else 1
    dy = abs(dy)
    steep = dy > dx
    a, b = (dy, dx) if steep else (dx, dy)
    err = a - 2 * b
    x, y = x1, y1
    out = []
    for _ in range(a + 1):
        out.append((x, y))
        minor = err < 0
        err += -2 * b + (2 * a if minor else 0)
        if steep:
            y += sy
            x += 1 if minor else 0
        else:
            x += 1
            y += sy if minor else 0
    return out


def polylines_oracle(polys: List[np.ndarray], output_size, value: int = 255, canvas=None) -> np.ndarray:
    """``for pts in polys: cv2.polylines(canvas, [pts], True, value, 1)`` (reference ``:455-456``)."""
    width, height = output_size
    if canvas is None:
        canvas = np.zeros((height, width), dtype=np.uint8)
    for pts in polys:
        n = len(pts)
        for i in range(n):
            for (x, y) in line_pixels(pts[i - 1], pts[i]):
                if 0 <= x < width and 0 <= y < height:
                    canvas[y, x] = value
    return canvas


def blur_kernel(size: int, blur_type: str) -> np.ndarray:
    """``Pouch._create_blur_kernel`` (reference ``core/pouch.py:494-513``)."""
    if blur_type == "motion":
        k = np.zeros((size, size), np.float32)
        k[size // 2, :] = 1.0 / size
        return k
    return np.ones((size, size), np.float32) / (size * size)


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def filter2d_oracle(src: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """``cv2.filter2D(src, -1, kernel)`` for uint8 ``src`` of shape (H, W) or (H, W, C)."""
    kernel = np.asarray(kernel, dtype=np.float32)
    kh, kw = kernel.shape
    ay, ax = kh // 2, kw // 2
    h, w = src.shape[:2]
    acc = np.zeros(src.shape, dtype=np.float32)
    ys, xs = np.arange(h), np.arange(w)
    for i in range(kh):
        for j in range(kw):
            if kernel[i, j] == 0:
                continue
            yy = _reflect101(ys + i - ay, h)
            xx = _reflect101(xs + j - ax, w)
            tap = src[yy][:, xx].astype(np.float64)
            # one fused multiply-add per tap: the float64 product and sum are exact (32 + 24 bits < 53), so a
            # single rounding to float32 remains
            acc = (acc.astype(np.float64) + np.float64(kernel[i, j]) * tap).astype(np.float32)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def dilate3_oracle(src: np.ndarray) -> np.ndarray:
    """``cv2.dilate(src, np.ones((3, 3), np.uint8), iterations=1)``."""
    h, w = src.shape
    pad = np.zeros((h + 2, w + 2), dtype=src.dtype)
    pad[1:-1, 1:-1] = src
    out = np.zeros_like(src)
    for dy in range(3):
        for dx in range(3):
            out = np.maximum(out, pad[dy:dy + h, dx:dx + w])
    return out


def edge_maps_oracle(polys: List[np.ndarray], output_size, blur_kernel_size: int = 3, blur_type: str = "mean"
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """The two static maps of the branch: ``edges_mask`` and ``edges_dilated`` (reference ``:455-460``)."""
    edges = polylines_oracle(polys, output_size)
    dil = dilate3_oracle(filter2d_oracle(edges, blur_kernel(blur_kernel_size, blur_type)))
    return edges, dil


def edge_effects_oracle(img: np.ndarray, polys: List[np.ndarray], output_size, with_border: bool, edge_blur: bool,
                        blur_kernel_size: int = 3, blur_type: str = "mean") -> np.ndarray:
    """Post-processing of the default-branch image ``img`` (H, W, 2) uint8 (reference ``:458-481``)."""
    img = img.copy()
    if not (with_border or edge_blur):
        return img
    edges = polylines_oracle(polys, output_size)
    if edge_blur:
        k = blur_kernel(blur_kernel_size, blur_type)
        dil = dilate3_oracle(filter2d_oracle(edges, k))
        m = dil / 255.0
        if not with_border:
            blurred = filter2d_oracle(img, k)
            m2 = np.stack([m, m], axis=-1)
            img = (img * (1 - m2) + blurred * m2).astype(np.uint8)
        else:
            img[:, :, 0] = np.where(edges > 0, img[:, :, 0] // 2, img[:, :, 0])
    if with_border and not edge_blur:
        # cv2.polylines(img[:, :, 0], ..., 0, 1): the reference passes a non-contiguous view, which
        # OpenCV >= 4.x rejects; the intended result is gray = 0 on the edges
        img[:, :, 0][edges > 0] = 0
    return img
