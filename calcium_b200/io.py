"""Host-side file writers for the batch path (first cut).

The reference intends ``images/<split>/<name>.png`` to be a 10-bit-in-16-bit gray+alpha PNG
(``utils/image_processing.py:252-269``: gray*4, alpha*257) but hands the two-channel array to
``cv2.imwrite``, which OpenCV rejects (SURVEY.md defects D2/D3).  Here the same samples are written
as a real PNG of colour type 4 (gray + alpha, 16 bit), with a small encoder on top of zlib.
Instance masks keep the reference's container: ``np.savez_compressed(path, mask=...)``
(``utils/sam2_data_generator.py:183-207``), which ``utils/sam2_dataloader.py`` reads back.
"""
from __future__ import annotations

import csv
import os
import struct
import zlib
from typing import List, Tuple

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_gray_alpha_png16(gray16: np.ndarray, alpha16: np.ndarray, level: int = 1) -> bytes:
    h, w = gray16.shape
    px = np.empty((h, w, 2), dtype=">u2")
    px[..., 0] = gray16
    px[..., 1] = alpha16
    raw = np.zeros((h, 1 + w * 4), dtype=np.uint8)       # filter byte 0 per scanline
    raw[:, 1:] = px.view(np.uint8).reshape(h, w * 4)
    ihdr = struct.pack(">IIBBBBB", w, h, 16, 4, 0, 0, 0)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr)
            + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))


def save_gray_alpha_png(image: np.ndarray, path: str, bit_depth: int = 10) -> str:
    """``image``: (H, W, 2) uint8 gray + alpha.  Gray is scaled 8 -> ``bit_depth`` bits by a left
    shift (x4 for 10 bit, as the reference), alpha to full 16 bit (x257)."""
    if image.ndim != 3 or image.shape[2] != 2 or image.dtype != np.uint8:
        raise ValueError(f"expected (H, W, 2) uint8, got {image.shape} {image.dtype}")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    gray = image[:, :, 0].astype(np.uint16) << max(0, bit_depth - 8)
    alpha = image[:, :, 1].astype(np.uint16) * 257
    with open(path, "wb") as fh:
        fh.write(encode_gray_alpha_png16(gray, alpha))
    return path


def save_instance_mask(instance_mask: np.ndarray, path: str) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, mask=instance_mask)
    return path


def save_csv_mapping(mappings: List[Tuple[str, str]], output_path: str, filename: str = "train.csv") -> str:
    """``ImageId,MaskId`` rows -- the format of reference ``utils/labeling.py:289-318``."""
    os.makedirs(output_path, exist_ok=True)
    if not filename.lower().endswith(".csv"):
        filename += ".csv"
    file_path = os.path.join(output_path, filename)
    with open(file_path, "w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow(["ImageId", "MaskId"])
        wr.writerows(mappings)
    return file_path


def scan_image_mask_pairs(batch_dir: str, split: str = "train") -> List[Tuple[str, str]]:
    """``(image, mask)`` basenames of every ``images/<split>/X.png`` that has a ``masks/<split>/X.npz`` --
    what the reference intends its CSV generator to find by scanning the dataset directory
    (``utils/generate_csv.py:32-55``; its own glob never matches the current layout).  Scanning
    instead of remembering makes the mapping cover every batch written into the same directory."""
    img_dir = os.path.join(batch_dir, "images", split)
    mask_dir = os.path.join(batch_dir, "masks", split)
    if not (os.path.isdir(img_dir) and os.path.isdir(mask_dir)):
        return []
    masks = {os.path.splitext(f)[0] for f in os.listdir(mask_dir) if f.endswith(".npz")}
    return [(f, os.path.splitext(f)[0] + ".npz") for f in sorted(os.listdir(img_dir))
            if f.endswith(".png") and os.path.splitext(f)[0] in masks]
