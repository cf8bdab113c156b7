"""TEST INFRASTRUCTURE -- CPU restatement of K3's file encoder (calcium_b200/csrc/encode.cu).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
never does.

What it pins, and against what:
  * the *content* of the files is pinned to the reference: decoding the PNG gives the array the
    reference builds in utils/image_processing.py:256-265 (gray*4 as big-endian uint16, alpha*257)
    from Pouch.generate_image's output, and ``np.load(npz)['mask']`` gives the array
    utils/sam2_data_generator.py:203 stores -- both checked with independent decoders (zlib, numpy's
    own npz reader, cv2.imdecode);
  * the *bytes* of the files are this repo's own choice (any DEFLATE parse is a valid PNG/zip), so
    the byte-exact comparison is GPU encoder vs this restatement of the same parse rules, written
    independently in numpy with zlib's crc32/adler32 as the checksum authority.

Parse rules (the same as encode.cu): rows are cut into 512-pixel segments; per pixel the mode is
"up" when it equals the pixel above, else "left" when it equals the previous pixel of the row, else
literal; a token is a maximal run of one mode, additionally cut every 256 bytes from the segment
start; a match shorter than 3 bytes is sent as a literal pixel; one final block.

Huffman code: ``"fixed"`` (RFC 1951 3.2.6) or ``"tuned"`` -- a dynamic-Huffman block whose code
lengths are static per stream kind: Huffman lengths (<= 15 bits) over the symbol frequencies in
``calcium_b200/csrc/enc_freq.inc`` (shared DATA; the two "distance" weights go to the symbols of the
left / up distance of the actual image width, every other distance symbol weighs 1), built with a
priority queue ordered by (weight, id) -- leaves 0..n-1, internal nodes numbered in creation order --
halving all weights (rounding up) while the deepest leaf exceeds the limit; canonical codes; the
block header sends all 316 code lengths one by one (no repeat codes) with a <= 7-bit code-length code
built the same way over the length values in use.
"""
from __future__ import annotations

import struct
import zlib
from typing import List, Tuple

import numpy as np

_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
              4097, 6145, 8193, 12289, 16385, 24577]
_DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def _rev(v: int, n: int) -> int:
    return int(format(v, f"0{n}b")[::-1], 2)


def fixed_symbol(sym: int) -> Tuple[int, int]:
    """(bits in stream order, count) of literal/length symbol ``sym`` in the fixed Huffman code (RFC 1951 3.2.6)."""
    if sym <= 143:
        code, nb = 0b00110000 + sym, 8
    elif sym <= 255:
        code, nb = 0b110010000 + (sym - 144), 9
    elif sym <= 279:
        code, nb = sym - 256, 7
    else:
        code, nb = 0b11000000 + (sym - 280), 8
    return _rev(code, nb), nb


def length_token(length: int, code=None) -> Tuple[int, int]:
    i = max(k for k in range(29) if _LEN_BASE[k] <= length)
    v, nb = fixed_symbol(257 + i) if code is None else code.ll[257 + i]
    return v | ((length - _LEN_BASE[i]) << nb), nb + _LEN_EXTRA[i]


def dist_symbol(dist: int) -> int:
    return max(k for k in range(30) if _DIST_BASE[k] <= dist)


def dist_token(dist: int, code=None) -> Tuple[int, int]:
    i = dist_symbol(dist)
    v, nb = (_rev(i, 5), 5) if code is None else code.d[i]
    return v | ((dist - _DIST_BASE[i]) << nb), nb + _DIST_EXTRA[i]


def _freq_tables():
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "calcium_b200", "csrc", "enc_freq.inc")
    text = open(path).read()
    out = {}
    for name in ("kFreqImgLL", "kFreqImgD", "kFreqMskLL", "kFreqMskD"):
        body = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, text, re.S).group(1)
        out[name] = [int(x) for x in re.findall(r"\d+", body)]
    return out


def huffman_lengths(weights: List[int], maxlen: int) -> List[int]:
    import heapq
    f = list(weights)
    n = len(f)
    if n == 1:
        return [1]
    while True:
        heap = [(w, i) for i, w in enumerate(f)]
        heapq.heapify(heap)
        parent = {}
        nxt = n
        while len(heap) > 1:
            a = heapq.heappop(heap)
            b = heapq.heappop(heap)
            parent[a[1]] = parent[b[1]] = nxt
            heapq.heappush(heap, (a[0] + b[0], nxt))
            nxt += 1
        lens = []
        for i in range(n):
            d, p = 0, i
            while p in parent:
                p = parent[p]
                d += 1
            lens.append(d)
        if max(lens) <= maxlen:
            return lens
        f = [(w + 1) // 2 for w in f]


def canonical(lens: List[int]) -> List[Tuple[int, int]]:
    """(stream-order bits, count) per symbol from code lengths (RFC 1951 3.2.2)."""
    count = [0] * 16
    for ln in lens:
        count[ln] += 1
    count[0] = 0
    code, nxt = 0, [0] * 16
    for b in range(1, 16):
        code = (code + count[b - 1]) << 1
        nxt[b] = code
    out = []
    for ln in lens:
        if ln:
            out.append((_rev(nxt[ln], ln), ln))
            nxt[ln] += 1
        else:
            out.append((0, 0))
    return out


class TunedCode:
    """The static dynamic-Huffman code of one stream kind ('img' | 'msk') at a given left / up distance."""

    def __init__(self, kind: str, dist_left: int, dist_up: int):
        fr = _freq_tables()
        ll_w = fr["kFreqImgLL" if kind == "img" else "kFreqMskLL"]
        d2 = fr["kFreqImgD" if kind == "img" else "kFreqMskD"]
        d_w = [1] * 30
        d_w[dist_symbol(dist_left)] = d2[0]
        if 1 <= dist_up <= 32768:
            d_w[dist_symbol(dist_up)] = d2[1]
        ll_len, d_len = huffman_lengths(ll_w, 15), huffman_lengths(d_w, 15)
        self.ll, self.d = canonical(ll_len), canonical(d_len)
        # block header: BFINAL=1, BTYPE=10, HLIT, HDIST, HCLEN, code-length code, the 316 lengths
        hdr: List[Tuple[int, int]] = [(1, 1), (2, 2), (286 - 257, 5), (30 - 1, 5), (19 - 4, 4)]
        every = ll_len + d_len
        used = sorted(set(every))
        weights = [every.count(v) for v in used]
        if len(used) == 1:
            used = sorted(used + [1 if used[0] == 0 else 0])
            weights = [every.count(v) or 1 for v in used]
        ul = huffman_lengths(weights, 7)
        cl_len = [0] * 19
        for v, ln in zip(used, ul):
            cl_len[v] = ln
        cl = canonical(cl_len)
        for k in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15):
            hdr.append((cl_len[k], 3))
        hdr += [cl[v] for v in every]
        self.header = hdr


def _pack(tokens: List[Tuple[int, int]]) -> bytes:
    """Concatenate (value, bit count) pairs LSB first."""
    vals = np.array([t[0] for t in tokens], dtype=np.uint64)
    nbs = np.array([t[1] for t in tokens], dtype=np.int64)
    k = np.arange(64, dtype=np.uint64)
    bits = ((vals[:, None] >> k[None, :]) & np.uint64(1)).astype(np.uint8)
    keep = np.arange(64)[None, :] < nbs[:, None]
    return np.packbits(bits[keep], bitorder="little").tobytes()


def deflate_rows(keys: np.ndarray, pixel_bytes, bpp: int, row_lead: int, head: bytes = b"", kind: str = "img",
                 huffman: str = "tuned") -> bytes:
    """Raw DEFLATE stream of ``head`` + rows of pixels.  ``keys`` (H, W) integers identify pixel
    values; ``pixel_bytes(key)`` gives the ``bpp`` bytes of a pixel; every row is preceded by
    ``row_lead`` zero bytes (the PNG filter byte)."""
    H, W = keys.shape
    maxpix = 256 // bpp
    up_dist = row_lead + bpp * W
    use_up = up_dist <= 32768
    code = TunedCode(kind, bpp, up_dist) if huffman == "tuned" else None
    lit = [fixed_symbol(b) if code is None else code.ll[b] for b in range(256)]
    tok: List[Tuple[int, int]] = [(3, 3)] if code is None else list(code.header)   # fixed: BFINAL=1, BTYPE=01
    tok += [lit[b] for b in head]
    d_left, d_up = dist_token(bpp, code), (dist_token(up_dist, code) if use_up else (0, 0))

    def literal_pixel(key):
        v, nb = 0, 0
        for b in pixel_bytes(int(key)):
            v |= lit[b][0] << nb
            nb += lit[b][1]
        return v, nb

    for y in range(H):
        row = keys[y]
        eq_up = (row == keys[y - 1]) if (y > 0 and use_up) else np.zeros(W, bool)
        eq_left = np.zeros(W, bool)
        eq_left[1:] = row[1:] == row[:-1]
        mode = np.where(eq_up, 2, np.where(eq_left, 1, 0))
        for _ in range(row_lead):
            tok.append(lit[0])
        for x0 in range(0, W, 512):
            x1 = min(W, x0 + 512)
            m = mode[x0:x1]
            n = x1 - x0
            start = np.zeros(n, bool)
            start[0] = True
            start[1:] = m[1:] != m[:-1]
            start |= m == 0
            start[np.arange(0, n, maxpix)] = True
            pos = np.flatnonzero(start)
            ends = np.append(pos[1:], n)
            for s, e in zip(pos, ends):
                ln = int(e - s)
                md = int(m[s])
                if md != 0 and ln * bpp >= 3:
                    lv, lnb = length_token(ln * bpp, code)
                    dv, dnb = d_up if md == 2 else d_left
                    tok.append((lv | (dv << lnb), lnb + dnb))
                else:
                    assert ln == 1
                    tok.append(literal_pixel(row[x0 + s]))
    tok.append(fixed_symbol(256) if code is None else code.ll[256])   # end of block
    return _pack(tok)


def _png_chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def png_scanlines(image: np.ndarray) -> bytes:
    """Filter-0 scanlines of the 16-bit gray+alpha picture the reference builds from an (H, W, 2)
    uint8 frame (utils/image_processing.py:256-265): gray*4, alpha*257, big endian."""
    h, w, _ = image.shape
    px = np.empty((h, w, 2), dtype=">u2")
    px[..., 0] = image[..., 0].astype(np.uint16) * 4
    px[..., 1] = image[..., 1].astype(np.uint16) * 257
    raw = np.zeros((h, 1 + 4 * w), dtype=np.uint8)
    raw[:, 1:] = px.view(np.uint8).reshape(h, 4 * w)
    return raw.tobytes()


def encode_png(image: np.ndarray, huffman: str = "tuned") -> bytes:
    """The .png K3 writes for ``image`` = Pouch.generate_image's (H, W, 2) uint8 array."""
    h, w, _ = image.shape
    keys = image[..., 0].astype(np.int64) | ((image[..., 1] != 0).astype(np.int64) << 8)

    def pixel_bytes(key):
        g, a = key & 255, 255 if key >> 8 else 0
        return [g >> 6, (g << 2) & 255, a, a]

    z = b"\x78\x01" + deflate_rows(keys, pixel_bytes, 4, 1, kind="img", huffman=huffman) + struct.pack(">I", zlib.adler32(png_scanlines(image)))
    return (b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 4, 0, 0, 0))
            + _png_chunk(b"IDAT", z) + _png_chunk(b"IEND", b""))


def npy_bytes(mask: np.ndarray) -> bytes:
    """.npy (format 1.0) bytes of a C-contiguous uint16 array, as numpy.lib.format writes them."""
    h, w = mask.shape
    d = "{'descr': '<u2', 'fortran_order': False, 'shape': (%d, %d), }" % (h, w)
    pad = (64 - ((10 + len(d) + 1) % 64)) % 64
    hdr = d + " " * pad + "\n"
    return b"\x93NUMPY\x01\x00" + struct.pack("<H", len(hdr)) + hdr.encode("latin1") + mask.astype("<u2").tobytes()


def encode_npz(mask: np.ndarray, huffman: str = "tuned") -> bytes:
    """The .npz K3 writes for an instance mask: a zip archive with the single deflated member
    ``mask.npy`` (what np.savez_compressed(path, mask=mask) stores)."""
    h, w = mask.shape
    raw = npy_bytes(mask)
    head = raw[: len(raw) - 2 * h * w]
    comp = deflate_rows(mask.astype(np.int64), lambda k: [k & 255, k >> 8], 2, 0, head=head, kind="msk", huffman=huffman)
    crc = zlib.crc32(raw) & 0xFFFFFFFF
    name = b"mask.npy"
    local = struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, 8, 0, 0x21, crc, len(comp), len(raw), len(name), 0) + name
    central = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0, 8, 0, 0x21, crc, len(comp), len(raw),
                          len(name), 0, 0, 0, 0, 0x01800000, 0) + name
    end = struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 1, 1, len(central), len(local) + len(comp), 0)
    return local + comp + central + end


def decode_png(data: bytes) -> np.ndarray:
    """Independent decoder for the PNG subset above -> (H, W, 2) uint16 (gray16, alpha16).  Checks
    every chunk CRC and lets zlib check the Adler-32."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        (ln,), tag = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        (crc,) = struct.unpack(">I", data[pos + 8 + ln:pos + 12 + ln])
        assert crc == (zlib.crc32(tag + body) & 0xFFFFFFFF), f"bad CRC in {tag!r}"
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, interlace) == (16, 4, 0)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + ln
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    ftype = raw[:, 0]
    assert np.isin(ftype, (0, 2)).all()                      # K3 writes filter type 0, K3r type 2 (Up)
    px = raw[:, 1:].copy()
    for r in range(h):
        if ftype[r] == 2 and r > 0:
            px[r] += px[r - 1]                                # uint8 arithmetic wraps mod 256
    return px.view(">u2").reshape(h, w, 2).astype(np.uint16)


# ---------------------------------------------------------------------------------------------------------------
# K3r (calcium_b200/csrc/encode_raw.cu): PNG of an arbitrary (H, W, 2) uint8 gray + alpha frame.  The same parse,
# restated with plain Python loops (small frames only): filter type 2 on every scanline, fixed-Huffman DEFLATE with
# distance-1 run matches, a lane = a span of ceil(W / 32) pixels, one block per pass of scanlines closed by empty
# stored blocks so that the file length stays a multiple of 4.  zlib is the checksum authority.
# ---------------------------------------------------------------------------------------------------------------
_RAW_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_RAW_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def raw_rows_per_pass(H: int, W: int) -> int:
    """Scanlines per pass = per DEFLATE block: what fits the kernel's 100 KB of shared memory (its layout_for)."""
    ppl = (W + 31) // 32
    seg = ppl + 1 + (ppl & 1)
    for rpp in range(16, 0, -1):
        bits = 43 * 8 + 3 + rpp * (4 * W + 1) * 9 + 7 + 3
        words = (bits + 31) // 32 + 8
        if (256 + 260 + 256) * 4 + rpp * 32 * seg * 4 + words * 4 <= 100 * 1024:
            return rpp
    raise ValueError("frame too wide")


class _BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, value: int, nbits: int) -> None:          # LSB first
        self.acc |= value << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 0xff)
            self.acc >>= 8
            self.n -= 8

    def huff(self, code: int, nbits: int) -> None:           # Huffman codes go in MSB first
        self.put(int(format(code, f"0{nbits}b")[::-1], 2), nbits)

    def align(self) -> None:
        if self.n:
            self.put(0, 8 - self.n)


def _raw_literal(bw: _BitWriter, b: int) -> None:
    if b <= 143:
        bw.huff(0x30 + b, 8)
    else:
        bw.huff(0x190 + b - 144, 9)


def _raw_run(bw: _BitWriter, length: int) -> None:          # `length` more copies of the previous byte (distance 1)
    i = max(k for k in range(29) if _RAW_LEN_BASE[k] <= length)
    sym = 257 + i
    if sym <= 279:
        bw.huff(sym - 256, 7)
    else:
        bw.huff(0xC0 + sym - 280, 8)
    bw.put(length - _RAW_LEN_BASE[i], _RAW_LEN_EXTRA[i])
    bw.huff(0, 5)


def _raw_flush(bw: _BitWriter, value: int, run: int) -> None:
    while run >= 3:
        m = min(run, 258)
        _raw_run(bw, m)
        run -= m
    for _ in range(run):
        _raw_literal(bw, value)


def encode_raw_png(image: np.ndarray) -> bytes:
    """The file ``cab200_encode_raw_batch`` writes for ``image`` ((H, W, 2) uint8), byte for byte."""
    H, W = image.shape[:2]
    g = image[..., 0].astype(np.uint16)
    al = image[..., 1].astype(np.uint16)
    px = np.stack([g >> 6, (g << 2) & 0xff, al, al], axis=-1).astype(np.uint8)            # (H, W, 4) stream bytes
    up = np.concatenate([np.zeros((1, W, 4), np.uint8), px[:-1]])
    filt = (px.astype(np.int16) - up.astype(np.int16)).astype(np.uint8).reshape(H, 4 * W)
    raw = np.concatenate([np.full((H, 1), 2, np.uint8), filt], axis=1)                      # what zlib would be fed
    rpp = raw_rows_per_pass(H, W)
    ppl = (W + 31) // 32
    ihdr = struct.pack(">IIBBBBB", W, H, 16, 4, 0, 0, 0)
    head = (b"\x89PNG\r\n\x1a\n" + struct.pack(">I", 13) + b"IHDR" + ihdr + struct.pack(">I", zlib.crc32(b"IHDR" + ihdr))
            + b"\0\0\0\0" + b"IDAT" + b"\x78\x01")
    assert len(head) == 43
    body = bytearray(head)
    for r0 in range(0, H, rpp):
        bw = _BitWriter()
        bw.put(2, 3)                                          # BFINAL = 0, BTYPE = 01
        for r in range(r0, min(H, r0 + rpp)):
            row = filt[r]
            for lane in range(32):
                x0, x1 = lane * ppl, min(W, lane * ppl + ppl)
                if lane == 0:
                    _raw_literal(bw, 2)
                    prev = 2
                elif x1 > x0:
                    prev = int(row[4 * x0 - 1])
                else:
                    continue
                run = 0
                for b in row[4 * x0:4 * x1].tolist():
                    if b == prev:
                        run += 1
                    else:
                        _raw_flush(bw, prev, run)
                        run = 0
                        _raw_literal(bw, b)
                        prev = b
                _raw_flush(bw, prev, run)
        bw.put(0, 7)                                          # end of block
        bw.put(0, 3)                                          # empty stored block: aligns to a byte
        bw.align()
        bw.out += b"\x00\x00\xff\xff"
        while (len(body) + len(bw.out)) % 4:
            bw.out += b"\x00\x00\x00\xff\xff"
        body += bw.out
    body += b"\x01\x00\x00\xff\xff" + struct.pack(">I", zlib.adler32(raw.tobytes()))
    data_end = len(body)
    body[33:37] = struct.pack(">I", data_end - 41)
    body += struct.pack(">I", zlib.crc32(bytes(body[37:data_end])))
    body += struct.pack(">I", 0) + b"IEND" + struct.pack(">I", zlib.crc32(b"IEND"))
    return bytes(body)
