"""Offline: measures the symbol frequencies of K3's token streams on sample frames (oracle parse) and writes
calcium_b200/csrc/enc_freq.inc, the data the static Huffman codes are built from."""
import os, sys, heapq, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pouch_oracle as O, encode_oracle as EO
from calcium_b200.ensemble import make_specs, DEFAULT_FRAME_STEPS
O.build()
def tokens(keys, pixel_bytes, bpp, row_lead, head=b""):
    H, W = keys.shape; maxpix = 256 // bpp; up = row_lead + bpp * W
    lit = collections.Counter(); ln = collections.Counter(); ds = collections.Counter(); extra = 0
    for b in head: lit[b] += 1
    for y in range(H):
        row = keys[y]
        eq_up = (row == keys[y-1]) if y > 0 else np.zeros(W, bool)
        eq_left = np.zeros(W, bool); eq_left[1:] = row[1:] == row[:-1]
        mode = np.where(eq_up, 2, np.where(eq_left, 1, 0))
        for _ in range(row_lead): lit[0] += 1
        start = np.zeros(W, bool); start[0] = True; start[1:] = mode[1:] != mode[:-1]; start |= mode == 0; start[np.arange(0, W, maxpix)] = True
        pos = np.flatnonzero(start); ends = np.append(pos[1:], W)
        for s, e in zip(pos, ends):
            n = int(e - s); md = int(mode[s])
            if md and n * bpp >= 3:
                L = n * bpp
                i = max(k for k in range(29) if EO._LEN_BASE[k] <= L); ln[257 + i] += 1; extra += EO._LEN_EXTRA[i]
                D = up if md == 2 else bpp
                j = max(k for k in range(30) if EO._DIST_BASE[k] <= D); ds[j] += 1; extra += EO._DIST_EXTRA[j]
            else:
                for b in pixel_bytes(int(row[s])): lit[b] += 1
    ll = lit + ln; ll[256] += 1
    return ll, ds, extra
def huff_lengths(freq, maxlen=15):
    # simple Huffman, then clamp (approximate); returns dict sym->len
    items = [(f, [s]) for s, f in freq.items() if f > 0]
    if len(items) == 1: return {items[0][1][0]: 1}
    h = [(f, i, syms) for i, (f, syms) in enumerate(items)]; heapq.heapify(h); L = collections.Counter(); c = len(h)
    while len(h) > 1:
        a = heapq.heappop(h); b = heapq.heappop(h)
        for s in a[2] + b[2]: L[s] += 1
        heapq.heappush(h, (a[0] + b[0], c, a[2] + b[2])); c += 1
    return L
specs = make_specs(8, size="medium"); g = specs[0].geometry
limg = O.paint_label_map_cv2(g.vertices, (512, 512)).astype(np.uint16); lmsk = O.cells_mask_oracle(g.vertices, (512, 512)).astype(np.uint16)
zero = np.zeros(g.n_cells)
LLi = collections.Counter(); DSi = collections.Counter(); LLm = collections.Counter(); DSm = collections.Counter(); per = []
for s in specs[:8]:
    fr = O.simulate_c(g.csr, O.pack_params(s.params), zero, zero, s.r0, s.vplc, 18000, DEFAULT_FRAME_STEPS)
    for f in (8, 30, 52, 74):
        img, _, inst, nact = O.raster_frame_c(fr[f,0], limg, lmsk, 0.1, True, want_active=False)
        keys = img[...,0].astype(np.int64) | ((img[...,1] != 0).astype(np.int64) << 8)
        ll, ds, ex = tokens(keys, lambda k: [(k&255) >> 6, ((k&255) << 2) & 255, 255 if k >> 8 else 0, 255 if k >> 8 else 0], 4, 1)
        llm, dsm, exm = tokens(inst.astype(np.int64), lambda k: [k & 255, k >> 8], 2, 0, head=EO.npy_bytes(inst)[:128])
        per.append((ll, ds, ex, llm, dsm, exm, nact)); LLi += ll; DSi += ds; LLm += llm; DSm += dsm
def cost(ll, ds, ex, Lll, Lds):
    return sum(f * Lll.get(s, 15) for s, f in ll.items()) + sum(f * Lds.get(s, 15) for s, f in ds.items()) + ex
def fixed_cost(ll, ds, ex):
    return sum(f * EO.fixed_symbol(s)[1] for s, f in ll.items()) + sum(f * 5 for f in ds.values()) + ex
Li, Ldi, Lm, Ldm = huff_lengths(LLi), huff_lengths(DSi), huff_lengths(LLm), huff_lengths(DSm)
print("max code len img", max(Li.values()), "mask", max(Lm.values()), "n syms img", len(Li), "mask", len(Lm))
fi = sum(fixed_cost(p[0], p[1], p[2]) for p in per) / 8 / len(per); ti = sum(cost(p[0], p[1], p[2], Li, Ldi) for p in per) / 8 / len(per)
fm = sum(fixed_cost(p[3], p[4], p[5]) for p in per if p[6]) / 8 / len(per); tm = sum(cost(p[3], p[4], p[5], Lm, Ldm) for p in per if p[6]) / 8 / len(per)
print(f"image: fixed {fi:.0f} B  tuned-static {ti:.0f} B ({ti/fi:.2f});  mask: fixed {fm:.0f} B tuned {tm:.0f} B ({tm/max(fm,1):.2f})")
print("img extra bits share", sum(p[2] for p in per)/8/len(per))
# ---- emit frequency tables (coarse: scaled so that the largest is 1<<16, floor 1) ----
def table(cnt, n):
    mx = max(cnt.values())
    return [max(1, int(round(cnt.get(s, 0) * 65536.0 / mx))) for s in range(n)]
out = []
for name, cnt, n in (("kFreqImgLL", LLi, 286), ("kFreqImgD", DSi, 30), ("kFreqMskLL", LLm, 286), ("kFreqMskD", DSm, 30)):
    t = table(cnt, n)
    out.append(f"static const uint32_t {name}[{n}] = {{\n" + "\n".join("    " + ", ".join(str(v) for v in t[i:i+16]) + "," for i in range(0, n, 16)) + "\n};\n")
open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "calcium_b200", "csrc", "enc_freq.inc"), "w").write(
"// enc_freq.inc -- symbol frequencies K3's static Huffman codes are built from (DEFLATE literal/length and\n"
"// distance alphabets, image and mask stream), measured on 32 frames of medium pouches of the four sim\n"
"// types with the token parse of encode.cu, scaled to a maximum of 65536 with a floor of 1 (every symbol\n"
"// keeps a code, so the tables are valid for any input).  DATA, not logic: oracle/encode_oracle.py reads the\n"
"// same numbers.  Regenerate with scripts/make_enc_freq.py.\n" + "\n".join(out))
print("written")
