// encode.cu -- K3: rasterise AND losslessly encode every (pouch, frame) on the device, so that what
// crosses PCIe is the file the reference's batch path stores, not the raw pixels.
//
// Replaces, per sampled frame of the batch path (reference main.py:270-330):
//   Pouch.generate_image + save_image(..., format='png')   core/pouch.py:412-453,
//        utils/image_processing.py:252-269  (gray*4 | alpha*257, 16-bit gray+alpha PNG)
//   process_batch_for_sam2's mask branch                   utils/sam2_data_generator.py:315-335
//        (get_active_cells, get_cell_masks(active_only), create_instance_mask_from_cells,
//         np.savez_compressed(path, mask=instance_mask); no file when no cell is active)
// Output per frame: one complete .png byte string and one complete .npz byte string in a device
// arena, plus a record (offset, size).  A 512x512 frame is 1.5 MiB of raw pixels; encoded it is
// typically 30-60 KiB, which is what makes the end-to-end path compute- instead of PCIe-bound.
//
// One CTA per (pouch, frame).  It builds the per-frame look-up tables (frame_lut.cuh), then encodes
// the two streams one after the other into a shared-memory staging buffer:
//   * the pixel rows are never materialised: a warp handles one 512-pixel row segment.  The fast path parses
//     only the segment's static EVENTS (the pixels whose label differs from the one above, plus the forced cuts;
//     1 to K per lane, chosen per row) -- every other pixel is in mode "up" whatever the frame's values; rows of
//     the mask without any instance id are written without reading even those.  The general path (K = 0) reads
//     16 consecutive pixels per lane from the static label map (L2 resident) through the LUT;
//   * LZ77 parse restricted to two match distances: the previous pixel (runs inside a cell) and the
//     pixel above (rows of flat-shaded polygons mostly repeat the row above).  Per pixel the mode is
//     "up" if it equals the pixel above, else "left" if it equals the previous pixel, else literal;
//     a token is a maximal run of one mode, cut at multiples of 256 bytes and at segment ends;
//   * DEFLATE with a static Huffman code per stream kind (tuned, sent as a dynamic-block header; or the fixed
//     code), one final block: token bit lengths are known from small
//     tables, a warp scan + a CTA scan over the 16 rows of a pass give every token its bit
//     position, and the bits are OR-ed into the zeroed staging buffer with shared-memory atomics;
//   * checksums on the fly: Adler-32 of the PNG scanlines (plain sums), CRC-32 of the .npy bytes
//     (per-lane table CRC, moved to its place in the message by multiplication with x^(8k) mod P in
//     GF(2)[x] -- CRC is linear), CRC-32 of the finished IDAT chunk (pieces combined the same way);
//   * the finished file is copied to the arena at a cursor advanced with one atomicAdd.
// A frame whose encoding does not fit the staging buffer is flagged; the host falls back to K2 + a
// host encoder for it (never observed for polygon images; covered by a test with a tiny buffer).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "frame_lut.cuh"

namespace cab200 {

constexpr int ENC_NT = 512;
constexpr int ENC_NW = ENC_NT / 32;
constexpr int ENC_MAXDIM = 2048;
constexpr int ENC_NPIECE = 4096;         // 64-byte pieces of a staged file: 256 KB, more than an SM's shared memory
constexpr uint32_t CRC_POLY = 0xEDB88320u;

struct EncStream {
    uint32_t lit[256];       // literal: reversed Huffman code | bits << 24
    uint32_t len[260];       // match length 3..258: reversed code + extra bits | bits << 24
    uint32_t hdr[96];        // DEFLATE block header (BFINAL, BTYPE and, for a dynamic block, the code lengths), LSB first
    int hdr_bits;
    uint32_t eob;            // end-of-block symbol: reversed code | bits << 24
    uint8_t prefix[64];
    uint8_t suffix[96];
    uint8_t head[192];       // uncompressed bytes in front of the pixel rows (.npy header), sent as literals
    uint16_t head_off[192];  // bit offset of head byte j's literal; head_bits = all of them
    int head_bits;
    int prefix_len, suffix_len, head_len;
    uint32_t dist_left, dist_up;   // distance code + extra bits (low 24 bits) | bit count << 24; count 0 = unusable
    int bpp;                 // bytes per pixel in the uncompressed stream
    int seg_lanes;           // a forced token start every seg_lanes lanes (16 pixels each): 256 bytes
    int row_lead;            // PNG: one filter byte (0) in front of every row
    uint32_t crc_const;      // npz: CRC-32 terms that do not depend on the pixels
    uint32_t raw_len;        // length of the uncompressed stream
    int p_len_prefix;        // PNG: position of the big-endian IDAT length in the prefix
    int p_crc_prefix, p_csize_prefix;                    // zip local header fields (little endian)
    int p_crc_suffix, p_csize_suffix, p_cdoff_suffix;    // zip central directory / end record
};

struct EncTables {
    uint32_t crc[256];       // reflected CRC-32 table
    uint32_t pow8[32];       // x^(8 * 2^k) mod P
    uint32_t rowshift[ENC_MAXDIM];        // mask: x^(8 * bytes after row y)
    uint32_t colshift[ENC_MAXDIM / 16];   // mask: x^(8 * bytes after 16-pixel chunk c inside its row)
    uint32_t pow64[ENC_NPIECE];           // x^(8 * 64 * j): the IDAT CRC is combined from 64-byte pieces counted from the end
    uint32_t pow1[64];                    // x^(8 * k), k < 64
    EncStream img, msk;
    int H, W;
};

struct EncRecord {
    unsigned long long offset;
    uint32_t size;
    uint32_t flags;
};

struct EncodeArgs {
    const PouchEntry *pouches;
    const uint16_t *label_img, *label_mask;
    const double *ca_frames;
    const EncTables *tab;
    const unsigned char *geom_tabs;   // per-geometry checksum tables (enc_geometry_kernel)
    const long long *geom_tab_off;    // [G] byte offsets into geom_tabs
    int n_frames, H, W;
    double threshold;
    int quirk, want_img, want_msk;
    int lut_bytes, stage_cap;
    uint8_t *arena;
    unsigned long long *cursor;
    unsigned long long arena_bytes;
    EncRecord *records;      // [P*F][2]: image, mask
    int32_t *n_active_out;
};

// ---- GF(2)[x] mod the CRC-32 polynomial, reflected representation (bit 31 = x^0) ------------------
__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll
    for (int i = 31; i >= 0; --i) {
        p ^= b & (0u - ((a >> i) & 1u));
        b = (b >> 1) ^ (CRC_POLY & (0u - (b & 1u)));
    }
    return p;
}

// x^(8*n) from the table of x^(8*2^k)
__device__ __forceinline__ uint32_t gf2_xpow8(const uint32_t *pow8, uint32_t n)
{
    uint32_t r = 0x80000000u;   // x^0
    for (int k = 0; n; ++k, n >>= 1)
        if (n & 1u) r = gf2_mulmod(r, pow8[k]);
    return r;
}

__device__ __forceinline__ void put_bits(uint32_t *w, uint32_t pos, unsigned long long val, int nb)
{
    const uint32_t idx = pos >> 5, sh = pos & 31u;
    const unsigned long long v0 = val << sh;
    atomicOr(&w[idx], (uint32_t)v0);
    if (sh + nb > 32) atomicOr(&w[idx + 1], (uint32_t)(v0 >> 32));
    if (sh + nb > 64) atomicOr(&w[idx + 2], (uint32_t)(val >> (64 - sh)));
}

// Table look-ups through explicit 32-bit shared addresses: with generic pointers the compiler rebuilds the
// shared window base (S2R + LEA + ...) in front of every predicated load -- 6 instructions per look-up.
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;                                  // a wider destination is zero-extended: no PRMT afterwards
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void put_le32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ void put_be32(uint8_t *p, uint32_t v)
{
    p[3] = (uint8_t)v; p[2] = (uint8_t)(v >> 8); p[1] = (uint8_t)(v >> 16); p[0] = (uint8_t)(v >> 24);
}

// Static per-(geometry, H, W) checksum tables, one block per geometry:
//   Q[n+1] u32 | cnt[n+1] u32 | (8-byte aligned) M[n+1] u64      (see encode_stream's finalisation)
struct GeomTab {
    const uint32_t *Q, *cnt;
    const unsigned long long *M;
    const uint2 *ev_img, *ev_msk;      // static event lists (enc_events_kernel), [H * nseg][ENC_ECAP] each
    const uint32_t *span;              // [n+1][2] rows a mask label touches: {max (H-1-y), max y}; both 0 = absent or row 0 only
    const uint16_t *cnt_img, *cnt_msk; // [H * nseg] events of each row segment (capped at ENC_ECAP)
};
constexpr int ENC_ECAP = 256;
__host__ __device__ inline size_t geom_tab_sums_bytes(int n) { return ((((size_t)(n + 1) * 8 + 7) & ~(size_t)7) + (size_t)(n + 1) * 8 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t geom_tab_events_bytes(int H, int W) { return (size_t)H * ((W + 511) / 512) * ENC_ECAP * sizeof(uint2); }
__host__ __device__ inline size_t geom_tab_span_bytes(int n) { return ((size_t)(n + 1) * 8 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t geom_tab_rowcnt_bytes(int H, int W) { return ((size_t)H * ((W + 511) / 512) * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t geom_tab_bytes(int n, int H, int W)
{   // sums | two event lists | two event counters | row spans | two per-row event counts
    return geom_tab_sums_bytes(n) + 2 * geom_tab_events_bytes(H, W) + 16 + geom_tab_span_bytes(n) + 2 * geom_tab_rowcnt_bytes(H, W);
}
__device__ __forceinline__ GeomTab geom_tab_carve(const unsigned char *p, int n, int H, int W)
{
    GeomTab t;
    t.Q = reinterpret_cast<const uint32_t *>(p);
    t.cnt = t.Q + (n + 1);
    t.M = reinterpret_cast<const unsigned long long *>(p + (((size_t)(n + 1) * 8 + 7) & ~(size_t)7));
    t.ev_img = reinterpret_cast<const uint2 *>(p + geom_tab_sums_bytes(n));
    t.ev_msk = reinterpret_cast<const uint2 *>(p + geom_tab_sums_bytes(n) + geom_tab_events_bytes(H, W));
    t.span = reinterpret_cast<const uint32_t *>(p + geom_tab_sums_bytes(n) + 2 * geom_tab_events_bytes(H, W) + 16);
    t.cnt_img = reinterpret_cast<const uint16_t *>(p + geom_tab_sums_bytes(n) + 2 * geom_tab_events_bytes(H, W) + 16 + geom_tab_span_bytes(n));
    t.cnt_msk = t.cnt_img + geom_tab_rowcnt_bytes(H, W) / 2;
    return t;
}

// Static event list of one label map: per 512-pixel row segment the pixels whose label differs from the
// left or the upper neighbour, plus the forced cut positions (every `maxpix` pixels), ENC_ECAP slots:
// x = {x | o << 16}, y = {o_left | o_up << 16}, o = BYTE OFFSET of the label's entry in the frame's 16-bit key
// LUT (2 * label); a missing neighbour and the padding slots (x = segment end) point at entry n + 1, which
// encode_kernel fills with a key no pixel has -- the look-ups and comparisons of the parse need no special
// cases.  Every other pixel has the label of both neighbours, hence -- whatever a frame's values -- the mode
// "up" (or "left" where no row above is usable).  One warp per segment; max_count receives the largest number
// of events in a segment (a list is truncated at ENC_ECAP: the caller then must not use it).
__global__ void __launch_bounds__(256) enc_events_kernel(const uint16_t *label, int H, int W, int n, int maxpix, int use_up, int one_px_literal,
                                                         uint2 *events, uint16_t *row_count, int *max_count)
{
    const int lane = threadIdx.x & 31;
    const int nseg = (W + 511) / 512;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= H * nseg) return;
    const int y = item / nseg, x0 = (item % nseg) * 512, xend = min(W, x0 + 512);
    uint2 *out = events + (size_t)item * ENC_ECAP;
    const uint32_t none = 2u * (uint32_t)(n + 1);
    int count = 0;
    for (int c = x0; c < xend; c += 32) {
        const int x = c + lane;
        const bool inb = x < xend;
        const uint32_t lab = inb ? label[(size_t)y * W + x] : 0u;
        const uint32_t labL = (inb && x > 0) ? label[(size_t)y * W + x - 1] : 0xffffu;
        const uint32_t labU = (inb && y > 0 && use_up) ? label[(size_t)(y - 1) * W + x] : 0xffffu;
        // With a usable row above, a pixel that has the label of its upper neighbour has its value too, whatever
        // the frame: mode "up", like every non-event pixel -- also where its LEFT neighbour differs.  Such
        // pixels are no events (44 instead of 77 per row of a medium 512 x 512 map, at most 88 instead of 127).
        // A token that starts at such a pixel takes its key from the event in front of it; the key only matters
        // for a match too short for DEFLATE (one 2-byte pixel), so where a pixel is 2 bytes (`one_px_literal`)
        // the pixel is dropped only if its right neighbour continues the run (same label, same label above, no
        // forced cut): then no token that starts there is a single pixel.
        bool staticup = labU != 0xffffu && lab == labU;
        if (staticup && one_px_literal) {
            const int xn = x + 1;
            staticup = xn < xend && (xn % maxpix) != 0 && label[(size_t)y * W + xn] == lab && label[(size_t)(y - 1) * W + xn] == lab;
        }
        const bool isev = inb && ((x % maxpix) == 0 || labL == 0xffffu || (staticup ? false : (lab != labL || (labU != 0xffffu && lab != labU))));
        const unsigned b = __ballot_sync(0xffffffffu, isev);
        const int idx = count + __popc(b & ((1u << lane) - 1u));
        if (isev && idx < ENC_ECAP)
            out[idx] = make_uint2((uint32_t)x | ((2u * lab) << 16),
                                  (labL == 0xffffu ? none : 2u * labL) | ((labU == 0xffffu ? none : 2u * labU) << 16));
        count += __popc(b);
    }
    for (int j = min(count, ENC_ECAP) + lane; j < ENC_ECAP; j += 32) out[j] = make_uint2((uint32_t)xend | (none << 16), none | (none << 16));
    if (lane == 0) { atomicMax(max_count, count); row_count[item] = (uint16_t)min(count, ENC_ECAP); }
}

// Fills one geometry's block (zeroed by the caller).  One thread per pixel, atomics on ~n addresses;
// runs once per (geometry, output size).
__global__ void __launch_bounds__(256) enc_geometry_kernel(const uint16_t *label_img, const uint16_t *label_mask, int H,
                                                           int W, int n, const EncTables *tab, unsigned char *out)
{
    extern __shared__ uint32_t colpx[];     // [W] x^(8 * bytes after pixel x inside its row)
    for (int x = threadIdx.x; x < W; x += blockDim.x) colpx[x] = gf2_xpow8(tab->pow8, 2u * (uint32_t)(W - 1 - x));
    __syncthreads();
    uint32_t *Q = reinterpret_cast<uint32_t *>(out);
    uint32_t *cnt = Q + (n + 1);
    unsigned long long *M = reinterpret_cast<unsigned long long *>(out + (((size_t)(n + 1) * 8 + 7) & ~(size_t)7));
    const unsigned long long N = tab->img.raw_len;
    const unsigned long long row_bytes = 1ull + 4ull * (unsigned long long)W;
    const int npix = H * W;
    uint32_t *span = reinterpret_cast<uint32_t *>(out + geom_tab_sums_bytes(n) + 2 * geom_tab_events_bytes(H, W) + 16);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const int li = label_img[i], lm = label_mask[i];
        if (li > 0 && li <= n) {
            atomicAdd(&cnt[li], 1u);
            atomicAdd(&M[li], N - ((unsigned long long)y * row_bytes + 1ull + 4ull * (unsigned long long)x));
        }
        if (lm <= n) {
            atomicXor(&Q[lm], gf2_mulmod(tab->rowshift[y], colpx[x]));
            // first / last row of the label (a new extreme is rare: test before the atomic)
            if (span[2 * lm] < (uint32_t)(H - 1 - y)) atomicMax(&span[2 * lm], (uint32_t)(H - 1 - y));
            if (span[2 * lm + 1] < (uint32_t)y) atomicMax(&span[2 * lm + 1], (uint32_t)y);
        }
    }
}

// shared-memory tables of one CTA
constexpr int ENC_TCAP = 64;     // token descriptors per warp and batch (event parse)
struct EncSmem {
    uint32_t *len_s, *lit_s, *crc_s;
    uint32_t *toklist;           // [ENC_NW][ENC_TCAP]
    uint8_t *stage;
    uint32_t len_a, lit_a, tok_a;   // the same tables as 32-bit shared addresses (lds_u32)
};

// 16 labels of one lane -> 8 packed keys (two 16-bit keys per word)
template <bool IMG>
__device__ __forceinline__ void load_keys(const uint16_t *row, int xl, int nvalid, bool vec_ok, const FrameLut &lut,
                                          uint32_t k[8])
{
    const uint32_t kl_a = smem_addr(IMG ? lut.rank1 : lut.inst_lut);    // image: key LUT built over the rank scratch
    uint32_t w[8];
    if (nvalid == 16 && vec_ok) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row + xl));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(row + xl) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t l0 = (2 * q < nvalid) ? (uint32_t)__ldg(row + xl + 2 * q) : 0u;
            const uint32_t l1 = (2 * q + 1 < nvalid) ? (uint32_t)__ldg(row + xl + 2 * q + 1) : 0u;
            w[q] = l0 | (l1 << 16);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t l0 = w[q] & 0xffffu, l1 = w[q] >> 16;
        k[q] = lds_u16(kl_a + 2u * l0) | (lds_u16(kl_a + 2u * l1) << 16);
    }
}

template <bool IMG>
__device__ __forceinline__ uint32_t one_key(const uint16_t *p, const FrameLut &lut)
{
    const uint32_t l = __ldg(p);
    return lds_u16(smem_addr(IMG ? lut.rank1 : lut.inst_lut) + 2u * l);
}

// Encodes one stream into sm.stage.  Returns the total file size in bytes, or 0 when the staging
// buffer is too small (uniform over the CTA).
// Phase B of a pass: every lane knows the bit count of its tokens; a warp scan and a scan over the 16
// rows of the pass give the lane's absolute bit position.  false = the staging buffer would overflow
// (uniform over the CTA).
__device__ __forceinline__ bool place_rows(uint32_t lanebits, uint32_t *rowbits, uint32_t bitpos, uint32_t tail_room,
                                           uint32_t cap, uint32_t &pos, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = lanebits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) rowbits[warp] = incl;
    __syncthreads();
    const uint32_t mine = (lane < ENC_NW) ? rowbits[lane] : 0u;
    uint32_t sc = mine;
#pragma unroll
    for (int o = 1; o < ENC_NW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += t;
    }
    const uint32_t before = __shfl_sync(0xffffffffu, sc - mine, warp);
    total = __shfl_sync(0xffffffffu, sc, ENC_NW - 1);
    pos = bitpos + before + incl - lanebits;
    return (bitpos + total + 7u + 7u) / 8u + tail_room <= cap;
}

// The same for the event parse, whose tokens are interleaved over the lanes: only the row's total is needed
// (warp-uniform), and what comes back is the absolute bit position of the row's first token.
__device__ __forceinline__ bool place_row_totals(uint32_t rowtotal, uint32_t *rowbits, uint32_t bitpos, uint32_t tail_room,
                                                 uint32_t cap, uint32_t &rowpos, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) rowbits[warp] = rowtotal;
    __syncthreads();
    const uint32_t mine = (lane < ENC_NW) ? rowbits[lane] : 0u;
    uint32_t sc = mine;
#pragma unroll
    for (int o = 1; o < ENC_NW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += t;
    }
    rowpos = bitpos + __shfl_sync(0xffffffffu, sc - mine, warp);
    total = __shfl_sync(0xffffffffu, sc, ENC_NW - 1);
    return (bitpos + total + 7u + 7u) / 8u + tail_room <= cap;
}

// One token -> (bits, count).  A match of `len` pixels in mode up/left, or the literal pixel `key`
// (also for a one-pixel match of a 2-byte pixel, which is too short for DEFLATE).
template <bool IMG>
__device__ __forceinline__ void make_token(const EncSmem &sm, bool is_match, bool is_up, int len, uint32_t key,
                                           uint32_t dist_up, uint32_t dist_left, unsigned long long &val, int &nb)
{
    constexpr int BPP = IMG ? 4 : 2;
    if (is_match && len * BPP >= 3) {
        const uint32_t lt = lds_u32(sm.len_a + 4u * (uint32_t)(len * BPP));
        const uint32_t dd = is_up ? dist_up : dist_left;
        const int ln = (int)(lt >> 24);
        val = (unsigned long long)(lt & 0xffffffu) | ((unsigned long long)(dd & 0xffffffu) << ln);
        nb = ln + (int)(dd >> 24);
    } else if (IMG) {
        // 16-bit gray*4 and alpha*257, big endian: four literals
        const uint32_t g = key & 0xffu, al = (key >> 8) ? 255u : 0u;
        const uint32_t t0 = lds_u32(sm.lit_a + 4u * (g >> 6)), t1 = lds_u32(sm.lit_a + 4u * ((g << 2) & 0xffu)),
                       t2 = lds_u32(sm.lit_a + 4u * al);
        const int n0 = (int)(t0 >> 24), n1 = (int)(t1 >> 24), n2 = (int)(t2 >> 24);
        val = (unsigned long long)(t0 & 0xffffffu) | ((unsigned long long)(t1 & 0xffffffu) << n0) |
              ((unsigned long long)(t2 & 0xffffffu) << (n0 + n1)) | ((unsigned long long)(t2 & 0xffffffu) << (n0 + n1 + n2));
        nb = n0 + n1 + 2 * n2;
    } else {
        const uint32_t t0 = lds_u32(sm.lit_a + 4u * (key & 0xffu)), t1 = lds_u32(sm.lit_a + 4u * (key >> 8));
        const int n0 = (int)(t0 >> 24);
        val = (unsigned long long)(t0 & 0xffffffu) | ((unsigned long long)(t1 & 0xffffffu) << n0);
        nb = n0 + (int)(t1 >> 24);
    }
}

// K = 0: parse every pixel of a row segment (16 per lane).  K > 0: parse only the static events of the
// segment (K per lane); same tokens, a fraction of the look-ups.
template <bool IMG, int K>
__device__ uint32_t encode_stream(const EncodeArgs &a, const EncStream &d, const EncSmem &sm, const FrameLut &lut,
                                  const uint16_t *lmap, const uint2 *events, const GeomTab &gt, int n)
{
    __shared__ uint32_t rowbits[2][ENC_NW];
    __shared__ unsigned long long red64[2][ENC_NW];
    __shared__ uint32_t s_crc;
    __shared__ uint32_t rowact[ENC_MAXDIM / 32];     // mask stream: bit y = some label of row y has an instance id
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W;
    uint32_t *stage32 = reinterpret_cast<uint32_t *>(sm.stage);
    const uint32_t *lit_s = sm.lit_s, *crc_s = sm.crc_s;

    // ---- zero the staging buffer, place the container prefix ------------------------------------
    for (int i = tid; i < a.stage_cap / 16; i += ENC_NT)
        reinterpret_cast<uint4 *>(sm.stage)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s_crc = 0;
    if constexpr (!IMG && K > 0) {
        // label 0 with an id (the reference's "background instance", D5) makes every row active
        for (int i = tid; i < ENC_MAXDIM / 32; i += ENC_NT) rowact[i] = lut.inst_lut[0] ? 0xffffffffu : 0u;
    }
    __syncthreads();
    if constexpr (!IMG && K > 0) {
        for (int v = 1 + tid; v <= n; v += ENC_NT) {
            if (!lut.inst_lut[v]) continue;
            const int y0 = H - 1 - (int)gt.span[2 * v], y1 = (int)gt.span[2 * v + 1];      // rows the label touches
            for (int w = y0 >> 5; w <= (y1 >> 5); ++w) {
                const int lo = max(y0 - 32 * w, 0), hi = min(y1 - 32 * w, 31);
                atomicOr(&rowact[w], (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
            }
        }
    }
    for (int i = tid; i < d.prefix_len; i += ENC_NT) sm.stage[i] = d.prefix[i];
    // this stream's Huffman tables
    for (int i = tid; i < 260; i += ENC_NT) sm.len_s[i] = d.len[i];
    for (int i = tid; i < 256; i += ENC_NT) sm.lit_s[i] = d.lit[i];
    __syncthreads();
    uint32_t bitpos = 8u * (uint32_t)d.prefix_len;
    // block header: BFINAL = 1 and either BTYPE = 01 (fixed code) or BTYPE = 10 + the code lengths
    if (tid < (d.hdr_bits + 31) / 32) {
        const int nb = min(32, d.hdr_bits - 32 * tid);
        put_bits(stage32, bitpos + 32u * (uint32_t)tid, (unsigned long long)d.hdr[tid], nb);
    }
    bitpos += (uint32_t)d.hdr_bits;
    // uncompressed head bytes (.npy header) as literals
    if (tid < d.head_len) {
        const uint32_t t = lit_s[d.head[tid]];
        put_bits(stage32, bitpos + d.head_off[tid], (unsigned long long)(t & 0xffffffu), (int)(t >> 24));
    }
    bitpos += (uint32_t)d.head_bits;

    const int nseg = (W + 511) / 512;
    const int nitems = H * nseg;
    const bool vec_ok = (W % 8) == 0;
    constexpr int BPP = IMG ? 4 : 2;                 // == d.bpp
    constexpr int SEG_LANES = IMG ? 4 : 8;           // == d.seg_lanes: a forced cut every 256 bytes
    constexpr int MAXPIX = 256 / BPP;                // the same cut in pixels
    const uint32_t dist_up = d.dist_up, dist_left = d.dist_left;
    const uint32_t tail_room = (uint32_t)d.suffix_len + 24u;
    const bool use_up = (dist_up >> 24) != 0;
    const uint32_t filt = lit_s[0];                               // the PNG filter byte 0 as a literal
    bool overflow = false;
    int par = 0;

    for (int item0 = 0; item0 < nitems; item0 += ENC_NW) {           // par flips with every pass that has a barrier
        const int item = item0 + warp;
        const bool valid = item < nitems;
        const int y = valid ? (nseg == 1 ? item : item / nseg) : 0;
        const int x0 = (valid && nseg > 1) ? (item % nseg) * 512 : 0;
        const int xl = x0 + 16 * lane;
        const int nvalid = valid ? max(0, min(16, W - xl)) : 0;
        const int xend = min(W, x0 + 512);
        uint32_t pass_bits = 0;

        if constexpr (K == 0) {
            // ---- phase A: keys, modes, token starts ------------------------------------------------
            uint32_t kc[8];
            uint32_t mU = 0, mL = 0;
            const uint32_t V = (nvalid >= 16) ? 0xffffu : ((1u << nvalid) - 1u);
            if (nvalid > 0) load_keys<IMG>(lmap + (size_t)y * W, xl, nvalid, vec_ok, lut, kc);
            else {
    #pragma unroll
                for (int q = 0; q < 8; ++q) kc[q] = 0;
            }
            if (nvalid > 0 && y > 0 && use_up) {
                uint32_t ku[8];
                load_keys<IMG>(lmap + (size_t)(y - 1) * W, xl, nvalid, vec_ok, lut, ku);
    #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t x = kc[q] ^ ku[q];
                    mU |= ((x & 0xffffu) == 0u ? 1u : 0u) << (2 * q);
                    mU |= ((x >> 16) == 0u ? 1u : 0u) << (2 * q + 1);
                }
                mU &= V;
            }
            {
                // previous pixel of this lane's first pixel: the previous lane's last key, or (first lane of
                // a later segment) the pixel in front of the segment; none at the start of a row
                uint32_t prev = __shfl_up_sync(0xffffffffu, kc[7] >> 16, 1);
                bool has_prev = true;
                if (lane == 0) {
                    has_prev = valid && x0 > 0;
                    prev = has_prev ? one_key<IMG>(lmap + (size_t)y * W + x0 - 1, lut) : 0u;
                }
                uint32_t pk = prev;
    #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t lo = kc[q] & 0xffffu, hi = kc[q] >> 16;
                    mL |= (lo == pk ? 1u : 0u) << (2 * q);
                    mL |= (hi == lo ? 1u : 0u) << (2 * q + 1);
                    pk = hi;
                }
                if (!has_prev) mL &= ~1u;
                mL &= V & ~mU;
            }
            // token starts: literal pixels, mode changes, forced cuts every 256 bytes / at the segment start
            uint32_t S;
            {
                const uint32_t cu = __shfl_up_sync(0xffffffffu, mU >> 15, 1) & 1u;
                const uint32_t cl = __shfl_up_sync(0xffffffffu, mL >> 15, 1) & 1u;
                const uint32_t pU = (mU << 1) | (lane ? cu : 0u);
                const uint32_t pL = (mL << 1) | (lane ? cl : 0u);
                S = (V & ~mU & ~mL) | (mU & ~pU) | (mL & ~pL);
                if ((lane & (SEG_LANES - 1)) == 0) S |= 1u;
                S &= V;
            }
            // end of a token that runs past this lane: the first start in a later lane, or the segment end
            int nxt;
            {
                int fs = S ? (xl + __ffs((int)S) - 1) : 0x7fffffff;
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_down_sync(0xffffffffu, fs, o);
                    if (lane + o < 32) fs = min(fs, t);
                }
                nxt = __shfl_down_sync(0xffffffffu, fs, 1);
                if (lane == 31) nxt = 0x7fffffff;
                nxt = min(nxt, xend);
            }

            // token walk: WRITE = false sums the bit counts, WRITE = true ORs the bits in at `pos`
            auto walk = [&](bool write, uint32_t pos) -> uint32_t {
                uint32_t bits = 0;
                if (IMG && lane == 0 && x0 == 0 && valid) {
                    if (write) put_bits(stage32, pos, (unsigned long long)(filt & 0xffffffu), (int)(filt >> 24));
                    bits += filt >> 24;
                }
                uint32_t rest = S;
                while (rest) {
                    const int j = __ffs((int)rest) - 1;
                    rest &= rest - 1;
                    const int len = (rest ? (__ffs((int)rest) - 1) : (nxt - xl)) - j;
                    const bool isU = (mU >> j) & 1u, isL = (mL >> j) & 1u;
                    const int q = j >> 1;
                    uint32_t w = kc[0];
                    w = (q == 1) ? kc[1] : w; w = (q == 2) ? kc[2] : w; w = (q == 3) ? kc[3] : w;
                    w = (q == 4) ? kc[4] : w; w = (q == 5) ? kc[5] : w; w = (q == 6) ? kc[6] : w;
                    w = (q == 7) ? kc[7] : w;
                    const uint32_t key = (j & 1) ? (w >> 16) : (w & 0xffffu);
                    unsigned long long val;
                    int nb;
                    make_token<IMG>(sm, isU || isL, isU, len, key, dist_up, dist_left, val, nb);
                    if (write) put_bits(stage32, pos + bits, val, nb);
                    bits += (uint32_t)nb;
                }
                return bits;
            };
            const uint32_t lanebits = walk(false, 0u);
            uint32_t pos;
            if (!place_rows(lanebits, rowbits[par], bitpos, tail_room, (uint32_t)a.stage_cap, pos, pass_bits)) { overflow = true; break; }
            if (valid) walk(true, pos);
        } else {
            // ---- phase A (events): K static events per lane; values only decide each event's own mode ------
            const uint32_t gm = (y > 0 && use_up) ? 2u : 1u;            // mode of every non-event pixel: up, else left
            // Mask stream: a row none of whose labels has an instance id in this frame, under a row of the same
            // kind, is all zeros over zeros -- every pixel in mode "up", tokens only at the forced cuts.  Its
            // descriptors are known without looking at a single event (most rows of most frames: the active set
            // of a pouch is small).
            bool fast = false;
            if constexpr (!IMG) {
                // a whole pass of such rows (CTA-uniform: rows item0 - 1 .. item0 + 15 without an id) needs neither
                // the per-row bit counts nor the barrier that exchanges them: every row is the same R bits
                if (nseg == 1 && use_up && item0 >= 1 && item0 + ENC_NW <= nitems) {
                    const int yb = item0 - 1;
                    const unsigned long long two = (unsigned long long)rowact[yb >> 5] |
                                                   ((unsigned long long)rowact[min((yb >> 5) + 1, ENC_MAXDIM / 32 - 1)] << 32);
                    if (((two >> (yb & 31)) & ((1ull << (ENC_NW + 1)) - 1ull)) == 0ull) {
                        const int ncut = (xend + MAXPIX - 1) / MAXPIX;          // <= 4 tokens per row (W <= 512)
                        unsigned long long val = 0; int nb = 0;
                        if (lane < ncut) make_token<false>(sm, true, true, min(MAXPIX, xend - lane * MAXPIX), 0u, dist_up, dist_left, val, nb);
                        uint32_t inc = (uint32_t)nb;
#pragma unroll
                        for (int o = 1; o < 4; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += u;
                        }
                        const uint32_t R = __shfl_sync(0xffffffffu, inc, 3);
                        if ((bitpos + (uint32_t)ENC_NW * R + 7u + 7u) / 8u + tail_room > (uint32_t)a.stage_cap) { overflow = true; break; }
                        if (lane < ncut) put_bits(stage32, bitpos + (uint32_t)warp * R + inc - (uint32_t)nb, val, nb);
                        bitpos += (uint32_t)ENC_NW * R;
                        continue;
                    }
                }
                if (valid && nseg == 1 && use_up && y > 0)
                    fast = (((rowact[y >> 5] >> (y & 31)) | (rowact[(y - 1) >> 5] >> ((y - 1) & 31))) & 1u) == 0u;
            }
            const uint32_t list_a = sm.tok_a + (uint32_t)(warp * ENC_TCAP) * 4u;
            const bool lead = IMG && lane == 0 && x0 == 0 && valid;           // the PNG filter byte of the row
            uint32_t S = 0;
            uint32_t desc[2 * K];
            int ntok, rank0 = 0;
            if (fast) {
                const int ncut = (xend + MAXPIX - 1) / MAXPIX;                  // <= 512 / 128
                if (lane < ncut) sts_u32(list_a + 4u * (uint32_t)lane, (uint32_t)min(MAXPIX, xend - lane * MAXPIX) | (2u << 12));
                __syncwarp();
                ntok = ncut;
            } else {
                // events per lane in THIS row: its static event count spread over the 32 lanes (a medium row has
                // 44 events on average, 88 at most: mostly 2 per lane, K only where the map is busiest); every
                // per-event block below is skipped, warp-uniformly, beyond kr
                const int kr = valid ? min(K, max(1, ((int)__ldg((IMG ? gt.cnt_img : gt.cnt_msk) + item) + 31) >> 5)) : 1;
                uint32_t info[K];                                           // x | mode << 12 | key << 16
                uint32_t ex[K];
                {
                    const uint2 *ev = events + (size_t)item * ENC_ECAP + lane * kr;
                    const uint32_t kl_a = smem_addr(IMG ? lut.rank1 : lut.inst_lut);
                    const uint32_t none = 2u * (uint32_t)(n + 1);
                    uint32_t rx[K], ry[K];
#pragma unroll
                    for (int e = 0; e < K; ++e) {
                        rx[e] = (uint32_t)xend | (none << 16); ry[e] = none | (none << 16);
                        if (valid && e < kr) {
                            const uint2 r = __ldg(ev + e);
                            rx[e] = r.x; ry[e] = r.y;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < K; ++e) {
                        ex[e] = (uint32_t)xend; info[e] = (uint32_t)xend;
                        if (e < kr) {
                            // a missing neighbour (and a padding slot) reads the LUT's extra entry n + 1, a key no
                            // pixel has: three unconditional look-ups, two plain comparisons
                            const uint32_t x = rx[e] & 0xffffu;
                            const uint32_t k = lds_u16(kl_a + (rx[e] >> 16));
                            const uint32_t kL = lds_u16(kl_a + (ry[e] & 0xffffu));
                            const uint32_t kU = lds_u16(kl_a + (ry[e] >> 16));
                            const uint32_t m = (k == kU) ? 2u : ((k == kL) ? 1u : 0u);
                            ex[e] = x;
                            info[e] = x | (m << 12) | (k << 16);
                        }
                    }
                }
                // token starts: slot 2e = at the event pixel (forced cut, literal, or mode differs from the pixel
                // before), slot 2e+1 = at the pixel after it (the gap's static mode differs from the event's)
                {
                    uint32_t last = info[0];                               // the lane's last event: info[kr - 1]
#pragma unroll
                    for (int e = 1; e < K; ++e) last = (e < kr) ? info[e] : last;
                    const uint32_t pinfo = __shfl_up_sync(0xffffffffu, last, 1);
                    uint32_t nx0 = __shfl_down_sync(0xffffffffu, ex[0], 1);
                    if (lane == 31) nx0 = (uint32_t)xend;
                    uint32_t px = pinfo & 0xfffu, pm = (pinfo >> 12) & 3u;
#pragma unroll
                    for (int e = 0; e < K; ++e) {
                        if (e < kr) {
                            const uint32_t x = ex[e], m = (info[e] >> 12) & 3u;
                            const bool isev = (int)x < xend;
                            const uint32_t before = (x - px > 1u) ? gm : pm;     // mode of pixel x-1
                            const bool forced = ((x & (uint32_t)(MAXPIX - 1)) == 0u);
                            const uint32_t nx = (e + 1 < K && e + 1 < kr) ? ex[(e + 1 < K) ? e + 1 : e] : nx0;
                            if (isev && (forced || m == 0u || m != before)) S |= 1u << (2 * e);
                            if (isev && nx - x > 1u && m != gm) S |= 2u << (2 * e);
                            px = x; pm = m;
                        }
                    }
                }
                // ---- tokens, compacted.  A lane's events start 0..2K tokens (most: none), so walking them
                // per lane would cost every lane the longest walk of the warp.  Instead each lane drops its
                // token descriptors (pixels | mode << 12 | key << 16) at their rank into a per-warp list and
                // the warp then handles them 32 at a time, one token per lane.
                {
                    // first token start at or after the next lane's first event: the end of this lane's last token
                    int fs = 0x7fffffff;
    #pragma unroll
                    for (int sidx = 2 * K - 1; sidx >= 0; --sidx)
                        if ((S >> sidx) & 1u) fs = (int)ex[sidx >> 1] + (sidx & 1);
    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_down_sync(0xffffffffu, fs, o);
                        if (lane + o < 32) fs = min(fs, t);
                    }
                    int nextpos = __shfl_down_sync(0xffffffffu, fs, 1);
                    if (lane == 31) nextpos = 0x7fffffff;
                    nextpos = min(nextpos, xend);
    #pragma unroll
                    for (int sidx = 2 * K - 1; sidx >= 0; --sidx) {
                        if ((sidx >> 1) < kr) {
                            const int pos = (int)ex[sidx >> 1] + (sidx & 1);
                            const uint32_t m = (sidx & 1) ? gm : ((info[sidx >> 1] >> 12) & 3u);
                            desc[sidx] = ((uint32_t)(nextpos - pos) & 0xfffu) | (m << 12) | (info[sidx >> 1] & 0xffff0000u);
                            if ((S >> sidx) & 1u) nextpos = pos;
                        }
                    }
                }
                {
                    const int mine = __popc(S) + (lead ? 1 : 0);
                    int incl = mine;
    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    ntok = __shfl_sync(0xffffffffu, incl, 31);
                    rank0 = incl - mine;
                }
            }
            // descriptors with rank in [b0, b0 + ENC_TCAP) -> list (rows with more tokens go batch by batch);
            // the rank of slot s is rank0 + (lead) + the number of starts below it
            auto fill = [&](int b0) {
                const int base = rank0 - b0 + (lead ? 1 : 0);
                if (lead && (unsigned)(rank0 - b0) < (unsigned)ENC_TCAP) sts_u32(list_a + 4u * (uint32_t)(rank0 - b0), 0xffffffffu);
#pragma unroll
                for (int sidx = 0; sidx < 2 * K; ++sidx) {
                    if ((S >> sidx) == 0u) break;                  // warp-divergent at worst: nothing above this slot
                    const int at = base + __popc(S & ((1u << sidx) - 1u));
                    if (((S >> sidx) & 1u) && (unsigned)at < (unsigned)ENC_TCAP) sts_u32(list_a + 4u * (uint32_t)at, desc[sidx]);
                }
                __syncwarp();
            };
            auto token_at = [&](int t, unsigned long long &val, int &nb) {
                const uint32_t dsc = lds_u32(list_a + 4u * (uint32_t)t);
                if (dsc == 0xffffffffu) { val = filt & 0xffffffu; nb = (int)(filt >> 24); return; }
                const uint32_t m = (dsc >> 12) & 3u;
                make_token<IMG>(sm, m != 0u, m == 2u, (int)(dsc & 0xfffu), dsc >> 16, dist_up, dist_left, val, nb);
            };
            // A row with at most ENC_TCAP tokens (all but pathological ones) builds every token ONCE: bits, bit
            // count and the position relative to the row start (a warp scan per round of 32 tokens) stay in
            // registers across the pass-wide prefix sum, and the write phase is one put_bits per token.
            const bool onebatch = ntok <= ENC_TCAP;                            // warp-uniform
            unsigned long long cval[ENC_TCAP / 32];
            uint32_t cmeta[ENC_TCAP / 32];                                     // relative bit position | bit count << 24
            uint32_t rowtotal = 0;
            if (onebatch) {
                if (!fast) fill(0);
#pragma unroll
                for (int r = 0; r < ENC_TCAP / 32; ++r) {
                    cval[r] = 0; cmeta[r] = 0;
                    if (r * 32 < ntok) {                                       // warp-uniform
                        const int t = r * 32 + lane;
                        unsigned long long val = 0; int nb = 0;
                        if (t < ntok) token_at(t, val, nb);
                        uint32_t inc = (uint32_t)nb;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += u;
                        }
                        cval[r] = val;
                        cmeta[r] = (rowtotal + inc - (uint32_t)nb) | ((uint32_t)nb << 24);
                        rowtotal += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
            } else {
                // bit counts, one token per lane and round
                uint32_t lanebits = 0;
                for (int b0 = 0; b0 < ntok; b0 += ENC_TCAP) {
                    if (b0) __syncwarp();
                    fill(b0);
                    const int nt = min(ntok - b0, ENC_TCAP);
                    for (int t = lane; t < nt; t += 32) {
                        unsigned long long val; int nb;
                        token_at(t, val, nb);
                        lanebits += (uint32_t)nb;
                    }
                }
                rowtotal = __reduce_add_sync(0xffffffffu, lanebits);
            }
            uint32_t rowpos;
            if (!place_row_totals(rowtotal, rowbits[par], bitpos, tail_room, (uint32_t)a.stage_cap, rowpos, pass_bits)) { overflow = true; break; }
            if (onebatch) {
#pragma unroll
                for (int r = 0; r < ENC_TCAP / 32; ++r) {
                    const int nb = (int)(cmeta[r] >> 24);
                    if (nb && valid) put_bits(stage32, rowpos + (cmeta[r] & 0xffffffu), cval[r], nb);
                }
            } else {
                // tokens are interleaved over the lanes by round, not contiguous per lane: scan round by round
                for (int b0 = 0; b0 < ntok; b0 += ENC_TCAP) {
                    __syncwarp(); fill(b0);
                    const int nt = min(ntok - b0, ENC_TCAP);
                    for (int r0 = 0; r0 < nt; r0 += 32) {
                        const int t = r0 + lane;
                        unsigned long long val = 0; int nb = 0;
                        if (t < nt) token_at(t, val, nb);
                        uint32_t inc = (uint32_t)nb;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += u;
                        }
                        if (t < nt && valid) put_bits(stage32, rowpos + inc - (uint32_t)nb, val, nb);
                        rowpos += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
            }
        }
        bitpos += pass_bits;
        par ^= 1;
    }
    __syncthreads();
    if (overflow) return 0u;

    // end-of-block symbol, then pad to a byte
    if (tid == 0) put_bits(stage32, bitpos, (unsigned long long)(d.eob & 0xffffffu), (int)(d.eob >> 24));
    bitpos += d.eob >> 24;
    __syncthreads();
    const uint32_t plen = (uint32_t)d.prefix_len;
    const uint32_t L = (bitpos - 8u * plen + 7u) / 8u;          // deflate bytes
    uint32_t total_bytes;
    if (IMG) {
        // Adler-32 of the scanlines (big endian, ends the zlib stream) without touching a pixel:
        // a = 1 + sum_i d_i and b = N + sum_i (N - i) d_i are sums over pixels of terms that depend only
        // on the pixel's label (through its gray level) and position, so with the static per-label pixel
        // count cnt_l and position weight M_l = sum_{px in l} (N - pos_px):
        //   sum d = sum_l cnt_l * S(g_l),   sum (N - i) d_i = sum_l (M_l * S(g_l) - cnt_l * w(g_l))
        // S = byte sum of a pixel, w = sum_k k * byte_k.  Background pixels are all-zero bytes.
        unsigned long long ad1 = 0, ad2 = 0;
        for (int l = 1 + tid; l <= n; l += ENC_NT) {
            const uint32_t g = lut.gray_lut[l];
            const uint32_t b1 = (g << 2) & 0xffu;
            const unsigned long long S = (g >> 6) + b1 + 510u, w = b1 + 1275u;
            const unsigned long long cnt = gt.cnt[l];
            ad1 += cnt * S;
            ad2 += gt.M[l] * S - cnt * w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ad1 += __shfl_xor_sync(0xffffffffu, ad1, o);
            ad2 += __shfl_xor_sync(0xffffffffu, ad2, o);
        }
        if (lane == 0) { red64[0][warp] = ad1; red64[1][warp] = ad2; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long s1 = 1, s2 = d.raw_len;
            for (int w = 0; w < ENC_NW; ++w) { s1 += red64[0][w]; s2 += red64[1][w]; }
            const uint32_t adler = (uint32_t)(s1 % 65521ull) | ((uint32_t)(s2 % 65521ull) << 16);
            put_be32(sm.stage + plen + L, adler);
            put_be32(sm.stage + d.p_len_prefix, L + 6u);          // zlib header 2 + deflate + Adler 4
        }
        __syncthreads();
        // CRC-32 of the IDAT chunk (type + data).  crc0 (CRC from a zero register, no final xor) is linear and
        // crc0(A | B) = crc0(A) * x^(8|B|) ^ crc0(B), so the range is cut into 64-byte pieces counted from its
        // END (the first one may be shorter), piece j contributes crc0(piece) * x^(8 * 64 * j) with a power
        // from a static table, and the contributions are XOR-ed: one polynomial product per piece, none of
        // them dependent on another.
        const uint32_t r0 = (uint32_t)d.p_len_prefix + 4u, r1 = plen + L + 4u;
        const uint32_t Lr = r1 - r0;
        {
            uint32_t acc = 0;
            for (uint32_t j = (uint32_t)tid; j * 64u < Lr; j += ENC_NT) {
                const uint32_t e = r1 - 64u * j;
                const uint32_t b = (e - r0 >= 64u) ? e - 64u : r0;
                uint32_t c = 0;
                for (uint32_t i = b; i < e; ++i) c = crc_s[(c ^ sm.stage[i]) & 0xffu] ^ (c >> 8);
                acc ^= gf2_mulmod(c, a.tab->pow64[j]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0 && acc) atomicXor(&s_crc, acc);
        }
        __syncthreads();
        if (tid == 0) {
            // CRC(M) = crc0(M) ^ CRC(zeros(|M|));  CRC(zeros(n)) = (0xFFFFFFFF * x^(8n)) ^ 0xFFFFFFFF
            const uint32_t xl = gf2_mulmod(a.tab->pow64[Lr >> 6], a.tab->pow1[Lr & 63u]);
            const uint32_t aff = gf2_mulmod(0xffffffffu, xl) ^ 0xffffffffu;
            put_be32(sm.stage + r1, s_crc ^ aff);
            for (int i = 0; i < d.suffix_len; ++i) sm.stage[r1 + 4 + i] = d.suffix[i];
        }
        total_bytes = r1 + 4u + (uint32_t)d.suffix_len;
    } else {
        // CRC-32 of the .npy bytes without touching a pixel: CRC is linear over GF(2), so the pixels
        // of one label value v contribute crc0(bytes of id_v) * Q_v with the static polynomial
        // Q_v = sum_{px with label v} x^(8 * bytes after px) mod P
        uint32_t crc_acc = 0;
        for (int v = tid; v <= n; v += ENC_NT) {
            const uint32_t id = lut.inst_lut[v];
            if (id) {
                uint32_t c = crc_s[id & 0xffu];
                c = crc_s[(c ^ (id >> 8)) & 0xffu] ^ (c >> 8);
                crc_acc ^= gf2_mulmod(c, gt.Q[v]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) crc_acc ^= __shfl_xor_sync(0xffffffffu, crc_acc, o);
        if (lane == 0 && crc_acc) atomicXor(&s_crc, crc_acc);
        __syncthreads();
        if (tid == 0) {
            const uint32_t crc = s_crc ^ d.crc_const;
            uint8_t *suf = sm.stage + plen + L;
            for (int i = 0; i < d.suffix_len; ++i) suf[i] = d.suffix[i];
            put_le32(sm.stage + d.p_crc_prefix, crc);
            put_le32(sm.stage + d.p_csize_prefix, L);
            put_le32(suf + d.p_crc_suffix, crc);
            put_le32(suf + d.p_csize_suffix, L);
            put_le32(suf + d.p_cdoff_suffix, plen + L);
        }
        total_bytes = plen + L + (uint32_t)d.suffix_len;
    }
    __syncthreads();
    return total_bytes;
}

// Copies a finished file from the staging buffer into the arena and fills its record.
__device__ void emit_file(const EncodeArgs &a, const uint8_t *stage, uint32_t total, EncRecord *rec)
{
    __shared__ unsigned long long s_off;
    const int tid = threadIdx.x;
    const uint32_t total16 = (total + 15u) & ~15u;
    if (tid == 0) {
        unsigned long long off = ~0ull;
        uint32_t flags = CAB200_ENC_STAGE_FULL;
        if (total) {
            off = atomicAdd(a.cursor, (unsigned long long)total16);
            flags = CAB200_ENC_OK;
            if (off + total16 > a.arena_bytes) { off = ~0ull; flags = CAB200_ENC_ARENA_FULL; }
        }
        s_off = off;
        rec->offset = (off == ~0ull) ? 0ull : off;
        rec->size = (off == ~0ull) ? 0u : total;
        rec->flags = flags;
    }
    __syncthreads();
    const unsigned long long off = s_off;
    if (off != ~0ull) {
        uint4 *dst = reinterpret_cast<uint4 *>(a.arena + off);
        const uint4 *src = reinterpret_cast<const uint4 *>(stage);
        for (uint32_t i = tid; i < total16 / 16; i += ENC_NT) dst[i] = src[i];
    }
    __syncthreads();
}

template <int K>
__global__ void __launch_bounds__(ENC_NT, 2) encode_kernel(const EncodeArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int pf = blockIdx.x;
    const int f = pf % a.n_frames;
    const int pi = pf / a.n_frames;
    const PouchEntry pe = a.pouches[pi];
    const int n = pe.n;
    const double *ca = a.ca_frames + (size_t)pe.cell_off * 4 * a.n_frames + (size_t)f * 4 * n;

    const FrameLut lut = frame_lut_carve(smem_raw, n);
    EncSmem sm;
    unsigned char *p = smem_raw + a.lut_bytes;
    sm.len_s = reinterpret_cast<uint32_t *>(p); p += 260 * 4;
    sm.lit_s = reinterpret_cast<uint32_t *>(p); p += 256 * 4;
    sm.crc_s = reinterpret_cast<uint32_t *>(p); p += 256 * 4;
    sm.toklist = reinterpret_cast<uint32_t *>(p); p += (K > 0 ? ENC_NW * ENC_TCAP * 4 : 0);
    sm.stage = p;                                   // 16-byte aligned: lut_bytes is, 260*4 = 65*16
    sm.len_a = smem_addr(sm.len_s); sm.lit_a = smem_addr(sm.lit_s); sm.tok_a = smem_addr(sm.toklist);

    const int n_active = frame_lut_build<ENC_NT>(lut, ca, n, a.threshold, a.quirk);
    if (tid == 0 && a.n_active_out) a.n_active_out[pf] = n_active;
    for (int i = tid; i < 256; i += ENC_NT) sm.crc_s[i] = a.tab->crc[i];
    // image pixel key per label value: gray | 0x100 (alpha set), 0 for the background; the rank
    // scratch of the LUT build is free now
    for (int v = tid; v <= n; v += ENC_NT) lut.rank1[v] = v ? (uint16_t)(lut.gray_lut[v] | 0x100u) : (uint16_t)0;
    // entry n + 1 of both key LUTs: "no such neighbour" in the static event lists (image keys are < 0x200,
    // instance ids <= n <= 32766)
    if (tid == 0) { lut.rank1[n + 1] = 0xffffu; lut.inst_lut[n + 1] = 0xffffu; }
    __syncthreads();
    EncRecord *rec = a.records + (size_t)pf * 2;
    const size_t gbase = (size_t)pe.geom * a.H * a.W;
    const GeomTab gt = geom_tab_carve(a.geom_tabs + a.geom_tab_off[pe.geom], n, a.H, a.W);
    if (a.want_img) {
        const uint32_t total = encode_stream<true, K>(a, a.tab->img, sm, lut, a.label_img + gbase, gt.ev_img, gt, n);
        emit_file(a, sm.stage, total, rec);
    } else if (tid == 0) {
        rec[0].offset = 0; rec[0].size = 0; rec[0].flags = CAB200_ENC_SKIPPED;
    }
    if (a.want_msk && n_active > 0) {           // the reference writes no mask file for a frame without active cells
        const uint32_t total = encode_stream<false, K>(a, a.tab->msk, sm, lut, a.label_mask + gbase, gt.ev_msk, gt, n);
        emit_file(a, sm.stage, total, rec + 1);
    } else if (tid == 0) {
        rec[1].offset = 0; rec[1].size = 0; rec[1].flags = CAB200_ENC_SKIPPED;
    }
}

// Device -> pinned host memory with plain stores (zero copy) instead of the DMA engine: the few
// bytes the host must see before it can size a chunk's copy-out must not queue behind megabytes of
// earlier copies on the same engine.
__global__ void __launch_bounds__(256) publish_kernel(uint4 *dst, const uint4 *src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---- host side: tables and container templates -----------------------------------------------------
static uint32_t rev_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// fixed Huffman code lengths of the literal/length alphabet (RFC 1951, 3.2.6)
static int fixed_len(int sym) { return sym <= 143 ? 8 : sym <= 255 ? 9 : sym <= 279 ? 7 : 8; }

// Huffman code lengths for weights w[0..n) (all > 0), at most maxlen bits.  Deterministic so that the
// oracle can restate it: priority queue ordered by (weight, id), leaves have ids 0..n-1, internal nodes get
// the next ids in creation order; if the deepest leaf exceeds maxlen all weights are halved (rounding up)
// and the construction is repeated.
static void huffman_lengths(const uint32_t *w, int n, int maxlen, int *len_out)
{
    std::vector<unsigned long long> f(w, w + n);
    for (;;) {
        struct Node { unsigned long long w; int id; };
        auto cmp = [](const Node &a, const Node &b) { return a.w != b.w ? a.w > b.w : a.id > b.id; };
        std::vector<Node> heap;
        std::vector<int> parent(2 * n, -1);
        for (int i = 0; i < n; ++i) heap.push_back({f[i], i});
        std::make_heap(heap.begin(), heap.end(), cmp);
        int next = n;
        while (heap.size() > 1) {
            std::pop_heap(heap.begin(), heap.end(), cmp); Node a = heap.back(); heap.pop_back();
            std::pop_heap(heap.begin(), heap.end(), cmp); Node b = heap.back(); heap.pop_back();
            parent[a.id] = next; parent[b.id] = next;
            heap.push_back({a.w + b.w, next++});
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
        int deepest = 0;
        for (int i = 0; i < n; ++i) {
            int d = 0;
            for (int p = parent[i]; p >= 0; p = parent[p]) ++d;
            len_out[i] = (n == 1) ? 1 : d;
            deepest = std::max(deepest, len_out[i]);
        }
        if (deepest <= maxlen) return;
        for (int i = 0; i < n; ++i) f[i] = (f[i] + 1) / 2;
    }
}

// canonical codes (RFC 1951, 3.2.2) for lengths len[0..n) (0 = unused); code_out holds the code MSB first
static void canonical_codes(const int *len, int n, uint32_t *code_out)
{
    int bl_count[16] = {0};
    for (int i = 0; i < n; ++i) bl_count[len[i]]++;
    bl_count[0] = 0;
    uint32_t next_code[16] = {0}, code = 0;
    for (int b = 1; b <= 15; ++b) { code = (code + bl_count[b - 1]) << 1; next_code[b] = code; }
    for (int i = 0; i < n; ++i) code_out[i] = len[i] ? next_code[len[i]]++ : 0u;
}

struct BitWriter {
    uint32_t *w; int cap_bits; int n = 0;
    void put(uint32_t v, int nb) { for (int i = 0; i < nb; ++i, ++n) if (n < cap_bits && ((v >> i) & 1u)) w[n >> 5] |= 1u << (n & 31); }
};

static const int kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99,
                                 115, 131, 163, 195, 227, 258};
static const int kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const int kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025,
                                  1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const int kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12,
                                   12, 13, 13};
static int dist_symbol(int dist) { int i = 29; while (dist < kDistBase[i]) --i; return i; }

#include "enc_freq.inc"

// Fills a stream's code tables, block header and distance entries.  tuned: a dynamic-Huffman block whose
// code lengths come from the measured symbol frequencies of enc_freq.inc (static per stream kind, so the
// kernel stays table driven); else the fixed code of RFC 1951 3.2.6.
static void build_codes(EncStream &d, int dist_left, int dist_up, const uint32_t *freq_ll, const uint32_t *freq_d2, bool tuned)
{
    int ll_len[286], d_len[30];
    uint32_t ll_code[286], d_code[30];
    if (tuned) {
        uint32_t fd[30];
        for (int i = 0; i < 30; ++i) fd[i] = 1;
        fd[dist_symbol(dist_left)] = freq_d2[0];
        if (dist_up >= 1 && dist_up <= 32768) fd[dist_symbol(dist_up)] = freq_d2[1];
        huffman_lengths(freq_ll, 286, 15, ll_len);
        huffman_lengths(fd, 30, 15, d_len);
    } else {
        for (int i = 0; i < 286; ++i) ll_len[i] = fixed_len(i);
        for (int i = 0; i < 30; ++i) d_len[i] = 5;
    }
    if (tuned) {
        canonical_codes(ll_len, 286, ll_code);
        canonical_codes(d_len, 30, d_code);
    } else {
        int full[288];
        uint32_t fc[288];
        for (int i = 0; i < 288; ++i) full[i] = fixed_len(i);
        canonical_codes(full, 288, fc);                  // the fixed code is canonical over all 288 symbols
        for (int i = 0; i < 286; ++i) ll_code[i] = fc[i];
        for (int i = 0; i < 30; ++i) d_code[i] = (uint32_t)i;
    }
    for (int i = 0; i < 256; ++i) d.lit[i] = rev_bits(ll_code[i], ll_len[i]) | ((uint32_t)ll_len[i] << 24);
    d.eob = rev_bits(ll_code[256], ll_len[256]) | ((uint32_t)ll_len[256] << 24);
    for (int l = 3; l <= 258; ++l) {
        int i = 28;
        while (l < kLenBase[i]) --i;
        const int nb = ll_len[257 + i];
        d.len[l] = rev_bits(ll_code[257 + i], nb) | ((uint32_t)(l - kLenBase[i]) << nb) | ((uint32_t)(nb + kLenExtra[i]) << 24);
    }
    auto dist_entry = [&](int dist) -> uint32_t {
        if (dist < 1 || dist > 32768) return 0u;
        const int i = dist_symbol(dist), nb = d_len[i];
        if (nb + kDistExtra[i] > 24) return 0u;          // does not fit the entry: that distance is not used
        return rev_bits(d_code[i], nb) | ((uint32_t)(dist - kDistBase[i]) << nb) | ((uint32_t)(nb + kDistExtra[i]) << 24);
    };
    d.dist_left = dist_entry(dist_left);
    d.dist_up = dist_entry(dist_up);
    // block header
    memset(d.hdr, 0, sizeof(d.hdr));
    BitWriter bw{d.hdr, (int)sizeof(d.hdr) * 8};
    bw.put(1, 1);                                        // BFINAL
    if (!tuned) {
        bw.put(1, 2);                                    // BTYPE = 01
    } else {
        bw.put(2, 2);                                    // BTYPE = 10: HLIT, HDIST, HCLEN, code-length code, lengths
        bw.put(286 - 257, 5); bw.put(30 - 1, 5); bw.put(19 - 4, 4);
        // code-length alphabet: the 316 lengths are sent one by one (no repeat codes 16-18)
        uint32_t clf[19] = {0};
        for (int i = 0; i < 286; ++i) clf[ll_len[i]]++;
        for (int i = 0; i < 30; ++i) clf[d_len[i]]++;
        int used[19], nused = 0;
        for (int v = 0; v < 19; ++v) if (clf[v]) used[nused++] = v;
        if (nused == 1) { used[nused++] = (used[0] == 0) ? 1 : 0; clf[used[1]] = 1; std::sort(used, used + nused); }
        uint32_t uw[19]; int ul[19];
        for (int k = 0; k < nused; ++k) uw[k] = clf[used[k]];
        huffman_lengths(uw, nused, 7, ul);
        int cl_len[19] = {0};
        uint32_t cl_code[19];
        for (int k = 0; k < nused; ++k) cl_len[used[k]] = ul[k];
        canonical_codes(cl_len, 19, cl_code);
        static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int k = 0; k < 19; ++k) bw.put((uint32_t)cl_len[order[k]], 3);
        for (int i = 0; i < 286; ++i) bw.put(rev_bits(cl_code[ll_len[i]], cl_len[ll_len[i]]), cl_len[ll_len[i]]);
        for (int i = 0; i < 30; ++i) bw.put(rev_bits(cl_code[d_len[i]], cl_len[d_len[i]]), cl_len[d_len[i]]);
    }
    d.hdr_bits = bw.n;
}

static uint32_t host_crc0(const uint32_t *tab, const uint8_t *p, size_t n, uint32_t c = 0)
{
    for (size_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c;
}
static uint32_t host_crc(const uint32_t *tab, const uint8_t *p, size_t n)
{
    return host_crc0(tab, p, n, 0xffffffffu) ^ 0xffffffffu;
}
static uint32_t host_xpow8(const uint32_t *pow8, unsigned long long n)
{
    uint32_t r = 0x80000000u;
    for (int k = 0; n; ++k, n >>= 1)
        if (n & 1ull) r = gf2_mulmod(r, pow8[k]);
    return r;
}
static void h_le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void h_le32(uint8_t *p, uint32_t v) { h_le16(p, v); h_le16(p + 2, v >> 16); }
static void h_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

static void build_tables(int H, int W, bool tuned, EncTables *t)
{
    memset(t, 0, sizeof(*t));
    t->H = H; t->W = W;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        t->crc[i] = c;
    }
    t->pow8[0] = 0x00800000u;                     // x^8
    for (int k = 1; k < 32; ++k) t->pow8[k] = gf2_mulmod(t->pow8[k - 1], t->pow8[k - 1]);
    for (int y = 0; y < H; ++y) t->rowshift[y] = host_xpow8(t->pow8, 2ull * W * (unsigned long long)(H - 1 - y));
    for (int c = 0; c * 16 < W; ++c) t->colshift[c] = host_xpow8(t->pow8, 2ull * (unsigned long long)(W - std::min(W, 16 * (c + 1))));
    t->pow1[0] = 0x80000000u;                     // x^0
    for (int k = 1; k < 64; ++k) t->pow1[k] = gf2_mulmod(t->pow1[k - 1], t->pow8[0]);
    t->pow64[0] = 0x80000000u;
    for (int j = 1; j < ENC_NPIECE; ++j) t->pow64[j] = gf2_mulmod(t->pow64[j - 1], t->pow8[6]);      // x^(8 * 64)

    // ---- PNG, colour type 4 (gray + alpha), 16 bit: signature | IHDR | IDAT(zlib) | IEND ---------------
    EncStream &g = t->img;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t *p = g.prefix;
    memcpy(p, sig, 8); p += 8;
    h_be32(p, 13); memcpy(p + 4, "IHDR", 4); h_be32(p + 8, (uint32_t)W); h_be32(p + 12, (uint32_t)H);
    p[16] = 16; p[17] = 4; p[18] = 0; p[19] = 0; p[20] = 0;
    h_be32(p + 21, host_crc(t->crc, p + 4, 17)); p += 25;
    g.p_len_prefix = (int)(p - g.prefix);
    h_be32(p, 0); memcpy(p + 4, "IDAT", 4); p += 8;
    p[0] = 0x78; p[1] = 0x01; p += 2;             // zlib header: deflate, 32K window, fastest
    g.prefix_len = (int)(p - g.prefix);
    static const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
    memcpy(g.suffix, iend, 12); g.suffix_len = 12;
    g.head_len = 0;
    g.bpp = 4; g.seg_lanes = 4; g.row_lead = 1;
    build_codes(g, 4, 1 + 4 * W, kFreqImgLL, kFreqImgD, tuned);
    g.raw_len = (uint32_t)((size_t)H * (1 + 4 * (size_t)W));
    g.p_crc_prefix = g.p_csize_prefix = g.p_crc_suffix = g.p_csize_suffix = g.p_cdoff_suffix = -1;

    // ---- NPZ: zip archive with one deflated member "mask.npy" (np.savez_compressed(path, mask=...)) -----
    EncStream &m = t->msk;
    char dict[160];
    snprintf(dict, sizeof(dict), "{'descr': '<u2', 'fortran_order': False, 'shape': (%d, %d), }", H, W);
    int dl = (int)strlen(dict);
    const int pad = 64 - ((10 + dl + 1) % 64);     // numpy.lib.format: header block is a multiple of 64 bytes
    const int hl = dl + (pad % 64) + 1;
    uint8_t *h = m.head;
    h[0] = 0x93; memcpy(h + 1, "NUMPY", 5); h[6] = 1; h[7] = 0; h_le16(h + 8, (uint32_t)hl);
    memcpy(h + 10, dict, dl); memset(h + 10 + dl, ' ', hl - dl - 1); h[10 + hl - 1] = '\n';
    m.head_len = 10 + hl;
    const unsigned long long data_len = 2ull * H * W;
    m.raw_len = (uint32_t)(m.head_len + data_len);
    // CRC(M) = crc0(head) * x^(8*|data|) ^ crc0(data) ^ CRC(zeros(|M|)): the first and last term are constant
    const uint32_t aff = gf2_mulmod(0xffffffffu, host_xpow8(t->pow8, m.raw_len)) ^ 0xffffffffu;
    m.crc_const = gf2_mulmod(host_crc0(t->crc, m.head, m.head_len), host_xpow8(t->pow8, data_len)) ^ aff;
    static const char name[] = "mask.npy";
    p = m.prefix;
    h_le32(p, 0x04034b50u); h_le16(p + 4, 20); h_le16(p + 6, 0); h_le16(p + 8, 8); h_le16(p + 10, 0);
    h_le16(p + 12, 0x21);                          // 1980-01-01 00:00
    m.p_crc_prefix = 14; m.p_csize_prefix = 18;
    h_le32(p + 14, 0); h_le32(p + 18, 0); h_le32(p + 22, m.raw_len); h_le16(p + 26, 8); h_le16(p + 28, 0);
    memcpy(p + 30, name, 8);
    m.prefix_len = 38;
    p = m.suffix;
    h_le32(p, 0x02014b50u); h_le16(p + 4, 20); h_le16(p + 6, 20); h_le16(p + 8, 0); h_le16(p + 10, 8);
    h_le16(p + 12, 0); h_le16(p + 14, 0x21);
    m.p_crc_suffix = 16; m.p_csize_suffix = 20;
    h_le32(p + 16, 0); h_le32(p + 20, 0); h_le32(p + 24, m.raw_len); h_le16(p + 28, 8); h_le16(p + 30, 0);
    h_le16(p + 32, 0); h_le16(p + 34, 0); h_le16(p + 36, 0); h_le32(p + 38, 0x01800000u); h_le32(p + 42, 0);
    memcpy(p + 46, name, 8);
    p += 54;
    h_le32(p, 0x06054b50u); h_le16(p + 4, 0); h_le16(p + 6, 0); h_le16(p + 8, 1); h_le16(p + 10, 1);
    h_le32(p + 12, 54); h_le32(p + 16, 0); h_le16(p + 20, 0);
    m.p_cdoff_suffix = 54 + 16;
    m.suffix_len = 54 + 22;
    m.bpp = 2; m.seg_lanes = 8; m.row_lead = 0;
    build_codes(m, 2, 2 * W, kFreqMskLL, kFreqMskD, tuned);
    for (int j = 0; j < m.head_len; ++j) { m.head_off[j] = (uint16_t)m.head_bits; m.head_bits += (int)(m.lit[m.head[j]] >> 24); }
    m.p_len_prefix = -1;
}

static inline size_t enc_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
static size_t enc_fixed_smem(int n_max, bool events)
{
    return enc_align(frame_lut_bytes(n_max), 16) + 260 * 4 + 256 * 4 + 256 * 4 + (events ? ENC_NW * ENC_TCAP * 4 : 0);
}

}  // namespace cab200

using namespace cab200;

extern "C" size_t cab200_encode_tables_bytes(void) { return sizeof(EncTables); }

extern "C" int cab200_encode_tables(int H, int W, int huffman, void *tables_host_out)
{
    if (H <= 0 || W <= 0 || H > ENC_MAXDIM || W > ENC_MAXDIM || !tables_host_out) {
        set_error("cab200_encode_tables: H and W must be in [1, %d]", ENC_MAXDIM);
        return CAB200_EINVAL;
    }
    if (huffman != CAB200_HUFFMAN_FIXED && huffman != CAB200_HUFFMAN_TUNED) { set_error("cab200_encode_tables: huffman must be CAB200_HUFFMAN_FIXED or CAB200_HUFFMAN_TUNED"); return CAB200_EINVAL; }
    build_tables(H, W, huffman == CAB200_HUFFMAN_TUNED, (EncTables *)tables_host_out);
    return CAB200_OK;
}

extern "C" size_t cab200_encode_scratch_bytes(int n_pouch, int n_geom)
{
    return enc_align(sizeof(PouchEntry) * (size_t)std::max(n_pouch, 1), 16) +
           enc_align(sizeof(long long) * (size_t)std::max(n_geom, 1), 16);
}

extern "C" size_t cab200_encode_geometry_bytes(int n_cells, int H, int W) { return enc_align(geom_tab_bytes(n_cells, H, W), 16); }

extern "C" int cab200_encode_geometry(int n_cells, const uint16_t *label_img, const uint16_t *label_mask, int H, int W,
                                      const void *tables, void *geom_out, int32_t *max_events_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_cells <= 0 || n_cells > 65534 || H <= 0 || W <= 0 || H > ENC_MAXDIM || W > ENC_MAXDIM || !label_img ||
        !label_mask || !tables || !geom_out || !max_events_out) {
        set_error("cab200_encode_geometry: bad argument");
        return CAB200_EINVAL;
    }
    unsigned char *out = (unsigned char *)geom_out;
    CAB_CUDA(cudaMemsetAsync(out, 0, geom_tab_sums_bytes(n_cells), stream));
    CAB_CUDA(cudaMemsetAsync(out + geom_tab_sums_bytes(n_cells) + 2 * geom_tab_events_bytes(H, W) + 16, 0,
                             geom_tab_span_bytes(n_cells), stream));
    enc_geometry_kernel<<<148 * 4, 256, sizeof(uint32_t) * (size_t)W, stream>>>(
        label_img, label_mask, H, W, n_cells, (const EncTables *)tables, out);
    CAB_CUDA(cudaGetLastError());
    // static event lists; the two counters (largest event count of a segment) live at the end of the block
    int *d_max = (int *)(out + geom_tab_sums_bytes(n_cells) + 2 * geom_tab_events_bytes(H, W));
    CAB_CUDA(cudaMemsetAsync(d_max, 0, 2 * sizeof(int), stream));
    const int items = H * ((W + 511) / 512);
    const bool up_img = 1 + 4 * W <= 32768, up_msk = 2 * W <= 32768;     // the distances DEFLATE can express
    uint2 *ev_img = (uint2 *)(out + geom_tab_sums_bytes(n_cells));
    uint2 *ev_msk = (uint2 *)(out + geom_tab_sums_bytes(n_cells) + geom_tab_events_bytes(H, W));
    uint16_t *rc_img = (uint16_t *)(out + geom_tab_sums_bytes(n_cells) + 2 * geom_tab_events_bytes(H, W) + 16 + geom_tab_span_bytes(n_cells));
    uint16_t *rc_msk = rc_img + geom_tab_rowcnt_bytes(H, W) / 2;
    enc_events_kernel<<<(items + 7) / 8, 256, 0, stream>>>(label_img, H, W, n_cells, 64, up_img ? 1 : 0, 0, ev_img, rc_img, d_max);
    enc_events_kernel<<<(items + 7) / 8, 256, 0, stream>>>(label_mask, H, W, n_cells, 128, up_msk ? 1 : 0, 1, ev_msk, rc_msk, d_max + 1);
    CAB_CUDA(cudaGetLastError());
    int h_max[2] = {0, 0};
    CAB_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, stream));
    CAB_CUDA(cudaStreamSynchronize(stream));
    // the event lists address the key LUT with 16-bit byte offsets: larger label ranges take the per-pixel parse
    *max_events_out = (n_cells > 32766) ? (1 << 30) : std::max(h_max[0], h_max[1]);
    return CAB200_OK;
}

extern "C" int cab200_encode_batch(
    int n_pouch, const int32_t *geom_id, int n_geom, const int32_t *g_ncell, const int64_t *cell_off,
    const uint16_t *label_img, const uint16_t *label_mask, int H, int W, int n_frames,
    const double *ca_frames, double threshold, int quirk_mode, int want_image, int want_mask,
    const void *tables, const void *geom_tabs, const int64_t *g_tab_off, const int32_t *g_max_events, uint8_t *arena, size_t arena_bytes, unsigned long long *cursor,
    void *records_out, int32_t *n_active_out, int stage_bytes, void *scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pouch < 0 || H <= 0 || W <= 0 || H > ENC_MAXDIM || W > ENC_MAXDIM || n_frames < 0) {
        set_error("cab200_encode_batch: bad sizes (H, W in [1, %d])", ENC_MAXDIM);
        return CAB200_EINVAL;
    }
    if (n_pouch == 0 || n_frames == 0) return CAB200_OK;
    if (!geom_id || !g_ncell || !cell_off || !label_img || !label_mask || !ca_frames || !tables || !geom_tabs ||
        !g_tab_off || !arena || !cursor || !records_out || !scratch || n_geom <= 0) {
        set_error("cab200_encode_batch: null pointer argument");
        return CAB200_EINVAL;
    }
    if (((uintptr_t)arena & 15u) != 0) { set_error("cab200_encode_batch: arena must be 16-byte aligned"); return CAB200_EINVAL; }
    if (scratch_bytes < cab200_encode_scratch_bytes(n_pouch, n_geom)) {
        set_error("cab200_encode_batch: scratch too small");
        return CAB200_ESIZE;
    }
    std::vector<PouchEntry> tab(n_pouch);
    int n_max = 0;
    for (int p = 0; p < n_pouch; ++p) {
        tab[p].cell_off = cell_off[p];
        tab[p].geom = geom_id[p];
        if (geom_id[p] < 0 || geom_id[p] >= n_geom) { set_error("cab200_encode_batch: pouch %d: bad geom_id", p); return CAB200_EINVAL; }
        tab[p].n = g_ncell[geom_id[p]];
        if (tab[p].n > 65534 || tab[p].n <= 0 || cell_off[p + 1] - cell_off[p] != tab[p].n) {
            set_error("cab200_encode_batch: pouch %d: bad cell count", p);
            return CAB200_EINVAL;
        }
        n_max = std::max(n_max, tab[p].n);
    }
    const long long pf = (long long)n_pouch * n_frames;
    if (pf > 0x7fffffffLL) { set_error("cab200_encode_batch: too many frames"); return CAB200_ESIZE; }
    // events per lane: the largest static event count of a row segment over the geometries of this call
    // (4 or 8 per lane); unknown or above 256 -> the per-pixel parse
    int max_ev = g_max_events ? 0 : ENC_ECAP + 1;
    if (g_max_events)
        for (int p = 0; p < n_pouch; ++p) max_ev = std::max(max_ev, g_max_events[geom_id[p]] > 0 ? g_max_events[geom_id[p]] : ENC_ECAP + 1);
    const size_t fixed = enc_fixed_smem(n_max, max_ev <= ENC_ECAP);
    // staging: by default half of an SM's shared memory per CTA (two CTAs per SM)
    size_t cap = (stage_bytes > 0) ? (size_t)stage_bytes : (size_t)(112 * 1024) - fixed - 4608;
    cap &= ~(size_t)15;
    if ((stage_bytes <= 0 && fixed + 4608 + 4096 > 112 * 1024) || cap < 1024 || fixed + cap > 224 * 1024) {
        set_error("cab200_encode_batch: staging buffer of %zu bytes does not fit (tables take %zu)", cap, fixed);
        return CAB200_ESIZE;
    }
    CAB_CUDA(cudaMemcpyAsync(scratch, tab.data(), sizeof(PouchEntry) * (size_t)n_pouch, cudaMemcpyHostToDevice, stream));
    long long *d_tab_off = (long long *)((unsigned char *)scratch + enc_align(sizeof(PouchEntry) * (size_t)n_pouch, 16));
    CAB_CUDA(cudaMemcpyAsync(d_tab_off, g_tab_off, sizeof(long long) * (size_t)n_geom, cudaMemcpyHostToDevice, stream));
    EncodeArgs a;
    a.geom_tabs = (const unsigned char *)geom_tabs;
    a.geom_tab_off = d_tab_off;
    a.pouches = (const PouchEntry *)scratch;
    a.label_img = label_img; a.label_mask = label_mask; a.ca_frames = ca_frames;
    a.tab = (const EncTables *)tables;
    a.n_frames = n_frames; a.H = H; a.W = W;
    a.threshold = threshold; a.quirk = quirk_mode; a.want_img = want_image; a.want_msk = want_mask;
    a.lut_bytes = (int)enc_align(frame_lut_bytes(n_max), 16);
    a.stage_cap = (int)cap;
    a.arena = arena; a.cursor = cursor; a.arena_bytes = arena_bytes;
    a.records = (EncRecord *)records_out;
    a.n_active_out = n_active_out;
    const size_t smem = fixed + cap;
    void (*kern)(const EncodeArgs) = (max_ev <= 64) ? encode_kernel<2> : (max_ev <= 96) ? encode_kernel<3> : (max_ev <= 128) ? encode_kernel<4>
                                     : (max_ev <= 256) ? encode_kernel<8> : encode_kernel<0>;      // events per lane
    CAB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)pf, ENC_NT, smem, stream>>>(a);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}

extern "C" int cab200_publish_to_host(void *dst_pinned_host, const void *src_device, size_t bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bytes == 0) return CAB200_OK;
    if (!dst_pinned_host || !src_device || (bytes & 15u) || ((uintptr_t)dst_pinned_host & 15u) || ((uintptr_t)src_device & 15u)) {
        set_error("cab200_publish_to_host: pointers and size must be 16-byte aligned");
        return CAB200_EINVAL;
    }
    void *dst_dev = nullptr;
    CAB_CUDA(cudaHostGetDevicePointer(&dst_dev, dst_pinned_host, 0));
    const size_t n16 = bytes / 16;
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 148 * 2);
    publish_kernel<<<grid, 256, 0, stream>>>((uint4 *)dst_dev, (const uint4 *)src_device, n16);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}
