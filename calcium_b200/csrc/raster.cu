// raster.cu -- K2 (per-frame image / active mask / instance mask) and K0 (static cells_mask).
//
// K2 replaces, per (pouch, frame): Pouch.generate_image default branch (core/pouch.py:412-453),
// get_active_cells (:563-565), get_cell_masks(active_only=True) (:540-546) and
// create_instance_mask_from_cells (utils/sam2_data_generator.py:31-35).  The reference repaints n
// polygons per frame; here a frame is a pure gather through two static label maps: per frame the
// CTA builds three small look-up tables in shared memory (gray level, active flag, instance id per
// label value) and then streams the pixels with 16-byte loads and stores.  HBM-write bound:
// 4 bytes written per pixel (image 2 + instance 2), +4 when the int32 active mask is requested;
// the label maps (4 bytes read per pixel) stay L2 resident per geometry.
#include <algorithm>
#include <climits>
#include <vector>

#include "common.cuh"
#include "frame_lut.cuh"

namespace cab200 {

struct RasterArgs {
    const PouchEntry *pouches;   // [P]
    const uint16_t *label_img;   // [G][H][W]
    const uint16_t *label_mask;  // [G][H][W]
    const double *ca_frames;
    int n_frames, npix, tiles;
    double threshold;
    int quirk;
    uint8_t *image_out;
    int32_t *active_out;
    uint16_t *instance_out;
    int32_t *n_active_out;
};

constexpr int RT = 256;   // threads per CTA

__global__ void __launch_bounds__(RT, 6) raster_kernel(const RasterArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tile = blockIdx.x % a.tiles;
    const int pf = blockIdx.x / a.tiles;
    const int f = pf % a.n_frames;
    const int pi = pf / a.n_frames;
    const PouchEntry pe = a.pouches[pi];
    const int n = pe.n;
    const double *ca = a.ca_frames + (size_t)pe.cell_off * 4 * a.n_frames + (size_t)f * 4 * n;

    // LUTs indexed by label value v in [0, n] (frame_lut.cuh)
    const FrameLut lut = frame_lut_carve(smem_raw, n);
    const uint16_t *inst_lut = lut.inst_lut;
    const uint8_t *gray_lut = lut.gray_lut, *act8 = lut.act8;
    const int tid = threadIdx.x;
    const int n_active = frame_lut_build<RT>(lut, ca, n, a.threshold, a.quirk);
    if (tid == 0 && tile == 0 && a.n_active_out) a.n_active_out[pf] = n_active;

    // ---- stream the pixels ---------------------------------------------------------------------
    const size_t gbase = (size_t)pe.geom * a.npix;
    const uint16_t *limg = a.label_img + gbase;
    const uint16_t *lmsk = a.label_mask + gbase;
    const size_t obase = (size_t)pf * a.npix;
    const int groups = a.npix >> 3;                       // 8-pixel groups
    const int gper = (groups + a.tiles - 1) / a.tiles;
    const int g0 = tile * gper, g1 = min(g0 + gper, groups);
    const bool vec_ok = ((a.npix & 7) == 0);
    if (vec_ok) {
        // two 8-pixel groups per thread and iteration: all four label loads are issued before any
        // look-up, which doubles the bytes in flight per thread (the kernel is latency/HBM bound)
        auto emit = [&](int g, const uint4 li, const uint4 lm) {
            const size_t px = (size_t)g * 8;
            if (a.image_out) {
                const uint32_t w[4] = {li.x, li.y, li.z, li.w};
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t l0 = w[q] & 0xffffu, l1 = w[q] >> 16;
                    const uint32_t p0 = l0 ? ((uint32_t)gray_lut[l0] | 0xff00u) : 0u;
                    const uint32_t p1 = l1 ? ((uint32_t)gray_lut[l1] | 0xff00u) : 0u;
                    o[q] = p0 | (p1 << 16);
                }
                __stcs(reinterpret_cast<uint4 *>(a.image_out + (obase + px) * 2),
                       make_uint4(o[0], o[1], o[2], o[3]));
            }
            if (a.instance_out || a.active_out) {
                const uint32_t w[4] = {lm.x, lm.y, lm.z, lm.w};
                if (a.instance_out) {
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        o[q] = (uint32_t)inst_lut[w[q] & 0xffffu] | ((uint32_t)inst_lut[w[q] >> 16] << 16);
                    __stcs(reinterpret_cast<uint4 *>(a.instance_out + obase + px),
                           make_uint4(o[0], o[1], o[2], o[3]));
                }
                if (a.active_out) {
                    int32_t o[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t l0 = w[q] & 0xffffu, l1 = w[q] >> 16;
                        o[2 * q] = act8[l0] ? (int32_t)l0 : 0;
                        o[2 * q + 1] = act8[l1] ? (int32_t)l1 : 0;
                    }
                    int4 *dst = reinterpret_cast<int4 *>(a.active_out + obase + px);
                    __stcs(dst, make_int4(o[0], o[1], o[2], o[3]));
                    __stcs(dst + 1, make_int4(o[4], o[5], o[6], o[7]));
                }
            }
        };
        const bool need_img = a.image_out != nullptr;
        const bool need_msk = a.instance_out != nullptr || a.active_out != nullptr;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int g = g0 + tid; g < g1; g += 2 * RT) {
            const int gb = g + RT;
            const bool two = gb < g1;
            const uint4 li0 = need_img ? __ldg(reinterpret_cast<const uint4 *>(limg + (size_t)g * 8)) : z4;
            const uint4 lm0 = need_msk ? __ldg(reinterpret_cast<const uint4 *>(lmsk + (size_t)g * 8)) : z4;
            const uint4 li1 = (need_img && two) ? __ldg(reinterpret_cast<const uint4 *>(limg + (size_t)gb * 8)) : z4;
            const uint4 lm1 = (need_msk && two) ? __ldg(reinterpret_cast<const uint4 *>(lmsk + (size_t)gb * 8)) : z4;
            emit(g, li0, lm0);
            if (two) emit(gb, li1, lm1);
        }
    } else {
        const int per = (a.npix + a.tiles - 1) / a.tiles;
        const int p0 = tile * per, p1 = min(p0 + per, a.npix);
        for (int px = p0 + tid; px < p1; px += RT) {
            if (a.image_out) {
                const uint32_t l = limg[px];
                a.image_out[(obase + px) * 2] = l ? gray_lut[l] : 0;
                a.image_out[(obase + px) * 2 + 1] = l ? 255 : 0;
            }
            const uint32_t m = lmsk[px];
            if (a.instance_out) a.instance_out[obase + px] = inst_lut[m];
            if (a.active_out) a.active_out[obase + px] = act8[m] ? (int32_t)m : 0;
        }
    }
}

// ---- K0: cells_mask with scikit-image's point_in_polygon rule -----------------------------------
// One CTA per polygon; integer pixel vertices make the O'Rourke crossing test exact in int64:
// the reference's quotient (x0*y1 - x1*y0)/(y1 - y0) is only compared with 0, and its eps=1e-12
// vertex test degenerates to equality on integers.
__global__ void __launch_bounds__(128) cells_mask_kernel(int n_cells, const int32_t *poly_off,
                                                         const int32_t *pts, int H, int W,
                                                         uint32_t *label32)
{
    constexpr int MAXV = 64;
    __shared__ int vx[MAXV], vy[MAXV];
    __shared__ int box[4];
    const int cell = blockIdx.x;
    const int b = poly_off[cell], nv = min(poly_off[cell + 1] - b, MAXV);
    if (nv <= 0) return;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) { vx[i] = pts[2 * (b + i)]; vy[i] = pts[2 * (b + i) + 1]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int minx = vx[0], maxx = vx[0], miny = vy[0], maxy = vy[0];
        for (int i = 1; i < nv; ++i) {
            minx = min(minx, vx[i]); maxx = max(maxx, vx[i]);
            miny = min(miny, vy[i]); maxy = max(maxy, vy[i]);
        }
        box[0] = max(0, minx); box[1] = min(W - 1, maxx);     // minc..maxc
        box[2] = max(0, miny); box[3] = min(H - 1, maxy);     // minr..maxr
    }
    __syncthreads();
    const int minc = box[0], maxc = box[1], minr = box[2], maxr = box[3];
    if (maxc < minc || maxr < minr) return;
    const int bw = maxc - minc + 1, bh = maxr - minr + 1;
    const uint32_t id = (uint32_t)cell + 1u;
    for (int q = threadIdx.x; q < bw * bh; q += blockDim.x) {
        const int y = minr + q / bw, x = minc + q % bw;
        long long x1 = vx[nv - 1] - x, y1 = vy[nv - 1] - y;
        unsigned rc = 0, lc = 0;
        bool vertex = false;
        for (int i = 0; i < nv; ++i) {
            const long long x0 = vx[i] - x, y0 = vy[i] - y;
            if (x0 == 0 && y0 == 0) { vertex = true; break; }
            if ((y0 > 0) != (y1 > 0) || (y0 < 0) != (y1 < 0)) {
                const long long num = x0 * y1 - x1 * y0, den = y1 - y0;
                const int sg = (num > 0 ? 1 : (num < 0 ? -1 : 0)) * (den > 0 ? 1 : -1);
                if ((y0 > 0) != (y1 > 0) && sg > 0) ++rc;
                if ((y0 < 0) != (y1 < 0) && sg < 0) ++lc;
            }
            x1 = x0; y1 = y0;
        }
        const bool inside = vertex || ((rc & 1) != (lc & 1)) || (rc & 1);
        if (inside) {
            // highest id wins (later cells overwrite, core/pouch.py:285): max on the u16 half
            const size_t px = (size_t)y * W + x;
            uint32_t *word = label32 + (px >> 1);
            const uint32_t sh = (px & 1) * 16;
            uint32_t old = *word;
            while (((old >> sh) & 0xffffu) < id) {
                const uint32_t want = (old & ~(0xffffu << sh)) | (id << sh);
                const uint32_t prev = atomicCAS(word, old, want);
                if (prev == old) break;
                old = prev;
            }
        }
    }
}

// ---- K0b: paint-order label map with OpenCV's fillPoly rule -------------------------------------
// generate_image paints every cell with cv2.fillPoly in cell order (core/pouch.py:441-453); which
// cell owns a pixel in the end is "the highest id whose fillPoly touches it".  For one polygon with
// integer vertices (shift 0, 8-connected, not anti-aliased) OpenCV's fillPoly is
//   (1) every edge drawn with cv::LineIterator, endpoints ordered left to right (CollectPolyEdges),
//   (2) scanline fill (FillEdgeCollection): an edge is active on rows y0 <= y < y1, its x advances
//       by dx = trunc(((x1 - x0) << 16) / (y1 - y0)) per row in 16.16 fixed point, and between
//       consecutive pairs of x-sorted active edges the pixels ceil(x_left) .. floor(x_right) are set.
// Restated here; tests compare it with the real cv2.fillPoly on every geometry / output size used
// and on random concave polygons.
__device__ __forceinline__ void label_max(uint32_t *label32, int W, int x, int y, uint32_t id)
{
    const size_t px = (size_t)y * W + x;
    uint32_t *word = label32 + (px >> 1);
    const uint32_t sh = (px & 1) * 16;
    uint32_t old = *word;
    while (((old >> sh) & 0xffffu) < id) {
        const uint32_t want = (old & ~(0xffffu << sh)) | (id << sh);
        const uint32_t prev = atomicCAS(word, old, want);
        if (prev == old) break;
        old = prev;
    }
}

__global__ void __launch_bounds__(128) paint_label_kernel(int n_cells, const int32_t *poly_off,
                                                          const int32_t *pts, int H, int W,
                                                          uint32_t *label32)
{
    constexpr int MAXV = 64;
    __shared__ int vx[MAXV], vy[MAXV];
    __shared__ int s_ymin, s_ymax;
    const int cell = blockIdx.x;
    const int b = poly_off[cell], nv = min(poly_off[cell + 1] - b, MAXV);
    if (nv <= 0) return;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) { vx[i] = pts[2 * (b + i)]; vy[i] = pts[2 * (b + i) + 1]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int lo = INT_MAX, hi = INT_MIN;
        for (int i = 0; i < nv; ++i) { lo = min(lo, vy[i]); hi = max(hi, vy[i]); }
        s_ymin = lo; s_ymax = hi;
    }
    __syncthreads();
    const uint32_t id = (uint32_t)cell + 1u;
    // (1) outlines: one thread per edge, cv::LineIterator (8-connected, left to right)
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        int x0 = vx[(i + nv - 1) % nv], y0 = vy[(i + nv - 1) % nv], x1 = vx[i], y1 = vy[i];
        int dx = x1 - x0, dy = y1 - y0;
        if (dx < 0) { x0 = x1; y0 = y1; dx = -dx; dy = -dy; }
        const int sy = dy >= 0 ? 1 : -1;
        dy = abs(dy);
        int x = x0, y = y0;
        if (dy > dx) {
            int err = dy - (dx + dx);
            const int plus = dy + dy, minus = -(dx + dx);
            for (int c = 0; c <= dy; ++c) {
                if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) label_max(label32, W, x, y, id);
                const bool m = err < 0;
                err += minus + (m ? plus : 0);
                y += sy;
                if (m) x += 1;
            }
        } else {
            int err = dx - (dy + dy);
            const int plus = dx + dx, minus = -(dy + dy);
            for (int c = 0; c <= dx; ++c) {
                if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) label_max(label32, W, x, y, id);
                const bool m = err < 0;
                err += minus + (m ? plus : 0);
                x += 1;
                if (m) y += sy;
            }
        }
    }
    // (2) interior: one thread per scanline
    const int ymin = max(s_ymin, 0), ymax = min(s_ymax, H);
    for (int y = ymin + threadIdx.x; y < ymax; y += blockDim.x) {
        long long xs[MAXV];
        int cnt = 0;
        for (int i = 0; i < nv; ++i) {
            int xa = vx[(i + nv - 1) % nv], ya = vy[(i + nv - 1) % nv], xb = vx[i], yb = vy[i];
            if (ya == yb) continue;
            const long long num = ((long long)(xb - xa)) << 16;
            const long long slope = num / (yb - ya);                 // C++ division truncates toward zero
            if (ya > yb) { const int t = ya; ya = yb; yb = t; xa = xb; }
            if (y < ya || y >= yb) continue;
            const long long xv = ((long long)xa << 16) + (long long)(y - ya) * slope;
            int j = cnt++;
            while (j > 0 && xs[j - 1] > xv) { xs[j] = xs[j - 1]; --j; }   // insertion sort by x
            xs[j] = xv;
        }
        for (int k = 0; k + 1 < cnt; k += 2) {
            long long a = (xs[k] + 65535) >> 16, bq = xs[k + 1] >> 16;
            if (a < W && bq >= 0) {
                if (a < 0) a = 0;
                if (bq >= W) bq = W - 1;
                for (long long x = a; x <= bq; ++x) label_max(label32, W, (int)x, y, id);
            }
        }
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace cab200

using namespace cab200;

extern "C" size_t cab200_raster_scratch_bytes(int n_pouch)
{
    return align_up(sizeof(PouchEntry) * (size_t)std::max(n_pouch, 1), 16);
}

extern "C" int cab200_raster_batch(
    int n_pouch, const int32_t *geom_id, const int32_t *g_ncell, const int64_t *cell_off,
    const uint16_t *label_img, const uint16_t *label_mask, int H, int W, int n_frames,
    const double *ca_frames, double threshold, int quirk_mode, uint8_t *image_out,
    int32_t *active_mask_out, uint16_t *instance_out, int32_t *n_active_out, void *scratch,
    size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pouch < 0 || H <= 0 || W <= 0 || n_frames < 0) {
        set_error("cab200_raster_batch: bad sizes");
        return CAB200_EINVAL;
    }
    if (n_pouch == 0 || n_frames == 0) return CAB200_OK;
    if (!geom_id || !g_ncell || !cell_off || !label_img || !label_mask || !ca_frames || !scratch) {
        set_error("cab200_raster_batch: null pointer argument");
        return CAB200_EINVAL;
    }
    if (scratch_bytes < cab200_raster_scratch_bytes(n_pouch)) {
        set_error("cab200_raster_batch: scratch too small");
        return CAB200_ESIZE;
    }
    if ((size_t)H * W >= (1u << 30)) { set_error("cab200_raster_batch: image too large"); return CAB200_ESIZE; }
    std::vector<PouchEntry> tab(n_pouch);
    int n_max = 0;
    for (int p = 0; p < n_pouch; ++p) {
        tab[p].cell_off = cell_off[p];
        tab[p].geom = geom_id[p];
        tab[p].n = g_ncell[geom_id[p]];
        if (tab[p].n > 65534 || tab[p].n <= 0 || cell_off[p + 1] - cell_off[p] != tab[p].n) {
            set_error("cab200_raster_batch: pouch %d: bad cell count", p);
            return CAB200_EINVAL;
        }
        n_max = std::max(n_max, tab[p].n);
    }
    CAB_CUDA(cudaMemcpyAsync(scratch, tab.data(), sizeof(PouchEntry) * (size_t)n_pouch,
                             cudaMemcpyHostToDevice, stream));
    RasterArgs a;
    a.pouches = (const PouchEntry *)scratch;
    a.label_img = label_img; a.label_mask = label_mask; a.ca_frames = ca_frames;
    a.n_frames = n_frames; a.npix = H * W;
    a.threshold = threshold; a.quirk = quirk_mode;
    a.image_out = image_out; a.active_out = active_mask_out; a.instance_out = instance_out;
    a.n_active_out = n_active_out;
    // whole frame per CTA when there are plenty of frames; split frames when few so all SMs stream
    const long long pf = (long long)n_pouch * n_frames;
    int tiles = 1;
    while ((long long)tiles * pf < 148 * 8 && tiles < 32) tiles *= 2;
    a.tiles = tiles;
    const size_t smem = frame_lut_bytes(n_max);
    if (smem > 48 * 1024)
        CAB_CUDA(cudaFuncSetAttribute((const void *)raster_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (pf * tiles > 0x7fffffffLL) { set_error("cab200_raster_batch: too many frames"); return CAB200_ESIZE; }
    raster_kernel<<<(unsigned)(pf * tiles), RT, smem, stream>>>(a);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}

extern "C" int cab200_paint_label_map(int n_cells, const int32_t *poly_off, const int32_t *pts, int H,
                                      int W, uint16_t *label_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_cells <= 0 || n_cells > 65534 || H <= 0 || W <= 0 || !poly_off || !pts || !label_out) {
        set_error("cab200_paint_label_map: bad argument");
        return CAB200_EINVAL;
    }
    if (((size_t)H * W) & 1) { set_error("cab200_paint_label_map: H*W must be even"); return CAB200_EINVAL; }
    CAB_CUDA(cudaMemsetAsync(label_out, 0, sizeof(uint16_t) * (size_t)H * W, stream));
    paint_label_kernel<<<n_cells, 128, 0, stream>>>(n_cells, poly_off, pts, H, W,
                                                    reinterpret_cast<uint32_t *>(label_out));
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}

extern "C" int cab200_cells_mask(int n_cells, const int32_t *poly_off, const int32_t *pts, int H,
                                 int W, uint16_t *label_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_cells <= 0 || n_cells > 65534 || H <= 0 || W <= 0 || !poly_off || !pts || !label_out) {
        set_error("cab200_cells_mask: bad argument");
        return CAB200_EINVAL;
    }
    if (((size_t)H * W) & 1) { set_error("cab200_cells_mask: H*W must be even"); return CAB200_EINVAL; }
    CAB_CUDA(cudaMemsetAsync(label_out, 0, sizeof(uint16_t) * (size_t)H * W, stream));
    cells_mask_kernel<<<n_cells, 128, 0, stream>>>(n_cells, poly_off, pts, H, W,
                                                   reinterpret_cast<uint32_t *>(label_out));
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}
