// edges.cu -- K0c (static edge maps) and K2b (per-frame image with the edge_blur / with_border
// branches of Pouch.generate_image, core/pouch.py:436-481) -- SURVEY.md section 8 row f1.
//
// The reference builds, per frame, an edge mask with one cv2.polylines per cell, blurs it
// (cv2.filter2D with the mean / motion kernel of :494-513), dilates it 3x3 and blends the blurred image
// into the sharp one with weight dilated/255.  The edge mask and its blurred + dilated version depend
// on the geometry only, so they are built once per (geometry, output size, kernel) by K0c; K2b then
// needs one filter2D of the frame's gray channel per pixel whose blend weight is not zero, fused with the
// gather through the paint-order label map (the sharp image is never materialised).  The alpha channel
// (255 inside the disc, 0 outside, whatever the frame) is blurred and blended once, in K0c.
//
// Arithmetic restated from OpenCV (4.13 in this image; not vendored by the reference):
//   polylines, thickness 1, shift 0, LINE_8: cv::LineIterator per segment (left to right) -- the same
//     iterator fillPoly outlines with (raster.cu, K0b);
//   filter2D, uint8 source, float32 kernel, ddepth -1: correlation with the anchor at ksize/2,
//     BORDER_REFLECT_101, float32 accumulator starting at 0, taps in row-major order of the kernel's
//     non-zero elements, one fused multiply-add per tap (the AVX2/FMA path of FilterVec_8u),
//     cvRound (half to even) and saturation to [0, 255];
//   dilate 3x3: maximum over the in-image neighbours;
//   blend (numpy): (img * (1 - m) + blurred * m).astype(uint8), m = dilated / 255.0, float64, one
//     rounding per operation, truncation.
// For odd kernel sizes (the GUI offers 3, 5, 7, 9) and powers of two no accumulated sum can come
// within rounding error of a .5 tie, so the result does not depend on the accumulation details; the
// oracle (oracle/edges_oracle.py) and tests/ pin all of it against the real cv2 calls.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "frame_lut.cuh"

namespace cab200 {

constexpr int ET = 256;          // threads per CTA (K2b); also the size of the blend tables
static_assert(ET == 256, "one blend-table entry per thread");
constexpr int EDGE_KMAX = 10;    // largest blur kernel (OpenCV switches to a DFT-based filter at 11)

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - i : i;
}

// ---- K0c (1): polygon outlines --------------------------------------------------------------------
__global__ void __launch_bounds__(64) edge_lines_kernel(int n_cells, const int32_t *poly_off, const int32_t *pts,
                                                        int H, int W, uint8_t *edges)
{
    const int cell = blockIdx.x;
    const int b = poly_off[cell], nv = poly_off[cell + 1] - b;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int ip = (i + nv - 1) % nv;
        int x0 = pts[2 * (b + ip)], y0 = pts[2 * (b + ip) + 1];
        const int x1 = pts[2 * (b + i)], y1 = pts[2 * (b + i) + 1];
        int dx = x1 - x0, dy = y1 - y0;
        if (dx < 0) { x0 = x1; y0 = y1; dx = -dx; dy = -dy; }
        const int sy = dy >= 0 ? 1 : -1;
        dy = abs(dy);
        const bool steep = dy > dx;
        const int major = steep ? dy : dx, minor = steep ? dx : dy;
        int err = major - 2 * minor, x = x0, y = y0;
        for (int c = 0; c <= major; ++c) {
            if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) edges[(size_t)y * W + x] = 255;
            const bool m = err < 0;
            err += -2 * minor + (m ? 2 * major : 0);
            if (steep) { y += sy; x += m ? 1 : 0; }
            else       { x += 1;  y += m ? sy : 0; }
        }
    }
}

// filter2D of a uint8 map at one pixel (taps in row-major kernel order, FMA chain)
__device__ __forceinline__ int blur_u8_at(const uint8_t *src, int H, int W, int y, int x, int ks, int motion, float kf)
{
    const int a = ks / 2;
    float acc = 0.0f;
    const int i0 = motion ? a : 0, i1 = motion ? a + 1 : ks;
    for (int i = i0; i < i1; ++i) {
        const int yy = reflect101(y + i - a, H);
        for (int j = 0; j < ks; ++j) {
            const int xx = reflect101(x + j - a, W);
            acc = __fmaf_rn(kf, (float)src[(size_t)yy * W + xx], acc);
        }
    }
    return min(max(__float2int_rn(acc), 0), 255);
}

// ---- K0c (2): dilate3x3(filter2D(edges)) and the static alpha channel ----------------------------------
// The alpha channel of the sharp image is 255 inside the disc and 0 outside whatever the frame, so its
// blurred and blended version (:463-467, channel 1) is a property of the geometry: computed here once.
__global__ void __launch_bounds__(256) edge_dilate_kernel(const uint8_t *edges, const uint16_t *label_img, int H, int W,
                                                          int ks, int motion, float kf, uint8_t *dilated,
                                                          uint8_t *alpha_out)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= H * W) return;
    const int y = px / W, x = px % W;
    int best = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
                best = max(best, blur_u8_at(edges, H, W, yy, xx, ks, motion, kf));
        }
    dilated[px] = (uint8_t)best;
    if (alpha_out) {
        int al = label_img[px] ? 255 : 0;
        if (best) {
            const int a = ks / 2;
            float acc = 0.0f;
            const int i0 = motion ? a : 0, i1 = motion ? a + 1 : ks;
            for (int i = i0; i < i1; ++i) {
                const int yy = reflect101(y + i - a, H);
                for (int j = 0; j < ks; ++j) {
                    const int xx = reflect101(x + j - a, W);
                    acc = __fmaf_rn(kf, label_img[(size_t)yy * W + xx] ? 255.0f : 0.0f, acc);
                }
            }
            const int ba = min(max(__float2int_rn(acc), 0), 255);
            const double m = __ddiv_rn((double)best, 255.0), om = __dsub_rn(1.0, m);
            al = (int)__dadd_rn(__dmul_rn((double)al, om), __dmul_rn((double)ba, m));
        }
        alpha_out[px] = (uint8_t)al;
    }
}

// ---- K2b ---------------------------------------------------------------------------------------
struct EdgeArgs {
    const PouchEntry *pouches;
    const uint16_t *label_img;   // [G][H][W]
    const uint8_t *edges;        // [G][H][W]
    const uint8_t *dilated;      // [G][H][W]
    const uint8_t *alpha;        // [G][H][W] static blended alpha channel (edge_blur)
    const double *ca_frames;
    int n_frames, H, W, tiles;
    int mode, ks, motion, band;  // band = output rows per shared-memory tile
    int stride;                  // floats per tile row
    float kf;
    uint8_t *image_out;
};

// one reflection is enough when the kernel is not larger than the image (checked on the host)
__device__ __forceinline__ int reflect_once(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

// with_border (+ edge_blur): gray halved (:470) or zeroed (:481) on the outline pixels
__device__ __forceinline__ void border_modes(const EdgeArgs &a, int tile, const uint16_t *limg, const uint8_t *edges,
                                             const uint8_t *gray_lut, uint8_t *out, size_t npix)
{
    const int tid = threadIdx.x;
    const bool halve = a.mode == CAB200_EDGE_BLUR_BORDER;
    if ((npix & 7) == 0) {
        const int groups = (int)(npix >> 3), gper = (groups + a.tiles - 1) / a.tiles;
        const int g0 = tile * gper, g1 = min(g0 + gper, groups);
        for (int g = g0 + tid; g < g1; g += ET) {
            const uint4 li = __ldg(reinterpret_cast<const uint4 *>(limg) + g);
            const uint2 ed = __ldg(reinterpret_cast<const uint2 *>(edges) + g);
            const uint32_t w[4] = {li.x, li.y, li.z, li.w};
            const uint32_t e[2] = {ed.x, ed.y};
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t pix[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t l = (w[q] >> (16 * h)) & 0xffffu;
                    uint32_t gr = l ? gray_lut[l] : 0u;
                    if ((e[q >> 1] >> (8 * (2 * (q & 1) + h))) & 0xffu) gr = halve ? (gr >> 1) : 0u;
                    pix[h] = l ? (gr | 0xff00u) : 0u;
                }
                o[q] = pix[0] | (pix[1] << 16);
            }
            __stcs(reinterpret_cast<uint4 *>(out) + g, make_uint4(o[0], o[1], o[2], o[3]));
        }
    } else {
        const size_t per = (npix + a.tiles - 1) / a.tiles;
        const size_t p0 = tile * per, p1 = min(p0 + per, npix);
        for (size_t px = p0 + tid; px < p1; px += ET) {
            const uint32_t l = limg[px];
            uint32_t g = l ? gray_lut[l] : 0;
            if (edges[px]) g = halve ? (g >> 1) : 0;
            reinterpret_cast<uchar2 *>(out)[px] = make_uchar2((unsigned char)g, l ? 255 : 0);
        }
    }
}

// KS > 0: compile-time kernel size, four pixels per thread (W % 4 == 0, KS <= min(H, W));
// KS == 0: any size / shape, one pixel per thread.
template <int KS, bool MOTION>
__global__ void __launch_bounds__(ET, 3) raster_edges_kernel(const EdgeArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile = blockIdx.x % a.tiles;
    const int pf = blockIdx.x / a.tiles;
    const int f = pf % a.n_frames;
    const int pi = pf / a.n_frames;
    const PouchEntry pe = a.pouches[pi];
    const int n = pe.n, H = a.H, W = a.W, tid = threadIdx.x;
    const double *ca = a.ca_frames + (size_t)pe.cell_off * 4 * a.n_frames + (size_t)f * 4 * n;

    const FrameLut lut = frame_lut_carve(smem_raw, n);
    frame_lut_build<ET>(lut, ca, n, 0.1, 1);          // only the gray table is used here
    const uint8_t *gray_lut = lut.gray_lut;
    unsigned char *after = smem_raw + ((frame_lut_bytes(n) + 15) & ~(size_t)15);

    const size_t npix = (size_t)H * W;
    const uint16_t *limg = a.label_img + (size_t)pe.geom * npix;
    uint8_t *out = a.image_out + (size_t)pf * npix * 2;

    if (a.mode != CAB200_EDGE_BLUR) {
        border_modes(a, tile, limg, a.edges + (size_t)pe.geom * npix, gray_lut, out, npix);
        return;
    }

    // ---- edge_blur: blend weights per dilated value, then band by band through a float tile of gray ----
    const uint8_t *dil = a.dilated + (size_t)pe.geom * npix;
    const uint8_t *alpha = a.alpha + (size_t)pe.geom * npix;
    double *m_tab = reinterpret_cast<double *>(after);      // m = d / 255.0          (:461)
    double *om_tab = m_tab + 256;                           // 1 - m                  (:466)
    float *tilebuf = reinterpret_cast<float *>(om_tab + 256);
    {
        const double m = __ddiv_rn((double)tid, 255.0);
        m_tab[tid] = m;
        om_tab[tid] = __dsub_rn(1.0, m);
    }
    const int ks = KS ? KS : a.ks, an = ks / 2, R = a.band, stride = a.stride;
    const int tw = W + ks - 1, th = R + ks - 1;
    const bool motion = KS ? MOTION : (a.motion != 0);
    const int i0 = motion ? an : 0, i1 = motion ? an + 1 : ks;
    const float kf = a.kf;
    const int bands = (H + R - 1) / R;
    const int warp = tid >> 5, lane = tid & 31;
    auto blend = [&](uint32_t g, float s, uint32_t d) -> uint32_t {
        const int bg = min(max(__float2int_rn(s), 0), 255);
        return (uint32_t)(int)__dadd_rn(__dmul_rn((double)g, om_tab[d]), __dmul_rn((double)bg, m_tab[d]));   // :467
    };
    for (int bnd = tile; bnd < bands; bnd += a.tiles) {
        const int y0 = bnd * R, rows = min(R, H - y0);
        __syncthreads();                                   // previous band's readers are done (and the tables are written)
        for (int tr = warp; tr < th; tr += ET / 32) {
            const int yy = KS ? reflect_once(y0 + tr - an, H) : reflect101(y0 + tr - an, H);
            const uint16_t *src = limg + (size_t)yy * W;
            float *dst = tilebuf + tr * stride;
            // labels come from L2: four independent loads in flight per lane before the first LUT look-up
            int tc = lane;
            for (; tc + 96 < tw; tc += 128) {
                uint32_t l[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int xx = KS ? reflect_once(tc + 32 * u - an, W) : reflect101(tc + 32 * u - an, W);
                    l[u] = __ldg(src + xx);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) dst[tc + 32 * u] = l[u] ? (float)gray_lut[l[u]] : 0.0f;
            }
            for (; tc < tw; tc += 32) {
                const int xx = KS ? reflect_once(tc - an, W) : reflect101(tc - an, W);
                const uint32_t l = __ldg(src + xx);
                dst[tc] = l ? (float)gray_lut[l] : 0.0f;
            }
        }
        __syncthreads();
        if (KS) {
            constexpr int KSc = KS ? KS : 1, WIN = KSc + 3, NV = (WIN + 3) / 4;
            const int gw = W >> 2;
            for (int ry = warp; ry < rows; ry += ET / 32) {
                const size_t rowpx = (size_t)(y0 + ry) * W;
                const uint32_t *drow = reinterpret_cast<const uint32_t *>(dil + rowpx);
                const uint32_t *arow = reinterpret_cast<const uint32_t *>(alpha + rowpx);
                uint32_t d4n = lane < gw ? __ldg(drow + lane) : 0u, a4n = lane < gw ? __ldg(arow + lane) : 0u;
                for (int xg = lane; xg < gw; xg += 32) {
                    const int x = xg * 4;
                    const uint32_t d4 = d4n, a4 = a4n;          // the next group's weights are in flight during this one
                    if (xg + 32 < gw) { d4n = __ldg(drow + xg + 32); a4n = __ldg(arow + xg + 32); }
                    float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    float c[4];
                    const int ib = d4 ? (MOTION ? KSc / 2 : 0) : KSc / 2, ie = d4 ? (MOTION ? KSc / 2 + 1 : KSc) : KSc / 2 + 1;
                    for (int i = ib; i < ie; ++i) {
                        float win[NV * 4];
                        const float4 *row = reinterpret_cast<const float4 *>(tilebuf + (ry + i) * stride + x);
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const float4 t = row[v];
                            win[4 * v] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
                        }
                        if (i == KSc / 2) {
#pragma unroll
                            for (int p = 0; p < 4; ++p) c[p] = win[p + KSc / 2];
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p)
#pragma unroll
                            for (int j = 0; j < KSc; ++j) s[p] = __fmaf_rn(kf, win[p + j], s[p]);
                    }
                    uint32_t o[2];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const uint32_t d = (d4 >> (8 * p)) & 0xffu;
                        uint32_t g = (uint32_t)c[p];
                        if (d) g = blend(g, s[p], d);
                        const uint32_t pix = g | (((a4 >> (8 * p)) & 0xffu) << 8);
                        if (p & 1) o[p >> 1] |= pix << 16; else o[p >> 1] = pix;
                    }
                    __stcs(reinterpret_cast<uint2 *>(out + (rowpx + x) * 2), make_uint2(o[0], o[1]));
                }
            }
        } else {
            for (int ry = warp; ry < rows; ry += ET / 32) {
                const size_t rowpx = (size_t)(y0 + ry) * W;
                for (int x = lane; x < W; x += 32) {
                    const uint32_t d = dil[rowpx + x];
                    uint32_t g = (uint32_t)tilebuf[(ry + an) * stride + x + an];
                    if (d) {
                        float sg = 0.0f;
                        for (int i = i0; i < i1; ++i) {
                            const float *row = tilebuf + (ry + i) * stride + x;
                            for (int j = 0; j < ks; ++j) sg = __fmaf_rn(kf, row[j], sg);
                        }
                        g = blend(g, sg, d);
                    }
                    reinterpret_cast<uchar2 *>(out)[rowpx + x] = make_uchar2((unsigned char)g, alpha[rowpx + x]);
                }
            }
        }
    }
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// the blur kernel's coefficient exactly as numpy builds it (core/pouch.py:506-513)
static float blur_coefficient(int ks, int motion)
{
    if (motion) return (float)(1.0 / (double)ks);          // float64 quotient assigned into a float32 array
    return 1.0f / (float)(ks * ks);                        // float32 array / int -> float32 division
}

}  // namespace cab200

using namespace cab200;

extern "C" int cab200_edge_maps(int n_cells, const int32_t *poly_off, const int32_t *pts, const uint16_t *label_img,
                                int H, int W, int ksize, int blur_type, uint8_t *edges_out, uint8_t *dilated_out,
                                uint8_t *alpha_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_cells <= 0 || H <= 0 || W <= 0 || !poly_off || !pts || !edges_out || (alpha_out && (!label_img || !dilated_out))) {
        set_error("cab200_edge_maps: bad argument");
        return CAB200_EINVAL;
    }
    if (dilated_out && (ksize < 1 || ksize > EDGE_KMAX)) {
        set_error("cab200_edge_maps: blur kernel size %d outside 1..%d", ksize, EDGE_KMAX);
        return CAB200_EINVAL;
    }
    if ((size_t)H * W >= (1u << 30)) { set_error("cab200_edge_maps: image too large"); return CAB200_ESIZE; }
    CAB_CUDA(cudaMemsetAsync(edges_out, 0, (size_t)H * W, stream));
    edge_lines_kernel<<<n_cells, 64, 0, stream>>>(n_cells, poly_off, pts, H, W, edges_out);
    CAB_CUDA(cudaGetLastError());
    if (dilated_out) {
        const int motion = blur_type == CAB200_BLUR_MOTION;
        edge_dilate_kernel<<<(H * W + 255) / 256, 256, 0, stream>>>(edges_out, label_img, H, W, ksize, motion,
                                                                   blur_coefficient(ksize, motion), dilated_out, alpha_out);
        CAB_CUDA(cudaGetLastError());
    }
    return CAB200_OK;
}

namespace {
typedef void (*EdgeKernel)(const EdgeArgs);
EdgeKernel pick_edge_kernel(int ks, bool motion, bool fast)
{
    if (fast) {
        switch (ks * 2 + (motion ? 1 : 0)) {
        case 6: return raster_edges_kernel<3, false>;
        case 7: return raster_edges_kernel<3, true>;
        case 10: return raster_edges_kernel<5, false>;
        case 11: return raster_edges_kernel<5, true>;
        case 14: return raster_edges_kernel<7, false>;
        case 15: return raster_edges_kernel<7, true>;
        case 18: return raster_edges_kernel<9, false>;
        case 19: return raster_edges_kernel<9, true>;
        default: break;
        }
    }
    return raster_edges_kernel<0, false>;
}
}  // namespace

extern "C" int cab200_raster_edges_batch(
    int n_pouch, const int32_t *geom_id, const int32_t *g_ncell, const int64_t *cell_off,
    const uint16_t *label_img, const uint8_t *edges, const uint8_t *dilated, const uint8_t *alpha_map,
    int H, int W, int n_frames, const double *ca_frames, int mode, int ksize, int blur_type, uint8_t *image_out,
    void *scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pouch < 0 || H <= 0 || W <= 0 || n_frames < 0) {
        set_error("cab200_raster_edges_batch: bad sizes");
        return CAB200_EINVAL;
    }
    if (mode != CAB200_EDGE_BLUR && mode != CAB200_EDGE_BLUR_BORDER && mode != CAB200_EDGE_BORDER) {
        set_error("cab200_raster_edges_batch: unknown mode %d", mode);
        return CAB200_EINVAL;
    }
    if (n_pouch == 0 || n_frames == 0) return CAB200_OK;
    if (!geom_id || !g_ncell || !cell_off || !label_img || !edges || !ca_frames || !image_out || !scratch ||
        (mode == CAB200_EDGE_BLUR && (!dilated || !alpha_map))) {
        set_error("cab200_raster_edges_batch: null pointer argument");
        return CAB200_EINVAL;
    }
    if (mode == CAB200_EDGE_BLUR && (ksize < 1 || ksize > EDGE_KMAX)) {
        set_error("cab200_raster_edges_batch: blur kernel size %d outside 1..%d", ksize, EDGE_KMAX);
        return CAB200_EINVAL;
    }
    if (scratch_bytes < sizeof(PouchEntry) * (size_t)n_pouch) {
        set_error("cab200_raster_edges_batch: scratch too small");
        return CAB200_ESIZE;
    }
    if ((size_t)H * W >= (1u << 30)) { set_error("cab200_raster_edges_batch: image too large"); return CAB200_ESIZE; }
    std::vector<PouchEntry> tab(n_pouch);
    int n_max = 0;
    for (int p = 0; p < n_pouch; ++p) {
        tab[p].cell_off = cell_off[p];
        tab[p].geom = geom_id[p];
        tab[p].n = g_ncell[geom_id[p]];
        if (tab[p].n > 65534 || tab[p].n <= 0 || cell_off[p + 1] - cell_off[p] != tab[p].n) {
            set_error("cab200_raster_edges_batch: pouch %d: bad cell count", p);
            return CAB200_EINVAL;
        }
        n_max = std::max(n_max, tab[p].n);
    }
    CAB_CUDA(cudaMemcpyAsync(scratch, tab.data(), sizeof(PouchEntry) * (size_t)n_pouch, cudaMemcpyHostToDevice, stream));
    EdgeArgs a;
    a.pouches = (const PouchEntry *)scratch;
    a.label_img = label_img; a.edges = edges; a.dilated = dilated; a.alpha = alpha_map; a.ca_frames = ca_frames;
    a.n_frames = n_frames; a.H = H; a.W = W;
    a.mode = mode; a.ks = ksize; a.motion = blur_type == CAB200_BLUR_MOTION;
    a.kf = blur_coefficient(ksize, a.motion);
    a.image_out = image_out;
    size_t smem = align16(frame_lut_bytes(n_max));
    a.band = 1;
    a.stride = 0;
    bool fast = false;
    if (mode == CAB200_EDGE_BLUR) {
        // tile of (band + ks - 1) rows of gray as float, W + ks - 1 wide (padded for 16-byte loads):
        // as many rows as ~48 KB hold (three CTAs per SM)
        fast = (W % 4 == 0) && (ksize % 2 == 1) && ksize >= 3 && ksize <= 9 && ksize <= std::min(H, W);
        a.stride = ((W + ksize - 1 + 3) & ~3) + 4;
        const size_t row_bytes = (size_t)a.stride * sizeof(float);
        long long rows_total = (long long)((48 * 1024) / row_bytes);
        rows_total = std::max<long long>(rows_total, ksize);
        a.band = (int)std::min<long long>(rows_total - (ksize - 1), H);
        smem += 2 * 256 * sizeof(double) + (size_t)(a.band + ksize - 1) * row_bytes;
    }
    if (smem > 227 * 1024) { set_error("cab200_raster_edges_batch: image too wide for one shared-memory tile"); return CAB200_ESIZE; }
    const long long pf = (long long)n_pouch * n_frames;
    const int bands = (H + a.band - 1) / a.band;
    int tiles = 1;
    while ((long long)tiles * pf < 148 * 3 && tiles * 2 <= std::max(bands, 1) && tiles < 64) tiles *= 2;
    a.tiles = tiles;
    const EdgeKernel kern = pick_edge_kernel(ksize, a.motion != 0, fast);
    if (smem > 48 * 1024)
        CAB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (pf * tiles > 0x7fffffffLL) { set_error("cab200_raster_edges_batch: too many frames"); return CAB200_ESIZE; }
    kern<<<(unsigned)(pf * tiles), ET, smem, stream>>>(a);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}
