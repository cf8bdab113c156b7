// edt.cu -- K4: exact squared Euclidean distance to the nearest pixel of a different class.
//
// Replaces the distance transforms of the reference's prompt generation:
//   generate_point_prompts_from_mask   utils/sam2_data_generator.py:71-94
//       distance_transform_edt(instance_mask == cell_id) for every instance (n_active full-image
//       transforms per frame in the reference)
//   generate_negative_prompts          utils/sam2_data_generator.py:152-159
//       distance_transform_edt(1 - (instance_mask > 0))
// Both are "distance from a pixel to the nearest pixel whose class differs from its own", with class =
// instance id (all per-instance transforms at once) or class = (instance id > 0).  scipy returns
// sqrt of the exact integer squared distance, which is what this kernel produces (the caller takes
// the correctly rounded double sqrt).  Pixels outside the image do not exist (no border effect),
// exactly as in scipy.
//
// Separable and exact.  Let g(y, x) be the vertical distance from (y, x) to the nearest pixel of its
// column whose class differs from class(y, x) (infinite if none).  For a query pixel p = (y, x) of
// class C, the nearest different-class pixel in column x' is at vertical distance 0 if
// class(y, x') != C and g(y, x') otherwise -- g is relative to class(y, x'), which then IS C.  Hence
//   d2(p) = min over x' of (x - x')^2 + [class(y, x') == C] * g(y, x')^2 .
// Pass 1 computes g with one thread per column; pass 2 scans each row outward from every query pixel
// and stops once (x - x')^2 alone exceeds the best value found (one CTA per row, row staged in
// shared memory).
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace cab200 {

constexpr uint16_t EDT_INF = 0xffff;

// class of pixel: lut ? lut[label] : label
__device__ __forceinline__ uint32_t edt_class(const uint16_t *lut, uint32_t label) { return lut ? (uint32_t)lut[label] : label; }

// grid (ceil(W/128), B): one thread per column
__global__ void __launch_bounds__(128) edt_columns_kernel(const uint16_t *label, const uint16_t *lut, int lut_stride, int H,
                                                          int W, uint16_t *g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int b = blockIdx.y;
    const uint16_t *l = lut ? lut + (size_t)b * lut_stride : nullptr;
    uint16_t *gb = g + (size_t)b * H * W;
    // downward: distance to the nearest different class above
    uint32_t prev = 0xffffffffu, run_start = 0;
    bool open_top = true;                       // the current run reaches the top edge: nothing different above
    for (int y = 0; y < H; ++y) {
        const uint32_t c = edt_class(l, label[(size_t)y * W + x]);
        if (y > 0 && c != prev) { run_start = (uint32_t)y; open_top = false; }
        gb[(size_t)y * W + x] = open_top ? EDT_INF : (uint16_t)min((uint32_t)y - run_start + 1u, 0xfffeu);
        prev = c;
    }
    // upward: distance to the nearest different class below, merged
    uint32_t run_end = (uint32_t)H - 1u;
    bool open_bottom = true;
    for (int y = H - 1; y >= 0; --y) {
        const uint32_t c = edt_class(l, label[(size_t)y * W + x]);
        if (y < H - 1 && c != prev) { run_end = (uint32_t)y; open_bottom = false; }
        if (!open_bottom) {
            const uint32_t d = min(run_end - (uint32_t)y + 1u, 0xfffeu);
            uint16_t *p = gb + (size_t)y * W + x;
            if (d < *p) *p = (uint16_t)d;
        }
        prev = c;
    }
}

// grid (H, B): one CTA per row; d2_out int32, INT_MAX where the whole image is one class
__global__ void __launch_bounds__(256) edt_rows_kernel(const uint16_t *label, const uint16_t *lut, int lut_stride, int H, int W,
                                                       const uint16_t *g, int32_t *d2_out)
{
    extern __shared__ uint32_t row[];           // class | g << 16
    const int y = blockIdx.x, b = blockIdx.y;
    const uint16_t *l = lut ? lut + (size_t)b * lut_stride : nullptr;
    const uint16_t *gr = g + ((size_t)b * H + y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x)
        row[x] = edt_class(l, label[(size_t)y * W + x]) | ((uint32_t)gr[x] << 16);
    __syncthreads();
    int32_t *out = d2_out + ((size_t)b * H + y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const uint32_t me = row[x] & 0xffffu;
        const uint32_t g0 = row[x] >> 16;
        long long best = (g0 == EDT_INF) ? (long long)INT_MAX : (long long)g0 * g0;
        for (int k = 1; (long long)k * k < best; ++k) {
            const long long k2 = (long long)k * k;
            bool any = false;
            if (x - k >= 0) {
                any = true;
                const uint32_t r = row[x - k];
                const uint32_t gg = r >> 16;
                const long long c = ((r & 0xffffu) != me) ? k2 : (gg == EDT_INF ? (long long)INT_MAX : k2 + (long long)gg * gg);
                best = min(best, c);
            }
            if (x + k < W) {
                any = true;
                const uint32_t r = row[x + k];
                const uint32_t gg = r >> 16;
                const long long c = ((r & 0xffffu) != me) ? k2 : (gg == EDT_INF ? (long long)INT_MAX : k2 + (long long)gg * gg);
                best = min(best, c);
            }
            if (!any) break;
        }
        out[x] = (int32_t)min(best, (long long)INT_MAX);
    }
}

}  // namespace cab200

using namespace cab200;

extern "C" size_t cab200_edt_scratch_bytes(int n_maps, int H, int W) { return (size_t)std::max(n_maps, 0) * H * W * sizeof(uint16_t); }

extern "C" int cab200_edt_classes(const uint16_t *label_map, const uint16_t *class_lut, int lut_stride, int n_maps, int H, int W,
                                  int32_t *d2_out, void *scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_maps < 0 || H <= 0 || W <= 0 || H > 65534 || W > 65534 || (class_lut && lut_stride <= 0)) {
        set_error("cab200_edt_classes: bad sizes");
        return CAB200_EINVAL;
    }
    if (n_maps == 0) return CAB200_OK;
    if (!label_map || !d2_out || !scratch) { set_error("cab200_edt_classes: null pointer argument"); return CAB200_EINVAL; }
    if (scratch_bytes < cab200_edt_scratch_bytes(n_maps, H, W)) { set_error("cab200_edt_classes: scratch too small"); return CAB200_ESIZE; }
    if (n_maps > 65535) { set_error("cab200_edt_classes: at most 65535 maps per call"); return CAB200_ESIZE; }
    if ((size_t)W * sizeof(uint32_t) > 200 * 1024) { set_error("cab200_edt_classes: row too long for shared memory"); return CAB200_ESIZE; }
    uint16_t *g = (uint16_t *)scratch;
    edt_columns_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)n_maps), 128, 0, stream>>>(label_map, class_lut, lut_stride, H, W, g);
    CAB_CUDA(cudaGetLastError());
    const size_t smem = (size_t)W * sizeof(uint32_t);
    if (smem > 48 * 1024)
        CAB_CUDA(cudaFuncSetAttribute((const void *)edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edt_rows_kernel<<<dim3((unsigned)H, (unsigned)n_maps), 256, smem, stream>>>(label_map, class_lut, lut_stride, H, W, g, d2_out);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}
