// defects.cu -- K5: the deterministic part of the reference's defect pipeline (apply_all_defects,
// utils/image_processing.py:130-175) on (H, W, 2) uint8 frames -- SURVEY.md section 8 row f3, partial.
//
// Between two defects the reference's image is uint8 again (every function ends in .astype(image.dtype)), and
// each supported defect is either a function of the pixel value and of one global statistic of the frame
// (np.max(image), or the test image.max() > 1.0) or the product with a static per-pixel factor:
//   uniform background fluorescence   artifacts/background.py:46-56     value, max
//   dynamic range compression (gamma) artifacts/sensor.py:117-135       value, max > 1
//   quantization                      artifacts/sensor.py:138-156       value, max
//   brightness / contrast             utils/image_processing.py:97-127  value, max > 1
//   vignetting                        artifacts/optical.py:49-83        value x vignette[y][x]   (float64, truncated)
// so a defect is one pass over the frames: look the pixel up in a 256-entry row chosen by the frame's statistic
// (the host builds the rows with the reference's own numpy expressions, calcium_b200/defects.py) or multiply by
// the static float64 map, and reduce the new maximum for the next pass on the way.  HBM bound: 2 bytes read
// and 2 written per pixel and pass.
#include <algorithm>

#include "common.cuh"

namespace cab200 {

constexpr int DT = 256;

__global__ void __launch_bounds__(DT) defect_pass_kernel(uint8_t *image, size_t bytes_per_frame, int W, int kind,
                                                         const uint8_t *table, const double *vig,
                                                         const uint32_t *max_in, uint32_t *max_out)
{
    __shared__ uint8_t row[256];
    const int frame = blockIdx.y, tid = threadIdx.x;
    uint8_t *img = image + (size_t)frame * bytes_per_frame;
    if (kind == CAB200_DEFECT_LUT_MAX || kind == CAB200_DEFECT_LUT_FLAG) {
        const uint32_t m = max_in[frame];
        const uint32_t r = (kind == CAB200_DEFECT_LUT_MAX) ? min(m, 255u) : (m > 1u ? 1u : 0u);
        row[tid] = table[r * 256u + tid];
        __syncthreads();
    }
    uint32_t mx = 0;
    const size_t n16 = bytes_per_frame / 16;
    for (size_t g = (size_t)blockIdx.x * DT + tid; g < n16; g += (size_t)gridDim.x * DT) {
        uint4 v = reinterpret_cast<const uint4 *>(img)[g];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (kind == CAB200_DEFECT_LUT_MAX || kind == CAB200_DEFECT_LUT_FLAG) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                w[q] = (uint32_t)row[w[q] & 0xffu] | ((uint32_t)row[(w[q] >> 8) & 0xffu] << 8) |
                       ((uint32_t)row[(w[q] >> 16) & 0xffu] << 16) | ((uint32_t)row[w[q] >> 24] << 24);
        } else if (kind == CAB200_DEFECT_MUL) {
            const size_t px0 = g * 8;                      // 8 pixels of 2 bytes; both channels share the factor
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t o = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double f = __ldg(vig + px0 + 2 * q + h);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t b = (w[q] >> (16 * h + 8 * c)) & 0xffu;
                        // image[:, :, i] * vignette assigned into the uint8 array: float64 product, C cast
                        const uint32_t nb = (uint32_t)(unsigned char)(long long)__dmul_rn((double)b, f);
                        o |= nb << (16 * h + 8 * c);
                    }
                }
                w[q] = o;
            }
        }
        if (kind != CAB200_DEFECT_MAX_ONLY)
            reinterpret_cast<uint4 *>(img)[g] = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mx = max(mx, w[q] & 0xffu); mx = max(mx, (w[q] >> 8) & 0xffu);
            mx = max(mx, (w[q] >> 16) & 0xffu); mx = max(mx, w[q] >> 24);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0 && mx) atomicMax(max_out + frame, mx);
}

}  // namespace cab200

using namespace cab200;

extern "C" int cab200_defect_pass(uint8_t *image, int n_frames, int H, int W, int kind, const uint8_t *table,
                                  const double *vignette, const uint32_t *max_in, uint32_t *max_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_frames < 0 || H <= 0 || W <= 0 || kind < CAB200_DEFECT_MAX_ONLY || kind > CAB200_DEFECT_MUL) {
        set_error("cab200_defect_pass: bad argument");
        return CAB200_EINVAL;
    }
    if (n_frames == 0) return CAB200_OK;
    if (!image || !max_out || ((kind == CAB200_DEFECT_LUT_MAX || kind == CAB200_DEFECT_LUT_FLAG) && (!table || !max_in)) ||
        (kind == CAB200_DEFECT_MUL && !vignette)) {
        set_error("cab200_defect_pass: null pointer argument");
        return CAB200_EINVAL;
    }
    const size_t bytes = (size_t)H * W * 2;
    if (bytes % 16) { set_error("cab200_defect_pass: H*W must be a multiple of 8"); return CAB200_EINVAL; }
    if (n_frames > 65535) { set_error("cab200_defect_pass: at most 65535 frames per call"); return CAB200_ESIZE; }
    CAB_CUDA(cudaMemsetAsync(max_out, 0, sizeof(uint32_t) * (size_t)n_frames, stream));
    const size_t n16 = bytes / 16;
    int tiles = (int)std::min<size_t>((n16 + DT - 1) / DT, 16);
    // enough CTAs to fill the machine when there are few frames
    while ((long long)tiles * n_frames < 148 * 8 && (size_t)tiles * DT < n16) tiles *= 2;
    defect_pass_kernel<<<dim3((unsigned)tiles, (unsigned)n_frames), DT, 0, stream>>>(
        image, bytes, W, kind, table, vignette, max_in, max_out);
    CAB_CUDA(cudaGetLastError());
    return CAB200_OK;
}
