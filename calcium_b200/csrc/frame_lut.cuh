// frame_lut.cuh -- the per-frame look-up tables shared by K2 (raster.cu) and K3 (encode.cu).
//
// For one (pouch, frame) a CTA of NT threads reduces the calcium values of the n cells to three
// tables indexed by label value v in [0, n] (0 = background, v = cell index + 1):
//   gray_lut[v]  int(255 * clip(ca / vmax, 0, 1)), vmax = max(max(ca), 0.5)    core/pouch.py:421-427
//   act8[v]      ca > threshold (strict)                                       core/pouch.py:563-565
//   inst_lut[v]  instance id, including the reference's 0-based/1-based id mismatch when `quirk`
//                (utils/sam2_data_generator.py:31-35 applied to get_cell_masks(active_only=True))
// rank1[] is scratch of the same size as inst_lut.  Returns the number of active cells.
#pragma once
#include "common.cuh"

namespace cab200 {

struct PouchEntry {
    int64_t cell_off;
    int32_t n;
    int32_t geom;
};

// bytes of dynamic shared memory the four tables take for n cells; offsets via frame_lut_carve
__host__ __device__ static inline size_t frame_lut_bytes(int n)
{
    return 2 * (size_t)((n + 2 + 7) & ~7) * 2 + 2 * (size_t)((n + 2 + 15) & ~15);
}

struct FrameLut {
    uint16_t *inst_lut, *rank1;
    uint8_t *gray_lut, *act8;
};

__device__ __forceinline__ FrameLut frame_lut_carve(unsigned char *smem, int n)
{
    FrameLut t;
    t.inst_lut = reinterpret_cast<uint16_t *>(smem);
    t.rank1 = t.inst_lut + ((n + 2 + 7) & ~7);
    t.gray_lut = reinterpret_cast<uint8_t *>(t.rank1 + ((n + 2 + 7) & ~7));
    t.act8 = t.gray_lut + ((n + 2 + 15) & ~15);
    return t;
}

template <int NT>
__device__ __forceinline__ int frame_lut_build(const FrameLut t, const double *ca, int n, double threshold, int quirk)
{
    __shared__ double red[NT / 32];
    __shared__ int wsum[NT / 32];
    __shared__ double s_vmax;
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- vmax = max(np.max(ca), 0.5)  (core/pouch.py:421-422) --------------------------------
    double mx = -INFINITY;
    for (int i = tid; i < n; i += NT) mx = fmax(mx, ca[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        double m = red[0];
        for (int w = 1; w < NT / 32; ++w) m = fmax(m, red[w]);
        s_vmax = (0.5 > m) ? 0.5 : m;
    }
    __syncthreads();
    const double vmax = s_vmax;

    // ---- per-cell gray level, active flag, 1-based rank among active cells ---------------------
    // contiguous chunk per thread so the scan keeps ascending cell order (np.where, :564)
    const int chunk = (n + NT - 1) / NT;
    const int beg = min(tid * chunk, n), end = min(beg + chunk, n);
    int cnt = 0;
    for (int i = beg; i < end; ++i) {
        const double v = ca[i];
        double nv = __ddiv_rn(__dsub_rn(v, 0.0), __dsub_rn(vmax, 0.0));        // :426
        nv = (nv < 0.0) ? 0.0 : ((nv > 1.0) ? 1.0 : nv);                       // np.clip
        t.gray_lut[i + 1] = (uint8_t)(int)__dmul_rn(255.0, nv);                // int(255*x), :427
        const int act = v > threshold;                                          // strict >, :564
        t.act8[i + 1] = (uint8_t)act;
        cnt += act;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < NT / 32; ++w) { const int u = wsum[w]; wsum[w] = run; run += u; }
        s_total = run;
        t.gray_lut[0] = 0; t.act8[0] = 0;
    }
    __syncthreads();
    int run = wsum[warp] + incl - cnt;
    for (int i = beg; i < end; ++i) {
        const int act = t.act8[i + 1];
        run += act;
        t.rank1[i] = (uint16_t)(act ? run : 0);      // rank1[cell], 1-based
    }
    if (tid == 0) t.rank1[n] = 0;
    __syncthreads();
    // instance id per label value
    for (int v = tid; v <= n; v += NT) {
        uint16_t id;
        if (quirk) {
            // reference semantics: masked value m = v if cell v-1 active else 0; then the loop
            // `instance[mask == cell_id] = new_id` over 0-based active ids (D5)
            const int m = (v > 0 && t.act8[v]) ? v : 0;
            id = (m < n && t.act8[m + 1]) ? t.rank1[m] : 0;
        } else {
            id = (v > 0 && t.act8[v]) ? t.rank1[v - 1] : 0;
        }
        t.inst_lut[v] = id;
    }
    const int total = s_total;
    __syncthreads();
    return total;
}

}  // namespace cab200
