// microbench.cu -- measured ceilings of the shared-memory data pipe (the pipe K1's gather lives on).
//
// K1's roofline is the LSU / shared-memory crossbar, for which MEASURED_PEAKS.json has no entry and
// the guides only state the nominal 128 B/clk/SM.  cab200_peak_smem measures it the way K1 uses it
// (LDS.128 / STS.128 from 16 resident warps, several independent accesses in flight per thread) and
// answers two layout questions with numbers instead of folklore:
//   * what does an access cost whose lanes all read ONE address (K1's zero-entry padding)?
//   * do warp shuffles offer a second data path for neighbour values?  (They do not: SHFL moves
//     128 B per instruction through the same crossbar -- compare modes 0 and 5.)
#include <vector>

#include "common.cuh"

namespace cab200 {

enum { MB_LDS128 = 0, MB_LDS128_SAME = 1, MB_LDS128_QUARTER_SAME = 2, MB_LDS128_HALF_ZERO = 3, MB_LDS64 = 4,
       MB_SHFL = 5, MB_STS128 = 6, MB_LDS128_2WAY = 7, MB_N };

template <int MODE>
__global__ void __launch_bounds__(512, 1) smem_kernel(int iters, unsigned long long *cycles, double *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double2 *buf = reinterpret_cast<double2 *>(sm);
    for (int i = tid; i < 4096; i += 512) buf[i] = make_double2(1e-9 * i, 2e-9 * i);
    __syncthreads();
    // per-thread entry index of access u (u = 0..7): chosen per mode
    int idx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int base = ((warp * 8 + u) * 32) & 4095;
        if (MODE == MB_LDS128 || MODE == MB_LDS64 || MODE == MB_STS128) idx[u] = base + lane;      // conflict free
        else if (MODE == MB_LDS128_SAME) idx[u] = base;                                              // one address per warp
        else if (MODE == MB_LDS128_QUARTER_SAME) idx[u] = base + (lane >> 3);                        // one per 8-lane phase
        else if (MODE == MB_LDS128_HALF_ZERO) idx[u] = (lane & 1) ? 4095 : base + lane;              // half the lanes pad
        else if (MODE == MB_LDS128_2WAY)                                                             // 2 entries per bank group
            idx[u] = (base + (lane >> 3) * 16 + (lane & 3) + 8 * ((lane >> 2) & 1)) & 4095;
        else idx[u] = base + lane;
    }
    double acc0 = 0, acc1 = 0;
    unsigned sh[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) sh[u] = (unsigned)lane * 17u + (unsigned)u;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == MB_SHFL) {
#pragma unroll
            for (int u = 0; u < 8; ++u) sh[u] = __shfl_sync(0xffffffffu, sh[u], (lane + u + 1) & 31);
        } else if (MODE == MB_STS128) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned addr = (unsigned)__cvta_generic_to_shared(buf + idx[u]);
                asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(acc0), "d"(acc1) : "memory");
            }
            acc0 += 1.0;
        } else if (MODE == MB_LDS64) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned addr = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double *>(sm) + idx[u]);
                asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(addr) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) { acc0 += v[u]; acc1 += v[u + 1]; }
        } else {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned addr = (unsigned)__cvta_generic_to_shared(buf + idx[u]);
                asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "r"(addr) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 8; u += 4) { acc0 += v[u].x + v[u + 2].x; acc1 += v[u + 1].y + v[u + 3].y; }
        }
    }
    const unsigned long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    unsigned shx = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) shx ^= sh[u];
    if (acc0 + acc1 == 123.456 || shx == 0x12345678u) sink[0] = acc0;
}

typedef void (*SmemFn)(int, unsigned long long *, double *);
static SmemFn smem_fn(int mode)
{
    switch (mode) {
        case MB_LDS128: return smem_kernel<MB_LDS128>;
        case MB_LDS128_SAME: return smem_kernel<MB_LDS128_SAME>;
        case MB_LDS128_QUARTER_SAME: return smem_kernel<MB_LDS128_QUARTER_SAME>;
        case MB_LDS128_HALF_ZERO: return smem_kernel<MB_LDS128_HALF_ZERO>;
        case MB_LDS64: return smem_kernel<MB_LDS64>;
        case MB_SHFL: return smem_kernel<MB_SHFL>;
        case MB_STS128: return smem_kernel<MB_STS128>;
        case MB_LDS128_2WAY: return smem_kernel<MB_LDS128_2WAY>;
    }
    return nullptr;
}

}  // namespace cab200

using namespace cab200;

// See include/cab200.h.
extern "C" int cab200_peak_smem(int mode, int iters, double out[3], void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    SmemFn fn = smem_fn(mode);
    if (!fn || iters <= 0 || !out) { set_error("cab200_peak_smem: bad argument (mode 0..%d)", MB_N - 1); return CAB200_EINVAL; }
    int dev = 0, sms = 0;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long *d_cyc = nullptr;
    double *sink = nullptr;
    CAB_CUDA(cudaMalloc(&d_cyc, sizeof(unsigned long long) * sms));
    if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) { cudaFree(d_cyc); set_error("cab200_peak_smem: cudaMalloc"); return CAB200_ECUDA; }
    const size_t smem = 4096 * 16;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = CAB200_OK;
    float best_ms = 1e30f;
    unsigned long long best_cyc = ~0ull;
    std::vector<unsigned long long> h(sms);
    do {
        if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { rc = CAB200_ECUDA; break; }
        fn<<<sms, 512, smem, stream>>>(iters / 8 + 1, d_cyc, sink);
        for (int rep = 0; rep < 3 && rc == CAB200_OK; ++rep) {
            cudaEventRecord(e0, stream);
            fn<<<sms, 512, smem, stream>>>(iters, d_cyc, sink);
            cudaEventRecord(e1, stream);
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = CAB200_ECUDA; break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            best_ms = fminf(best_ms, ms);
            if (cudaMemcpy(h.data(), d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = CAB200_ECUDA; break; }
            unsigned long long mx = 0;
            for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
            best_cyc = mx < best_cyc ? mx : best_cyc;
        }
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_cyc); cudaFree(sink);
    if (rc != CAB200_OK) { set_error("cab200_peak_smem: CUDA error %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    // bytes a warp-wide access moves: 512 (128-bit), 256 (64-bit), 128 (SHFL, 32-bit)
    const double bytes_per_instr = (mode == MB_LDS64) ? 256.0 : (mode == MB_SHFL) ? 128.0 : 512.0;
    const double instr = (double)iters * 8.0 * 16.0;                 // per SM: 8 accesses x 16 warps per iteration
    out[0] = instr * bytes_per_instr / (double)best_cyc;             // bytes per clk per SM (requested bytes)
    out[1] = instr * bytes_per_instr * sms / (best_ms * 1e-3) / 1e9; // GB/s over the device
    out[2] = (double)best_cyc / instr;                               // clk per warp-wide access per SM
    return CAB200_OK;
}
