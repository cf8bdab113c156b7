// api.cu -- error channel, version/device queries and the roofline micro-kernels of libcalcium_b200.
#include <stdarg.h>

#include "common.cuh"

namespace cab200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- FP64 pipe peak: 8 independent DFMA chains per thread, no memory traffic ---------------------
__global__ void __launch_bounds__(256) dfma_kernel(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;   // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) fill_kernel(uint4 *dst, size_t n16, uint32_t v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        __stcs(dst + i, make_uint4(v, v + 1, v + 2, v + 3));
}

// ---- self-test of the branch-free division used by K1 -------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D2049BB133111Bull;
    return z ^ (z >> 31);
}

// Operands: random mantissas, exponents uniform in [-exp_range, exp_range]; every 16th numerator 0.
__global__ void __launch_bounds__(256) divtest_kernel(uint64_t seed, int per_thread, int exp_range,
                                                      unsigned long long *mismatch)
{
    uint64_t st = seed + 0x1234567ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    unsigned long long bad = 0;
    for (int i = 0; i < per_thread; ++i) {
        const uint64_t ra = splitmix64(st), rb = splitmix64(st), re = splitmix64(st);
        const int ea = (int)(re % (2 * exp_range + 1)) - exp_range;
        const int eb = (int)((re >> 32) % (2 * exp_range + 1)) - exp_range;
        double a = __longlong_as_double((ra & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 + ea) << 52));
        const double b = __longlong_as_double((rb & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 + eb) << 52));
        if ((i & 15) == 0) a = 0.0;
        const double ref = __ddiv_rn(a, b);
        const double got = Strict<double>::fdiv(a, b);
        const Strict<double>::Recip rc = Strict<double>::recip(b);
        const double got2 = Strict<double>::div(a, rc);
        bad += (__double_as_longlong(ref) != __double_as_longlong(got)) ||
               (__double_as_longlong(ref) != __double_as_longlong(got2));
    }
    if (bad) atomicAdd(mismatch, bad);
}

}  // namespace cab200

using namespace cab200;

extern "C" int cab200_selftest_division(long long n_pairs, int exp_range, unsigned long long seed,
                                        unsigned long long *mismatches_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs <= 0 || exp_range < 0 || exp_range > 900 || !mismatches_out) {
        set_error("cab200_selftest_division: bad argument");
        return CAB200_EINVAL;
    }
    unsigned long long *d = nullptr;
    CAB_CUDA(cudaMalloc(&d, sizeof(*d)));
    const int blocks = 148 * 8, threads = 256;
    const int per_thread = (int)((n_pairs + (long long)blocks * threads - 1) / ((long long)blocks * threads));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(*d), stream);
    if (e == cudaSuccess) {
        divtest_kernel<<<blocks, threads, 0, stream>>>(seed, per_thread, exp_range, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(mismatches_out, d, sizeof(*d), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);                                   // on every path
    if (e != cudaSuccess) { set_error("cab200_selftest_division: %s", cudaGetErrorString(e)); return CAB200_ECUDA; }
    return CAB200_OK;
}

extern "C" int cab200_version(void) { return CAB200_VERSION; }
extern "C" const char *cab200_last_error(void) { return g_err; }

extern "C" int cab200_device_info(int device, int64_t out[4])
{
    if (!out) { set_error("cab200_device_info: null out"); return CAB200_EINVAL; }
    cudaDeviceProp prop;
    CAB_CUDA(cudaGetDeviceProperties(&prop, device));
    out[0] = prop.multiProcessorCount;
    out[1] = (int64_t)prop.sharedMemPerBlockOptin;
    out[2] = prop.major * 10 + prop.minor;
    out[3] = prop.l2CacheSize;
    return CAB200_OK;
}

extern "C" int cab200_peak_fp64(int iters, double *tflops_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (iters <= 0 || !tflops_out) { set_error("cab200_peak_fp64: bad argument"); return CAB200_EINVAL; }
    int dev = 0, sms = 0;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    CAB_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CAB_CUDA(cudaEventCreate(&e0));
    CAB_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    dfma_kernel<<<blocks, 256, 0, stream>>>(iters / 8 + 1, 1.0, sink);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CAB_CUDA(cudaEventRecord(e0, stream));
        dfma_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0, sink);
        CAB_CUDA(cudaEventRecord(e1, stream));
        CAB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CAB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
        best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops_out = best;
    return CAB200_OK;
}

extern "C" int cab200_peak_hbm_write(void *buf, size_t bytes, int reps, double *gbs_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!buf || bytes < 16 || reps <= 0 || !gbs_out) { set_error("cab200_peak_hbm_write: bad argument"); return CAB200_EINVAL; }
    int dev = 0, sms = 0;
    CAB_CUDA(cudaGetDevice(&dev));
    CAB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1;
    CAB_CUDA(cudaEventCreate(&e0));
    CAB_CUDA(cudaEventCreate(&e1));
    const size_t n16 = bytes / 16;
    fill_kernel<<<sms * 8, 256, 0, stream>>>((uint4 *)buf, n16, 0u);
    double best = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
        CAB_CUDA(cudaEventRecord(e0, stream));
        fill_kernel<<<sms * 8, 256, 0, stream>>>((uint4 *)buf, n16, (uint32_t)rep);
        CAB_CUDA(cudaEventRecord(e1, stream));
        CAB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CAB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = fmax(best, (double)(n16 * 16) / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gbs_out = best;
    return CAB200_OK;
}
