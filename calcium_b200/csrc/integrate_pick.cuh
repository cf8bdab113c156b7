// integrate_pick.cuh -- kernel selection tables for K1.  The template instantiations are spread over
// four translation units (FP64 / FP32, single CTA / cluster) so they compile in parallel.
#pragma once
#include "integrate.cuh"

namespace cab200 {

typedef void (*KernelFn)(const SimArgs);

template <typename R, int THREADS, int CPT, int NCTA>
static KernelFn pick_cluster_variant(int ep, int copies)
{
    if (copies == 2) {
        if (ep <= 5) return integrate_kernel<R, THREADS, CPT, 5, true, 2, 1, NCTA>;
        return integrate_kernel<R, THREADS, CPT, 8, true, 2, 1, NCTA>;
    }
    if (ep <= 5) return integrate_kernel<R, THREADS, CPT, 5, true, 1, 1, NCTA>;
    return integrate_kernel<R, THREADS, CPT, 8, true, 1, 1, NCTA>;
}

// Cluster kernels exist for four CTA shapes and cluster sizes 2 and 4, and for the 416- / 512-thread shapes as
// clusters of 8 (the portable maximum: pouches of up to 8 x 1536 = 12 288 cells, the reference's `xlarge`).
template <typename R>
static KernelFn pick_cluster_kernel(int threads, int cpt, int ncta, int ep, int copies)
{
#define CAB_CSHAPE(T_, C_) \
    if (threads == T_ && cpt == C_) { \
        if (ncta == 2) return pick_cluster_variant<R, T_, C_, 2>(ep, copies); \
        if (ncta == 4) return pick_cluster_variant<R, T_, C_, 4>(ep, copies); \
    }
    CAB_CSHAPE(256, 2)
    CAB_CSHAPE(416, 2)
    CAB_CSHAPE(512, 2)
    CAB_CSHAPE(512, 3)
#undef CAB_CSHAPE
    if (ncta == 8 && threads == 416 && cpt == 2) return pick_cluster_variant<R, 416, 2, 8>(ep, copies);
    if (ncta == 3 && threads == 512 && cpt == 3) return pick_cluster_variant<R, 512, 3, 3>(ep, copies);   // see cab200_cluster_capacity
    if (ncta == 8 && threads == 512 && cpt == 2) return pick_cluster_variant<R, 512, 2, 8>(ep, copies);
    if (ncta == 8 && threads == 512 && cpt == 3) return pick_cluster_variant<R, 512, 3, 8>(ep, copies);
    return nullptr;
}

// CTA shape for one part of a clustered pouch.
static inline void cluster_shape(int cells_per_cta, int *threads, int *cpt)
{
    *threads = 0; *cpt = 0;
    if (cells_per_cta <= 512) { *threads = 256; *cpt = 2; }
    else if (cells_per_cta <= 832) { *threads = 416; *cpt = 2; }      // 13 warps: a large pouch on 4 CTAs (814 cells each)
    else if (cells_per_cta <= 1024) { *threads = 512; *cpt = 2; }
    else if (cells_per_cta <= 1536) { *threads = 512; *cpt = 3; }
}

template <typename R, int THREADS, int CPT, int MINB>
static KernelFn pick_variant(int ep, bool unit, int copies)
{
    if (unit) {
        if (copies == 2) {
            if (ep <= 5) return integrate_kernel<R, THREADS, CPT, 5, true, 2, MINB>;
            return integrate_kernel<R, THREADS, CPT, 8, true, 2, MINB>;
        }
        if (ep <= 5) return integrate_kernel<R, THREADS, CPT, 5, true, 1, MINB>;
        return integrate_kernel<R, THREADS, CPT, 8, true, 1, MINB>;
    }
    return integrate_kernel<R, THREADS, CPT, 8, false, 1, MINB>;
}

template <typename R>
static KernelFn pick_kernel(int threads, int cpt, int ep, bool unit, int copies)
{
#define CAB_SHAPE(T_, C_, M_) \
    if (threads == T_ && cpt == C_) return pick_variant<R, T_, C_, M_>(ep, unit, copies);
    CAB_SHAPE(128, 1, 8)
    CAB_SHAPE(256, 1, 4)
    CAB_SHAPE(256, 2, 2)
    CAB_SHAPE(512, 1, 2)
    CAB_SHAPE(512, 2, 1)
    CAB_SHAPE(512, 3, 1)
    CAB_SHAPE(768, 2, 1)
    CAB_SHAPE(1024, 2, 1)
    CAB_SHAPE(1024, 3, 1)
    CAB_SHAPE(1024, 4, 1)
#undef CAB_SHAPE
    return nullptr;
}


KernelFn pick_kernel_f64(int threads, int cpt, int ep, bool unit, int copies);
KernelFn pick_kernel_f32(int threads, int cpt, int ep, bool unit, int copies);
KernelFn pick_cluster_kernel_f64(int threads, int cpt, int ncta, int ep, int copies);
KernelFn pick_cluster_kernel_f32(int threads, int cpt, int ncta, int ep, int copies);

}  // namespace cab200
