// integrate_cl32.cu -- one quarter of K1's template instantiations (see integrate_pick.cuh).
#include "integrate_pick.cuh"

namespace cab200 {

KernelFn pick_cluster_kernel_f32(int threads, int cpt, int ncta, int ep, int copies)
{
    return pick_cluster_kernel<float>(threads, cpt, ncta, ep, copies);
}

}  // namespace cab200
