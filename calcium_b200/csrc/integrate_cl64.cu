// integrate_cl64.cu -- one quarter of K1's template instantiations (see integrate_pick.cuh).
#include "integrate_pick.cuh"

namespace cab200 {

KernelFn pick_cluster_kernel_f64(int threads, int cpt, int ncta, int ep, int copies)
{
    return pick_cluster_kernel<double>(threads, cpt, ncta, ep, copies);
}

}  // namespace cab200
