// integrate_f32.cu -- one quarter of K1's template instantiations (see integrate_pick.cuh).
#include "integrate_pick.cuh"

namespace cab200 {

KernelFn pick_kernel_f32(int threads, int cpt, int ep, bool unit, int copies)
{
    return pick_kernel<float>(threads, cpt, ep, unit, copies);
}

}  // namespace cab200
