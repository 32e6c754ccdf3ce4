// tools/kdev/quad_eh.cu -- compile ONLY the bench kernel (KDEV_MODEL=1, EnergyHydrology) or the Richards quad
// (KDEV_MODEL=0) for SASS / register inspection:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -Xptxas -v -Iclimaland.jl_b200/csrc \
//        -o /tmp/kdev/quad.cubin tools/kdev/quad_eh.cu
// (tools/kdev/phase_mix.py attributes the instructions to source lines / phases)
#include "soil_pair.cuh"
#ifndef KDEV_MODEL
#define KDEV_MODEL 1
#endif
void *kdev_address()
{
    return (void *)clb::k_step_lanes<0, KDEV_MODEL, 15, 2, (KDEV_MODEL == 1) ? 14 : 11, 2, 256, 1, 4, false>;
}
