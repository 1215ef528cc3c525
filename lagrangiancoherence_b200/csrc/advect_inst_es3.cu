// Kernel instantiations of the departure-point integrator (see advect_kernels.cuh); one group per translation unit.
#include "advect_kernels.cuh"

namespace lcs {
cudaError_t lcs_launch_f64_es3(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, false, 3, kES>(P, nwindows, workspace, st);
}
}  // namespace lcs
