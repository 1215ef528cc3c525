// Kernel instantiations of the departure-point integrator (see advect_kernels.cuh); one group per translation unit.
#include "advect_kernels.cuh"

namespace lcs {
cudaError_t lcs_launch_f64_es2(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, false, 2, kES>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f64_es4(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, false, 4, kES>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f64_es5(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, false, 5, kES>(P, nwindows, workspace, st);
}
}  // namespace lcs
