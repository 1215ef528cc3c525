// Kernel instantiations of the departure-point integrator (see advect_kernels.cuh); one group per translation unit.
#include "advect_kernels.cuh"

namespace lcs {
cudaError_t lcs_launch_f32_es3(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, false, 3, kES>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_es1(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, false, 1, kES>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_p3(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, false, 3, kPair4>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_p3s(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, true, 3, kPair4>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_p1(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, false, 1, kPair4>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_p1s(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, true, 1, kPair4>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_f32_fast3(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<float, false, 3, kES32>(P, nwindows, workspace, st);
}
}  // namespace lcs
