// Kernel instantiations of the departure-point integrator (see advect_kernels.cuh): the f32 dtype-propagation
// variants (R32) of the f64 ES layout, every interpolation order, with scipy's tap order (STRICT: see stage_settls).
#include "advect_kernels.cuh"

namespace lcs {
cudaError_t lcs_launch_r32_es1(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, true, 1, kES, true>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_r32_es2(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, true, 2, kES, true>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_r32_es3(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, true, 3, kES, true>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_r32_es4(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, true, 4, kES, true>(P, nwindows, workspace, st);
}
cudaError_t lcs_launch_r32_es5(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    return launch_advect<double, true, 5, kES, true>(P, nwindows, workspace, st);
}
}  // namespace lcs
