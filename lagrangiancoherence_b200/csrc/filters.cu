// The two optional pre-/post-processing steps of LCS.__call__ that sit on the hot path's edges:
//  * lcs_time_lerp         -- `u.resample({time: f}).interpolate('linear')` (LCS.py:88-91): linear
//                             refinement of the wind series in time, evaluated as scipy 1.18.1's
//                             interp1d(kind='linear') does (the routine xarray calls):
//                             y = w_hi * y_hi + w_lo * y_lo,  w_hi = (x-x_lo)/(x_hi-x_lo), w_lo = (x_hi-x)/(x_hi-x_lo);
//  * lcs_gaussian_filter2d -- `scipy.ndimage.gaussian_filter(x_departure, sigma)` (LCS.py:187-190):
//                             separable, mode='reflect' (d c b a | a b c d | d c b a), truncate=4,
//                             axis 0 then axis 1, accumulated in scipy's symmetric-kernel order
//                             (centre tap first, then pairs from the outermost inwards), every
//                             operation a separate IEEE f64 op so the result equals scipy's bit for bit.
#include "lcs_internal.h"

namespace lcs {

template <typename Tin>
__global__ void __launch_bounds__(256)
time_lerp_kernel(const Tin* __restrict__ in, const int* __restrict__ lo, const double* __restrict__ w_hi,
                 const double* __restrict__ w_lo, long long plane, int nnew, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plane * nnew) return;
    const int k = (int)(idx / plane);
    const long long c = idx - (long long)k * plane;
    const int j = lo[k];
    const double ylo = (double)in[(size_t)j * plane + c];
    const double yhi = (double)in[(size_t)(j + 1) * plane + c];
    out[idx] = __dadd_rn(__dmul_rn(w_hi[k], yhi), __dmul_rn(w_lo[k], ylo));
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

// one 1-D pass along axis `along0 ? 0 : 1` of [nfields][n0][n1]
__global__ void __launch_bounds__(256)
gaussian_pass_kernel(const double* __restrict__ in, double* __restrict__ out, int nfields, int n0, int n1,
                     int along0, const double* __restrict__ w, int radius) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long plane = (long long)n0 * n1;
    if (idx >= plane * nfields) return;
    const long long f = idx / plane;
    const int rem = (int)(idx - f * plane);
    const int i = rem / n1, j = rem - i * n1;
    const double* base = in + f * plane;
    const int l = along0 ? i : j, n = along0 ? n0 : n1;
    const long long stride = along0 ? n1 : 1;
    const double* line = base + (along0 ? j : (long long)i * n1);
    double tmp = __dmul_rn(line[l * stride], w[radius]);
    for (int ii = -radius; ii < 0; ++ii) {
        const double a = line[reflect_idx(l + ii, n) * stride], b = line[reflect_idx(l - ii, n) * stride];
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), w[ii + radius]));
    }
    out[idx] = tmp;
}

}  // namespace lcs

using namespace lcs;

extern "C" int lcs_time_lerp(const void* in, int in_dtype, const int32_t* lo, const double* w_hi, const double* w_lo,
                             int nnew, int64_t plane, double* out, void* stream) {
    if (!in || !lo || !w_hi || !w_lo || !out) return lcs_fail(LCS_E_INVALID, "lcs_time_lerp: null argument");
    if (nnew < 1 || plane < 1) return lcs_fail(LCS_E_INVALID, "lcs_time_lerp: bad sizes");
    const long long n = (long long)plane * nnew;
    const unsigned gb = (unsigned)((n + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dtype == LCS_F64) time_lerp_kernel<double><<<gb, 256, 0, st>>>((const double*)in, lo, w_hi, w_lo, plane, nnew, out);
    else if (in_dtype == LCS_F32) time_lerp_kernel<float><<<gb, 256, 0, st>>>((const float*)in, lo, w_hi, w_lo, plane, nnew, out);
    else return lcs_fail(LCS_E_INVALID, "lcs_time_lerp: bad in_dtype");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_time_lerp");
    lcs_count_launches(1);
    return LCS_OK;
}

extern "C" int lcs_gaussian_filter2d(const double* in, double* out, double* scratch, int nfields, int n0, int n1,
                                     const double* weights, int radius, void* stream) {
    if (!in || !out || !scratch || !weights) return lcs_fail(LCS_E_INVALID, "lcs_gaussian_filter2d: null argument");
    if (nfields < 1 || n0 < 1 || n1 < 1 || radius < 0) return lcs_fail(LCS_E_INVALID, "lcs_gaussian_filter2d: bad sizes");
    if (in == out || in == scratch || out == scratch) return lcs_fail(LCS_E_INVALID, "lcs_gaussian_filter2d: buffers may not alias");
    const long long n = (long long)nfields * n0 * n1;
    const unsigned gb = (unsigned)((n + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    gaussian_pass_kernel<<<gb, 256, 0, st>>>(in, scratch, nfields, n0, n1, 1, weights, radius);
    gaussian_pass_kernel<<<gb, 256, 0, st>>>(scratch, out, nfields, n0, n1, 0, weights, radius);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_gaussian_filter2d");
    lcs_count_launches(2);
    return LCS_OK;
}
