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

// ---------------------------------------------------------------------------------------------
// Global-path regrid (LCS.py:105-114): linear interpolation along latitude, then along longitude, exactly as the two
// successive scipy interp1d passes xarray makes (w_hi*y_hi + w_lo*y_lo, no fused multiply-add), NaN outside the source
// range, and every NaN replaced by the nearest-label value (`u_interp.where(~isnan(u_interp), u_reindex)`).
namespace lcs {
struct AxisPlan { const int* lo; const double* w_hi; const double* w_lo; const unsigned char* valid; const int* nearest; };

template <typename T>
__global__ void __launch_bounds__(256)
regrid_kernel(const T* __restrict__ in, int nlev, int nlat_s, int nlon_s, AxisPlan py, AxisPlan px,
              int nlat_d, int nlon_d, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)nlat_d * nlon_d;
    if (idx >= per * nlev) return;
    const int lev = (int)(idx / per);
    const int rem = (int)(idx - (long long)lev * per);
    const int j = rem / nlon_d, i = rem - j * nlon_d;
    const T* f = in + (size_t)lev * nlat_s * nlon_s;
    double r = nan("");
    if (py.valid[j] && px.valid[i]) {
        const int r0 = py.lo[j], c0 = px.lo[i];
        const double wyh = py.w_hi[j], wyl = py.w_lo[j], wxh = px.w_hi[i], wxl = px.w_lo[i];
        const T* a = f + (size_t)r0 * nlon_s + c0;
        // pass 1 (latitude) at the two source longitudes, pass 2 (longitude) between them
        const double t0 = __dadd_rn(__dmul_rn(wyh, (double)a[nlon_s]), __dmul_rn(wyl, (double)a[0]));
        const double t1 = __dadd_rn(__dmul_rn(wyh, (double)a[nlon_s + 1]), __dmul_rn(wyl, (double)a[1]));
        r = __dadd_rn(__dmul_rn(wxh, t1), __dmul_rn(wxl, t0));
    }
    if (isnan(r)) r = (double)f[(size_t)py.nearest[j] * nlon_s + px.nearest[i]];
    out[idx] = r;
}
}  // namespace lcs

extern "C" int lcs_regrid_linear_nearest(const void* in, int in_dtype, int nlev, int nlat_src, int nlon_src,
                                         const int32_t* lat_lo, const double* lat_w_hi, const double* lat_w_lo,
                                         const uint8_t* lat_valid, const int32_t* lat_nearest,
                                         const int32_t* lon_lo, const double* lon_w_hi, const double* lon_w_lo,
                                         const uint8_t* lon_valid, const int32_t* lon_nearest,
                                         int nlat_dst, int nlon_dst, double* out, void* stream) {
    if (!in || !out || !lat_lo || !lat_w_hi || !lat_w_lo || !lat_valid || !lat_nearest ||
        !lon_lo || !lon_w_hi || !lon_w_lo || !lon_valid || !lon_nearest)
        return lcs_fail(LCS_E_INVALID, "lcs_regrid_linear_nearest: null argument");
    if (nlev < 1 || nlat_src < 2 || nlon_src < 2 || nlat_dst < 1 || nlon_dst < 1)
        return lcs_fail(LCS_E_INVALID, "lcs_regrid_linear_nearest: bad sizes");
    const long long n = (long long)nlev * nlat_dst * nlon_dst;
    const lcs::AxisPlan py = {lat_lo, lat_w_hi, lat_w_lo, lat_valid, lat_nearest};
    const lcs::AxisPlan px = {lon_lo, lon_w_hi, lon_w_lo, lon_valid, lon_nearest};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned gb = (unsigned)((n + 255) / 256);
    if (in_dtype == LCS_F64)
        lcs::regrid_kernel<double><<<gb, 256, 0, st>>>((const double*)in, nlev, nlat_src, nlon_src, py, px, nlat_dst, nlon_dst, out);
    else if (in_dtype == LCS_F32)
        lcs::regrid_kernel<float><<<gb, 256, 0, st>>>((const float*)in, nlev, nlat_src, nlon_src, py, px, nlat_dst, nlon_dst, out);
    else return lcs_fail(LCS_E_INVALID, "lcs_regrid_linear_nearest: bad in_dtype");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_regrid_linear_nearest");
    lcs_count_launches(1);
    return LCS_OK;
}

// ---------------------------------------------------------------------------------------------
// Ridge classification of find_ridges_spherical_hessian (tools.py:93-136), one thread per point.
// The reference loops np.linalg.eig over 2x2 Hessians (tools.py:105-121); LAPACK's dgeev reduces a
// symmetric [[a,b],[b,d]] with dlanv2, which fixes both the ORDER of the eigenvalues (rt1 is the one
// on a's side: z = p + sign(sqrt(p^2+b^2), p), p = (a-d)/2) and the eigenvector matrix
// [[cs,-sn],[sn,cs]], (cs,sn) = (z,b)/hypot(b,z).  Upstream then takes a ROW of that matrix
// (eig[1][argmin(eig[0])], tools.py:108) -- reproduced as executed.  Eigenvalues equal LAPACK's bit
// for bit; eigenvector entries to one ulp (dgeev renormalises), which only matters exactly at the
// |dt| = tolerance threshold.
namespace lcs {

__device__ __forceinline__ double clean_hess(double h) { return (isinf(h) || isnan(h)) ? 0.0 : h; }   // tools.py:93-94

__global__ void __launch_bounds__(256)
ridge_classify_kernel(const double* __restrict__ hxx, const double* __restrict__ hxy, const double* __restrict__ hyy,
                      const double* __restrict__ gx, const double* __restrict__ gy, long long n, double tol,
                      double* __restrict__ dt_prod, double* __restrict__ eigmin,
                      double* __restrict__ dt_raw, double* __restrict__ evec0, double* __restrict__ evec1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = clean_hess(hxx[i]), b = clean_hess(hxy[i]), d = clean_hess(hyy[i]);
    double rt1, rt2, cs, sn;
    if (b == 0.0) { rt1 = a; rt2 = d; cs = 1.0; sn = 0.0; }
    else {
        const double p = __dmul_rn(0.5, __dsub_rn(a, d));
        const double bc = fabs(b);
        const double scale = fmax(fabs(p), bc);
        double z = __dadd_rn(__dmul_rn(__ddiv_rn(p, scale), p), __dmul_rn(__ddiv_rn(bc, scale), bc));
        z = __dadd_rn(p, copysign(__dmul_rn(sqrt(scale), sqrt(z)), p));
        rt1 = __dadd_rn(d, z);
        rt2 = __dsub_rn(d, __dmul_rn(__ddiv_rn(bc, z), bc));
        const double tau = hypot(b, z);
        cs = __ddiv_rn(z, tau); sn = __ddiv_rn(b, tau);
    }
    const bool first = !(rt2 < rt1);                       // np.argmin: first index on ties
    const double r0 = first ? cs : sn, r1 = first ? -sn : cs;
    const double dt = __dadd_rn(__dmul_rn(r0, gx[i]), __dmul_rn(r1, gy[i]));                 // tools.py:116
    const double em = (fabs(rt1) >= fabs(rt2)) ? rt1 : rt2;                                  // tools.py:119
    eigmin[i] = em;
    dt_prod[i] = (!(fabs(dt) > tol) && em < 0.0) ? 1.0 : 0.0;                                 // tools.py:134-136
    if (dt_raw) dt_raw[i] = dt;                                                               // dt_prod_, tools.py:129
    if (evec0) { evec0[i] = r0; evec1[i] = r1; }                                              // the row taken at :108
}

}  // namespace lcs

extern "C" int lcs_ridge_classify(const double* hxx, const double* hxy, const double* hyy, const double* gx,
                                  const double* gy, int64_t n, double tolerance, double* dt_prod, double* eigmin,
                                  double* dt_raw, double* evec0, double* evec1, void* stream) {
    if (!hxx || !hxy || !hyy || !gx || !gy || !dt_prod || !eigmin) return lcs_fail(LCS_E_INVALID, "lcs_ridge_classify: null argument");
    if (n < 0) return lcs_fail(LCS_E_INVALID, "lcs_ridge_classify: negative n");
    if ((evec0 == nullptr) != (evec1 == nullptr)) return lcs_fail(LCS_E_INVALID, "lcs_ridge_classify: evec0/evec1 must both be set");
    if (n == 0) return LCS_OK;
    lcs::ridge_classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        hxx, hxy, hyy, gx, gy, n, tolerance, dt_prod, eigmin, dt_raw, evec0, evec1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_ridge_classify");
    lcs_count_launches(1);
    return LCS_OK;
}
