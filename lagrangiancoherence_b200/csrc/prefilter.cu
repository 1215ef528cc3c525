// Wind staging: cubic B-spline prefilter (once per level) and packing into the gather layout.
//
// scipy.ndimage.map_coordinates(order=3, mode='wrap') re-runs spline_filter over the whole field
// inside every call (tools.py:26-30: 18 calls per wind interval at SETTLS_order=4).  The
// coefficients depend only on the level, so here they are computed once per level:
// per axis  c *= (1-z)(1-1/z);  exact mirror causal initialisation;  c[i] += z c[i-1];
// c[n-1] = (z c[n-2] + c[n-1]) z/(z^2-1);  c[i] = z (c[i+1] - c[i]),  z = sqrt(3)-2, axis 0 first.
// Each operation is a separate IEEE f64 op in that order (no FMA) so the result equals the numpy
// restatement in oracle/lcs_oracle.py bit for bit (which is within a few ulp of scipy).
#include <math.h>
#include "lcs_internal.h"
#include "lcs_device.cuh"

namespace lcs {

template <typename Tin>
__device__ __forceinline__ void filter_line(const Tin* src, size_t sstride,
                                            double* dst, size_t dstride, int n,
                                            double z, double gain, double zn1) {
    if (n < 2) { if (n == 1) dst[0] = (double)src[0]; return; }
    // exact causal initialisation for the mirror extension
    double c0 = __dadd_rn(__dmul_rn((double)src[0], gain), __dmul_rn(zn1, __dmul_rn((double)src[(size_t)(n - 1) * sstride], gain)));
    double zi = z;
    for (int i = 1; i < n - 1; ++i) {
        if (zi == 0.0) break;                      // z^i underflowed: every later term adds exactly 0
        const double a = __dmul_rn((double)src[(size_t)i * sstride], gain);
        const double b = __dmul_rn((double)src[(size_t)(n - 1 - i) * sstride], gain);
        c0 = __dadd_rn(c0, __dmul_rn(zi, __dadd_rn(a, __dmul_rn(zn1, b))));
        zi = __dmul_rn(zi, z);
    }
    c0 = __ddiv_rn(c0, __dsub_rn(1.0, __dmul_rn(zn1, zn1)));
    // causal pass
    double prev = c0, prev2 = 0.0;
    // src may alias dst (the row pass is in place): element i is read before it is written
    double cur = c0;
    dst[0] = c0;
    for (int i = 1; i < n; ++i) {
        const double a = __dmul_rn((double)src[(size_t)i * sstride], gain);
        prev2 = prev;
        cur = __dadd_rn(a, __dmul_rn(z, prev));
        dst[(size_t)i * dstride] = cur;
        prev = cur;
    }
    // anticausal initialisation and pass
    double nxt = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(z, prev2), cur), z), __dsub_rn(__dmul_rn(z, z), 1.0));
    dst[(size_t)(n - 1) * dstride] = nxt;
    for (int i = n - 2; i >= 0; --i) {
        nxt = __dmul_rn(z, __dsub_rn(nxt, dst[(size_t)i * dstride]));
        dst[(size_t)i * dstride] = nxt;
    }
}

// axis 0: one thread per (plane, column); adjacent threads walk adjacent columns (coalesced)
template <typename Tin>
__global__ void __launch_bounds__(128)
prefilter_cols_kernel(const Tin* __restrict__ u, const Tin* __restrict__ v, double* cu, double* cv,
                      int nlev, int nlat, int nlon, double z, double gain, double zn1) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)2 * nlev * nlon;
    if (idx >= total) return;
    const int col = (int)(idx % nlon);
    const int pl = (int)(idx / nlon);
    const int lev = pl >> 1;
    const size_t off = (size_t)lev * nlat * nlon + col;
    const Tin* src = ((pl & 1) ? v : u) + off;
    double* dst = ((pl & 1) ? cv : cu) + off;
    filter_line<Tin>(src, (size_t)nlon, dst, (size_t)nlon, nlat, z, gain, zn1);
}

// axis 1, in place: one thread per (plane, row).  Accesses are line-strided; every 128-B line
// a warp touches is reused by its next 15 iterations out of L1.
__global__ void __launch_bounds__(128)
prefilter_rows_kernel(double* cu, double* cv, int nlev, int nlat, int nlon, double z, double gain, double zn1) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)2 * nlev * nlat;
    if (idx >= total) return;
    const int row = (int)(idx % nlat);
    const int pl = (int)(idx / nlat);
    const int lev = pl >> 1;
    double* line = ((pl & 1) ? cv : cu) + ((size_t)lev * nlat + row) * nlon;
    filter_line<double>(line, 1, line, 1, nlon, z, gain, zn1);
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
pack_pairs_kernel(const Tin* __restrict__ u, const Tin* __restrict__ v,
                  typename PairOf<Tout>::type* __restrict__ pairs, long long plane, int npairs) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plane * npairs) return;
    typename PairOf<Tout>::type o;
    o.x = (Tout)u[idx];
    o.y = (Tout)v[idx];
    o.z = (Tout)u[idx + plane];
    o.w = (Tout)v[idx + plane];
    pairs[idx] = o;
}

}  // namespace lcs

using namespace lcs;

extern "C" int lcs_prefilter(const void* u, const void* v, int in_dtype, double* coef_u, double* coef_v,
                             int nlev, int nlat, int nlon, void* stream) {
    if (!u || !v || !coef_u || !coef_v) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: null argument");
    if (nlev < 1 || nlat < 2 || nlon < 2) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: bad sizes");
    if (u == coef_u || v == coef_v) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: outputs may not alias inputs");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double z = sqrt(3.0) - 2.0;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    const long long ncol_threads = (long long)2 * nlev * nlon;
    const unsigned gb = (unsigned)((ncol_threads + 127) / 128);
    if (in_dtype == LCS_F64)
        prefilter_cols_kernel<double><<<gb, 128, 0, st>>>((const double*)u, (const double*)v, coef_u, coef_v,
                                                          nlev, nlat, nlon, z, gain, pow(z, nlat - 1));
    else if (in_dtype == LCS_F32)
        prefilter_cols_kernel<float><<<gb, 128, 0, st>>>((const float*)u, (const float*)v, coef_u, coef_v,
                                                         nlev, nlat, nlon, z, gain, pow(z, nlat - 1));
    else return lcs_fail(LCS_E_INVALID, "lcs_prefilter: bad in_dtype");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(cols)");
    const long long nrow_threads = (long long)2 * nlev * nlat;
    prefilter_rows_kernel<<<(unsigned)((nrow_threads + 127) / 128), 128, 0, st>>>(coef_u, coef_v, nlev, nlat, nlon,
                                                                                  z, gain, pow(z, nlon - 1));
    e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(rows)");
    return LCS_OK;
}

extern "C" int lcs_pack_pairs(const void* u, const void* v, int in_dtype, void* pairs, int pair_dtype,
                              int nlev, int nlat, int nlon, void* stream) {
    if (!u || !v || !pairs) return lcs_fail(LCS_E_INVALID, "lcs_pack_pairs: null argument");
    if (nlev < 2 || nlat < 1 || nlon < 1) return lcs_fail(LCS_E_INVALID, "lcs_pack_pairs: need at least two levels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long plane = (long long)nlat * nlon;
    const int npairs = nlev - 1;
    const unsigned gb = (unsigned)((plane * npairs + 255) / 256);
    if (in_dtype == LCS_F64 && pair_dtype == LCS_F64)
        pack_pairs_kernel<double, double><<<gb, 256, 0, st>>>((const double*)u, (const double*)v, (d4*)pairs, plane, npairs);
    else if (in_dtype == LCS_F64 && pair_dtype == LCS_F32)
        pack_pairs_kernel<double, float><<<gb, 256, 0, st>>>((const double*)u, (const double*)v, (float4*)pairs, plane, npairs);
    else if (in_dtype == LCS_F32 && pair_dtype == LCS_F64)
        pack_pairs_kernel<float, double><<<gb, 256, 0, st>>>((const float*)u, (const float*)v, (d4*)pairs, plane, npairs);
    else if (in_dtype == LCS_F32 && pair_dtype == LCS_F32)
        pack_pairs_kernel<float, float><<<gb, 256, 0, st>>>((const float*)u, (const float*)v, (float4*)pairs, plane, npairs);
    else return lcs_fail(LCS_E_INVALID, "lcs_pack_pairs: bad dtype");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_pack_pairs");
    return LCS_OK;
}
