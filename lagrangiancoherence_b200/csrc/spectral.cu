// Triangular spectral truncation of the global path (LCS.py:115-118: VectorWind(u, v).truncate(field, truncation) ->
// pyspharm grdtospec / spectogrd -> SPHEREPACK shaes / shses) as three small dense products per field:
//      Y = G . Fc            [nlat x nlon] . [nlon x (2T+1)]   zonal Fourier coefficients (mean, cos m, sin m; m <= T)
//      Z[:, c] = A_m(c) . Y[:, c]                              analysis + truncation + synthesis in colatitude
//      out = Z . Fi          [nlat x (2T+1)] . [(2T+1) x nlon]
// The tables A_m, Fc, Fi are built on the host in f64 (lagrangiancoherence_b200/spectral.py, where the algorithm and the
// status of its parity -- unpinned: SPHEREPACK is not available -- are described).  This is pre-processing, once per
// wind level: 53 MFLOP per 360 x 721 field in f64, ~1 % of the integration it precedes.  The 5th-generation tensor cores
// have no f64 path (tcgen05.mma: f16 / bf16 / tf32 / f8 / f6 / f4 only) and a split-precision emulation would buy
// nothing at this size, so these are plain FMA kernels with coalesced table reads and broadcast operands.
#include "lcs_internal.h"

namespace lcs {

// Y[f][i][c] = sum_j G[f][i][j] * Fc[j][c]; block = one (field, row), thread = one coefficient column
template <typename Tin>
__global__ void __launch_bounds__(64)
spectral_lon_forward_kernel(const Tin* __restrict__ g, const double* __restrict__ Fc, int nlat, int nlon, int ncoef,
                            double* __restrict__ Y) {
    const int c = threadIdx.x;
    const size_t row = (size_t)blockIdx.y * nlat + blockIdx.x;
    const Tin* gr = g + row * nlon;
    if (c >= ncoef) return;
    double acc0 = 0.0, acc1 = 0.0;
    int j = 0;
    for (; j + 1 < nlon; j += 2) {
        acc0 = fma((double)gr[j], __ldg(Fc + (size_t)j * ncoef + c), acc0);
        acc1 = fma((double)gr[j + 1], __ldg(Fc + (size_t)(j + 1) * ncoef + c), acc1);
    }
    if (j < nlon) acc0 = fma((double)gr[j], __ldg(Fc + (size_t)j * ncoef + c), acc0);
    Y[row * ncoef + c] = acc0 + acc1;
}

// Z[f][i][c] = sum_k At[m(c)][k][i] * Y[f][k][c]; At = A transposed so that threads (i fastest) read it coalesced
__global__ void __launch_bounds__(128)
spectral_lat_project_kernel(const double* __restrict__ At, const double* __restrict__ Y, int nlat, int ncoef,
                            double* __restrict__ Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, f = blockIdx.z;
    if (i >= nlat) return;
    const int m = (c + 1) >> 1;
    const double* a = At + (size_t)m * nlat * nlat + i;
    const double* y = Y + (size_t)f * nlat * ncoef + c;
    double acc0 = 0.0, acc1 = 0.0;
    int k = 0;
    for (; k + 1 < nlat; k += 2) {
        acc0 = fma(__ldg(a + (size_t)k * nlat), __ldg(y + (size_t)k * ncoef), acc0);
        acc1 = fma(__ldg(a + (size_t)(k + 1) * nlat), __ldg(y + (size_t)(k + 1) * ncoef), acc1);
    }
    if (k < nlat) acc0 = fma(__ldg(a + (size_t)k * nlat), __ldg(y + (size_t)k * ncoef), acc0);
    Z[((size_t)f * nlat + i) * ncoef + c] = acc0 + acc1;
}

// out[f][i][j] = sum_c Z[f][i][c] * Fi[c][j]
__global__ void __launch_bounds__(256)
spectral_lon_inverse_kernel(const double* __restrict__ Z, const double* __restrict__ Fi, int nlat, int nlon, int ncoef,
                            double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = (size_t)blockIdx.z * nlat + blockIdx.y;
    if (j >= nlon) return;
    const double* z = Z + row * ncoef;
    double acc = 0.0;
    for (int c = 0; c < ncoef; ++c) acc = fma(__ldg(z + c), __ldg(Fi + (size_t)c * nlon + j), acc);
    out[row * nlon + j] = acc;
}

}  // namespace lcs

using namespace lcs;

extern "C" size_t lcs_spectral_truncate_scratch_bytes(int nfields, int nlat, int ntrunc) {
    if (nfields < 1 || nlat < 1 || ntrunc < 0) return 0;
    return (size_t)2 * nfields * nlat * (2 * ntrunc + 1) * sizeof(double);
}

extern "C" int lcs_spectral_truncate(const void* in, int in_dtype, int nfields, int nlat, int nlon, int ntrunc,
                                     const double* At, const double* Fc, const double* Fi,
                                     void* scratch, size_t scratch_bytes, double* out, void* stream) {
    if (!in || !At || !Fc || !Fi || !scratch || !out) return lcs_fail(LCS_E_INVALID, "lcs_spectral_truncate: null argument");
    if (nfields < 1 || nlat < 3 || nlon < 4 || ntrunc < 0 || ntrunc > nlat - 1 || 2 * ntrunc + 1 > nlon || 2 * ntrunc + 1 > 64)
        return lcs_fail(LCS_E_INVALID, "lcs_spectral_truncate: bad sizes (need ntrunc <= min(nlat-1, (nlon-1)/2, 31))");
    if (nfields > 65535 || nlat > 65535) return lcs_fail(LCS_E_INVALID, "lcs_spectral_truncate: at most 65535 fields / rows per call");
    if (in_dtype != LCS_F64 && in_dtype != LCS_F32) return lcs_fail(LCS_E_INVALID, "lcs_spectral_truncate: bad in_dtype");
    if (scratch_bytes < lcs_spectral_truncate_scratch_bytes(nfields, nlat, ntrunc))
        return lcs_fail(LCS_E_WORKSPACE, "lcs_spectral_truncate: scratch too small");
    if (in == (const void*)out) return lcs_fail(LCS_E_INVALID, "lcs_spectral_truncate: out may not alias in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ncoef = 2 * ntrunc + 1;
    double* Y = static_cast<double*>(scratch);
    double* Z = Y + (size_t)nfields * nlat * ncoef;
    const dim3 g1((unsigned)nlat, (unsigned)nfields);
    if (in_dtype == LCS_F64) spectral_lon_forward_kernel<double><<<g1, 64, 0, st>>>((const double*)in, Fc, nlat, nlon, ncoef, Y);
    else spectral_lon_forward_kernel<float><<<g1, 64, 0, st>>>((const float*)in, Fc, nlat, nlon, ncoef, Y);
    const dim3 g2((unsigned)((nlat + 127) / 128), (unsigned)ncoef, (unsigned)nfields);
    spectral_lat_project_kernel<<<g2, 128, 0, st>>>(At, Y, nlat, ncoef, Z);
    const dim3 g3((unsigned)((nlon + 255) / 256), (unsigned)nlat, (unsigned)nfields);
    spectral_lon_inverse_kernel<<<g3, 256, 0, st>>>(Z, Fi, nlat, nlon, ncoef, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_spectral_truncate");
    lcs_count_launches(3);
    return LCS_OK;
}
