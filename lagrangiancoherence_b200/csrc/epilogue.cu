// Fused FTLE epilogue: flowmap_gradient (LCS.py:193-223) + six derivative_spherical_coords /
// fourth_order_derivative calls (tools.py:190-267) + the batched spectral norm (LCS.py:145-157)
// in one pass over the departure points.
//
// A warp owns a strip of 28 output columns (+2 halo lanes each side, periodic in x because the
// reference never overrides isglobal=True, tools.py:248) and marches down a chunk of rows with
// a 5-row register window of (X, Y, Z) in f32:
//   x-stencil : neighbours come from lanes l-2..l+2 by warp shuffle;
//   y-stencil : neighbours are the window rows; global rows 0,1 / n-2,n-1 use the halved
//               one-sided differences of tools.py:210-217.
// Precision follows the reference bit for bit: X,Y,Z in f64 (sincos) -> rounded to f32
// (tools.py:258) -> f32 differences -> f64 combination -> rounded to f32 -> divided by the f64
// metric spacing.  sigma_max of [[Xx,Xy,Yx],[Yy,Zx,Zy],[0,0,0]] (the as-executed scrambled 3x3,
// quirk Q7) is evaluated in closed form instead of LAPACK's SVD.
#include <math.h>
#include "lcs_internal.h"

namespace lcs {

constexpr int kStripCols = 28;
constexpr double kEarthR = 6371000.0;
constexpr double kPi = 3.141592653589793;

struct EpiParams {
    const double* x_dep;
    const double* y_dep;
    int nfields, nlat_global, nlon, in_row0, nrow_in, out_row0, nrow_out;
    const double* dx;
    double dy;
    const unsigned char* mask;
    int log_scale;
    double* sigma;
    double* jac;
    int* status;
    int nstrips, nchunks, rows_per_chunk;
};

struct XYZ { float x, y, z; };

__device__ __forceinline__ XYZ to_xyz(double lon_deg, double lat_deg) {
    const double LON = __ddiv_rn(__dmul_rn(lon_deg, kPi), 180.0);                   // LCS.py:195
    const double LAT = __ddiv_rn(__dmul_rn(__dsub_rn(lat_deg, 90.0), kPi), 180.0);  // LCS.py:196
    double sl, cl, so, co;
    sincos(LAT, &sl, &cl);
    sincos(LON, &so, &co);
    const double rs = __dmul_rn(kEarthR, sl);
    XYZ r;
    r.x = __double2float_rn(__dmul_rn(rs, co));                                      // LCS.py:197
    r.y = __double2float_rn(__dmul_rn(rs, so));                                      // LCS.py:198
    r.z = __double2float_rn(__dmul_rn(kEarthR, cl));                                 // LCS.py:199
    return r;                                                                        // f32: tools.py:258
}

// tools.py:204-207 / 225-228 under numba typing: f32 differences, f64 combination, f32 store
__device__ __forceinline__ float centred4(float p1, float m1, float p2, float m2) {
    const float d1 = __fsub_rn(p1, m1), d2 = __fsub_rn(p2, m2);
    const double a = __dmul_rn(__dmul_rn(4.0 / 3.0, (double)d1), 0.5);
    const double b = __dmul_rn(__dmul_rn(1.0 / 3.0, (double)d2), 0.25);
    return __double2float_rn(__dsub_rn(a, b));
}
__device__ __forceinline__ float onesided(float hi, float lo) {                      // tools.py:210-217
    return __double2float_rn(__dmul_rn((double)__fsub_rn(hi, lo), 0.5));
}

__device__ __forceinline__ float ddy(const float (&w)[5], int r, int n) {
    if (r >= n - 2) return onesided(w[2], w[1]);          // last two rows win (written last, :214)
    if (r < 2) return onesided(w[3], w[2]);
    return centred4(w[3], w[1], w[4], w[0]);
}

__device__ __forceinline__ float ddx(float c, unsigned lane) {
    const float m1 = __shfl_up_sync(0xffffffffu, c, 1), p1 = __shfl_down_sync(0xffffffffu, c, 1);
    const float m2 = __shfl_up_sync(0xffffffffu, c, 2), p2 = __shfl_down_sync(0xffffffffu, c, 2);
    (void)lane;
    return centred4(p1, m1, p2, m2);
}

__global__ void __launch_bounds__(128)
ftle_epilogue_kernel(const EpiParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long per_field = (long long)P.nstrips * P.nchunks;
    if (wg >= per_field * P.nfields) return;              // whole warp leaves together
    const int f = (int)(wg / per_field);
    const int rem = (int)(wg - (long long)f * per_field);
    const int chunk = rem / P.nstrips, strip = rem - chunk * P.nstrips;
    const int col_raw = strip * kStripCols + (int)lane - 2;
    int colw = col_raw % P.nlon;
    if (colw < 0) colw += P.nlon;
    const int r_lo = P.out_row0 + chunk * P.rows_per_chunk;
    const int r_end = P.out_row0 + P.nrow_out;
    const int r_hi = (r_lo + P.rows_per_chunk < r_end) ? r_lo + P.rows_per_chunk : r_end;
    const bool writer = lane >= 2 && lane < 2 + kStripCols && col_raw < P.nlon;
    const double* xin = P.x_dep + (size_t)f * P.nrow_in * P.nlon + colw;
    const double* yin = P.y_dep + (size_t)f * P.nrow_in * P.nlon + colw;

    float wx[5] = {0, 0, 0, 0, 0}, wy[5] = {0, 0, 0, 0, 0}, wz[5] = {0, 0, 0, 0, 0};
    for (int g = r_lo - 2; g < r_hi + 2; ++g) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { wx[k] = wx[k + 1]; wy[k] = wy[k + 1]; wz[k] = wz[k + 1]; }
        if (g >= 0 && g < P.nlat_global) {
            const size_t o = (size_t)(g - P.in_row0) * P.nlon;
            const XYZ v = to_xyz(__ldg(xin + o), __ldg(yin + o));
            wx[4] = v.x; wy[4] = v.y; wz[4] = v.z;
        }
        const int r = g - 2;                              // window is now rows r-2 .. r+2
        if (r < r_lo) continue;
        const float fXx = ddx(wx[2], lane), fYx = ddx(wy[2], lane), fZx = ddx(wz[2], lane);
        const float fXy = ddy(wx, r, P.nlat_global), fYy = ddy(wy, r, P.nlat_global), fZy = ddy(wz, r, P.nlat_global);
        if (!writer) continue;
        const double dxr = __ldg(P.dx + r);
        const double a = __ddiv_rn((double)fXx, dxr);     // dXdx   tools.py:264
        const double b = __ddiv_rn((double)fXy, P.dy);    // dXdy   tools.py:262
        const double c = __ddiv_rn((double)fYx, dxr);     // dYdx
        const double d = __ddiv_rn((double)fYy, P.dy);    // dYdy
        const double e = __ddiv_rn((double)fZx, dxr);     // dZdx
        const double ff = __ddiv_rn((double)fZy, P.dy);   // dZdy
        const size_t oo = (size_t)(r - P.out_row0) * P.nlon + col_raw;
        const size_t plane = (size_t)P.nrow_out * P.nlon;
        if (P.jac) {
            double* j = P.jac + (size_t)f * 6 * plane + oo;
            j[0] = a; j[plane] = b; j[2 * plane] = c; j[3 * plane] = d; j[4 * plane] = e; j[5 * plane] = ff;
        }
        double s;
        const bool anynan = isnan(a) || isnan(b) || isnan(c) || isnan(d) || isnan(e) || isnan(ff);
        if (anynan || (P.mask && !P.mask[oo])) {
            s = nan("");                                   // dropna('points') / crop, LCS.py:143-146
        } else {
            if (P.status && (isinf(a) || isinf(b) || isinf(c) || isinf(d) || isinf(e) || isinf(ff)))
                atomicOr(P.status, 1);                     // scipy.linalg.norm would raise, LCS.py:154
            // M = [[a,b,c],[d,e,f],[0,0,0]] (LCS.py:152-153 reshape); sigma_max^2 = lambda_max(M M^T)
            const double g11 = a * a + b * b + c * c;
            const double g22 = d * d + e * e + ff * ff;
            const double g12 = a * d + b * e + c * ff;
            const double df = g11 - g22;
            s = sqrt(0.5 * (g11 + g22 + sqrt(df * df + 4.0 * g12 * g12)));
            if (P.log_scale) s = 0.5 * log(s);             // callers' scaling, ideal_vortex.py:282
        }
        P.sigma[(size_t)f * plane + oo] = s;
    }
}

// --------------------------------------------------------------------------- seams
__global__ void __launch_bounds__(256)
fourth_order_derivative_kernel(const float* __restrict__ a, int n0, int n1, int dim, int isglobal, float* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n0 * n1) return;
    const int i = (int)(idx / n1), j = (int)(idx - (long long)i * n1);
    float r = 0.0f;                                       // np.zeros_like, tools.py:198
    if (dim == 0) {
        if (i >= n0 - 2) { if (i - 1 >= 0) r = onesided(a[idx], a[idx - n1]); }
        else if (i < 2) r = onesided(a[idx + n1], a[idx]);
        else r = centred4(a[idx + n1], a[idx - n1], a[idx + 2 * (long long)n1], a[idx - 2 * (long long)n1]);
    } else {
        const float* row = a + (size_t)i * n1;
        if (isglobal) {
            const int jp1 = (j + 1) % n1, jm1 = (j - 1 + n1) % n1, jp2 = (j + 2) % n1, jm2 = (j - 2 + 2 * n1) % n1;
            r = centred4(row[jp1], row[jm1], row[jp2], row[jm2]);
        } else {
            if (j >= n1 - 2) { if (j - 1 >= 0) r = onesided(row[j], row[j - 1]); }
            else if (j < 2) r = onesided(row[j + 1], row[j]);
            else r = centred4(row[j + 1], row[j - 1], row[j + 2], row[j - 2]);
        }
    }
    out[idx] = r;
}

// largest eigenvalue of a symmetric 3x3 by cyclic Jacobi (general seam; the hot path never has
// a non-zero third row and takes the closed form)
__device__ double sym3_lambda_max(double a00, double a01, double a02, double a11, double a12, double a22) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            if (A[p][q] == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double cth = 1.0 / sqrt(t * t + 1.0), sth = t * cth;
            const int r = 3 - p - q;
            const double app = A[p][p], aqq = A[q][q], apq = A[p][q], arp = A[r][p], arq = A[r][q];
            A[p][p] = app - t * apq;
            A[q][q] = aqq + t * apq;
            A[p][q] = A[q][p] = 0.0;
            A[r][p] = A[p][r] = cth * arp - sth * arq;
            A[r][q] = A[q][r] = sth * arp + cth * arq;
        }
    }
    return fmax(A[0][0], fmax(A[1][1], A[2][2]));
}

__global__ void __launch_bounds__(256)
spectral_norm_3x3_kernel(const double* __restrict__ v, long long n, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = v[(size_t)k * n + i];
    const double g00 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    const double g11 = m[3] * m[3] + m[4] * m[4] + m[5] * m[5];
    const double g01 = m[0] * m[3] + m[1] * m[4] + m[2] * m[5];
    double lam;
    if (m[6] == 0.0 && m[7] == 0.0 && m[8] == 0.0) {
        const double df = g00 - g11;
        lam = 0.5 * (g00 + g11 + sqrt(df * df + 4.0 * g01 * g01));
    } else {
        const double g22 = m[6] * m[6] + m[7] * m[7] + m[8] * m[8];
        const double g02 = m[0] * m[6] + m[1] * m[7] + m[2] * m[8];
        const double g12 = m[3] * m[6] + m[4] * m[7] + m[5] * m[8];
        lam = sym3_lambda_max(g00, g01, g02, g11, g12, g22);
    }
    out[i] = sqrt(lam);
}

}  // namespace lcs

using namespace lcs;

extern "C" int lcs_ftle_epilogue(const double* x_dep, const double* y_dep, int nfields,
                                 int nlat_global, int nlon, int in_row0, int nrow_in,
                                 int out_row0, int nrow_out, const double* dx, double dy,
                                 const uint8_t* mask, int log_scale,
                                 double* sigma, double* jac, int32_t* status, void* stream) {
    if (!x_dep || !y_dep || !dx || !sigma) return lcs_fail(LCS_E_INVALID, "lcs_ftle_epilogue: null argument");
    if (nfields < 1 || nlat_global < 5 || nlon < 5 || nrow_out < 1 || nrow_in < 1)
        return lcs_fail(LCS_E_INVALID, "lcs_ftle_epilogue: bad sizes (grid must be at least 5x5)");
    if (out_row0 < 0 || out_row0 + nrow_out > nlat_global) return lcs_fail(LCS_E_INVALID, "lcs_ftle_epilogue: output band outside grid");
    const int need_lo = out_row0 - 2 > 0 ? out_row0 - 2 : 0;
    const int need_hi = out_row0 + nrow_out + 2 < nlat_global ? out_row0 + nrow_out + 2 : nlat_global;
    if (in_row0 > need_lo || in_row0 + nrow_in < need_hi)
        return lcs_fail(LCS_E_INVALID, "lcs_ftle_epilogue: input band lacks the 2-row halo");
    EpiParams P{};
    P.x_dep = x_dep; P.y_dep = y_dep; P.nfields = nfields; P.nlat_global = nlat_global; P.nlon = nlon;
    P.in_row0 = in_row0; P.nrow_in = nrow_in; P.out_row0 = out_row0; P.nrow_out = nrow_out;
    P.dx = dx; P.dy = dy; P.mask = mask; P.log_scale = log_scale; P.sigma = sigma; P.jac = jac; P.status = status;
    P.nstrips = (nlon + kStripCols - 1) / kStripCols;
    // enough warps to cover the machine without excessive halo recomputation (4 extra rows per chunk)
    int rpc = lcs_env_int("LCS_EPILOGUE_ROWS", 0);
    if (rpc <= 0) {
        rpc = 64;                                   // 4 halo rows per chunk: 6 % of recomputed sincos at 64 rows, 25 % at 16
        while (rpc > 4 && (long long)nfields * P.nstrips * ((nrow_out + rpc - 1) / rpc) < 148LL * 16) rpc >>= 1;
    }
    P.rows_per_chunk = rpc;
    P.nchunks = (nrow_out + rpc - 1) / rpc;
    const long long warps = (long long)nfields * P.nstrips * P.nchunks;
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    ftle_epilogue_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_ftle_epilogue");
    lcs_count_launches(1);
    return LCS_OK;
}

extern "C" int lcs_fourth_order_derivative(const float* arr, int n0, int n1, int dim, int isglobal,
                                           float* out, void* stream) {
    if (!arr || !out || arr == out) return lcs_fail(LCS_E_INVALID, "lcs_fourth_order_derivative: bad pointers");
    if (n0 < 3 || n1 < 3 || (dim != 0 && dim != 1)) return lcs_fail(LCS_E_INVALID, "lcs_fourth_order_derivative: bad sizes/dim");
    const long long n = (long long)n0 * n1;
    fourth_order_derivative_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        arr, n0, n1, dim, isglobal, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_fourth_order_derivative");
    lcs_count_launches(1);
    return LCS_OK;
}

extern "C" int lcs_spectral_norm_3x3(const double* vals, int64_t n, double* out, void* stream) {
    if (!vals || !out) return lcs_fail(LCS_E_INVALID, "lcs_spectral_norm_3x3: null argument");
    if (n < 0) return lcs_fail(LCS_E_INVALID, "lcs_spectral_norm_3x3: negative n");
    if (n == 0) return LCS_OK;
    spectral_norm_3x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(vals, n, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_spectral_norm_3x3");
    lcs_count_launches(1);
    return LCS_OK;
}
