// xr_map_coordinates as a stand-alone device operation (tools.py:19-41), the gather-roofline
// microbenchmark, and the library's error plumbing.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "lcs_internal.h"
#include "lcs_device.cuh"

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";

int lcs_fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
int lcs_fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof g_err, "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
    return LCS_E_CUDA;
}
int lcs_env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
int lcs_sm_count() {
    // per device (a process may drive several GPUs from several threads); -1 = not queried yet
    static int cached[64];
    static std::once_flag once;
    std::call_once(once, [] { for (int& c : cached) c = -1; });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return 148; }
    int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (n < 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = 148; }
        __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);          // racing threads store the same value
    }
    return n;
}
static unsigned long long g_launches = 0;
void lcs_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
extern "C" unsigned long long lcs_kernel_launches(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int lcs_abi_version(void) { return LCS_ABI_VERSION; }
extern "C" const char* lcs_last_error(void) { return g_err; }

namespace lcs {

struct MapParams {
    const double* field;
    const double* coef;
    int nlat, nlon;
    double nlat_d, nlon_d, lat_min, lat_span, lon_min, lon_span;
    int order, nrow, ncol, row0, nrow_global;
    const double* pos_x;
    const double* pos_y;
    double* out;
};

__global__ void __launch_bounds__(256)
map_coordinates_kernel(const MapParams P) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)P.nrow * P.ncol) return;
    const int row = (int)(idx / P.ncol);
    const int grow = P.row0 + row;
    const bool pole = (grow < P.order) || (grow >= P.nrow_global - P.order);      // tools.py:31-33
    const double iy = index_map(P.pos_y[idx], P.lat_min, P.lat_span, P.nlat_d);   // tools.py:22
    const double ix = index_map(P.pos_x[idx], P.lon_min, P.lon_span, P.nlon_d);   // tools.py:21
    double o[1];
    if (pole) gather_linear_constant<Scalar64, true>(P.field, P.nlat, P.nlon, iy, ix, o);
    else if (P.order == 3) gather_cubic_wrap<Scalar64, true>(P.coef, P.nlat, P.nlon, iy, ix, o);
    else if (P.order == 1) gather_linear_wrap<Scalar64, true>(P.field, P.nlat, P.nlon, iy, ix, o);
    else if (P.order == 2) gather_spline_wrap<Scalar64, true, 2>(P.coef, P.nlat, P.nlon, iy, ix, o);
    else if (P.order == 4) gather_spline_wrap<Scalar64, true, 4>(P.coef, P.nlat, P.nlon, iy, ix, o);
    else gather_spline_wrap<Scalar64, true, 5>(P.coef, P.nlat, P.nlon, iy, ix, o);
    P.out[idx] = o[0];
}

// ------------------------------------------------------------------ gather roofline microbenchmark
// Same tap pattern as the integrator (TAPS x TAPS neighbourhood of packed elements around a displaced
// copy of the start grid) with nothing else in the loop: integer position arithmetic only (no index
// folding, no weights, no dependent position update), one add per loaded value.  The loads of all
// `iters` rounds are independent, so the measured rate is what the LSU/L1/L2 path can deliver for this
// access pattern.  Every round shifts the whole grid by a few cells (coherent across a warp, like the
// early sub-steps); `jitter` adds a per-particle pseudo-random displacement of up to +-jitter cells
// (decorrelated neighbours, like late sub-steps).  Bytes counted = particles * iters * TAPS^2 * sizeof(element).
template <typename E, int TAPS>
__global__ void __launch_bounds__(256, 4)
gather_peak_kernel(const typename E::type* __restrict__ pairs, int nlat, int nlon,
                   int nrow, int ncol, int jitter, int iters, int band_log2, double* sink) {
    const int r = threadIdx.x & ((1 << band_log2) - 1), c = threadIdx.x >> band_log2;
    const int row = blockIdx.y * (1 << band_log2) + r;
    const int col = blockIdx.x * (256 >> band_log2) + c;
    if (row >= nrow || col >= ncol) return;
    const int w = blockIdx.z;
    const int by = (int)((long long)row * (nlat - 1) / (nrow > 1 ? nrow - 1 : 1));
    const int bx = (int)((long long)col * (nlon - 1) / (ncol > 1 ? ncol - 1 : 1));
    const unsigned seed = (unsigned)(row * ncol + col) * 2654435761u + (unsigned)w * 40503u;
    double acc[E::NV];
#pragma unroll
    for (int v = 0; v < E::NV; ++v) acc[v] = 0.0;
    for (int it = 0; it < iters; ++it) {
        int iy = by + ((it * 3 + w) & 7) - (TAPS / 2 - 1), ix = bx + ((it * 5 + w) & 7) - (TAPS / 2 - 1);
        if (jitter > 0) {
            const unsigned h = seed ^ ((unsigned)it * 2246822519u);
            iy += (int)((h >> 8) % (unsigned)(2 * jitter + 1)) - jitter;
            ix += (int)((h >> 20) % (unsigned)(2 * jitter + 1)) - jitter;
        }
        iy = max(0, min(iy, nlat - TAPS));
        ix = max(0, min(ix, nlon - TAPS));
        const typename E::type* base = pairs + (size_t)iy * nlon + ix;
#pragma unroll
        for (int i = 0; i < TAPS; ++i) {
            double cc[TAPS][E::NV];
#pragma unroll
            for (int j = 0; j < TAPS; ++j) E::ld(base + (size_t)i * nlon + j, cc[j]);
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
#pragma unroll
                for (int v = 0; v < E::NV; ++v) acc[v] += cc[j][v];
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < E::NV; ++v) s += acc[v];
    if (s == 1.2345e308) sink[0] = s;                     // keeps the loads alive, never true
}

// The shared-memory alternative, measured instead of argued: the same block tiling (2 rows x 128 particles) and the same
// coherent per-round shift, but every round first stages the block's tap bounding box (5 rows x 131 columns of 16-B
// elements = 10.5 KB) into shared memory with coalesced loads, and the 4x4 taps are then LDS.128 reads.  This is the best
// case for a tile design (particles of a block stay a compact lattice, which holds only for the first sub-steps of a
// window).  Bytes counted as in gather_peak_kernel (taps only, not the staging traffic).
__global__ void __launch_bounds__(256, 4)
gather_peak_smem_kernel(const d2* __restrict__ pairs, int nlat, int nlon, int nrow, int ncol, int iters, double* sink) {
    constexpr int TH = 2, TW = 128, BH = TH + 3, BW = TW + 3;
    __shared__ d2 tile[BH][BW + 1];
    const int r = threadIdx.x & 1, c = threadIdx.x >> 1;
    const int row0 = blockIdx.y * TH, col0 = blockIdx.x * TW;
    const int row = row0 + r, col = col0 + c;
    const int w = blockIdx.z;
    // the block's first particle fixes the box; all particles of the block keep their lattice offsets (coherent shift)
    const int by0 = (int)((long long)row0 * (nlat - 1) / (nrow > 1 ? nrow - 1 : 1));
    const int bx0 = (int)((long long)col0 * (nlon - 1) / (ncol > 1 ? ncol - 1 : 1));
    double ax = 0.0, ay = 0.0;
    for (int it = 0; it < iters; ++it) {
        int oy = by0 + ((it * 3 + w) & 7) - 1, ox = bx0 + ((it * 5 + w) & 7) - 1;
        oy = max(0, min(oy, nlat - BH));
        ox = max(0, min(ox, nlon - BW));
        __syncthreads();                                   // previous round's reads are done
        for (int e = threadIdx.x; e < BH * BW; e += 256) {
            const int tr = e / BW, tc = e - tr * BW;
            double v[2];
            Vec2<double>::ld(pairs + (size_t)(oy + tr) * nlon + ox + tc, v);
            tile[tr][tc].x = v[0]; tile[tr][tc].y = v[1];
        }
        __syncthreads();
        if (row < nrow && col < ncol) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const d2 t = tile[r + i][c + j];
                    ax += t.x; ay += t.y;
                }
            }
        }
    }
    if (ax + ay == 1.2345e308) sink[0] = ax;               // keeps the loads alive, never true
}

}  // namespace lcs

using namespace lcs;

extern "C" int lcs_map_coordinates(const lcs_grid* g, const double* field, const double* coef, int order,
                                   const double* pos_x, const double* pos_y, int nrow, int ncol,
                                   int row0, int nrow_global, double* out, void* stream) {
    if (!g || !field || !pos_x || !pos_y || !out) return lcs_fail(LCS_E_INVALID, "lcs_map_coordinates: null argument");
    if (order < 1 || order > 5) return lcs_fail(LCS_E_UNSUPPORTED, "lcs_map_coordinates: order must be 1..5");
    if (order >= 2 && !coef) return lcs_fail(LCS_E_INVALID, "lcs_map_coordinates: coef required for orders >= 2");
    if (g->nlat < 4 || g->nlon < 4 || nrow < 1 || ncol < 1) return lcs_fail(LCS_E_INVALID, "lcs_map_coordinates: bad sizes");
    MapParams P{};
    P.field = field; P.coef = coef; P.nlat = g->nlat; P.nlon = g->nlon;
    P.nlat_d = (double)g->nlat; P.nlon_d = (double)g->nlon;
    P.lat_min = g->lat_min; P.lat_span = g->lat_max - g->lat_min;
    P.lon_min = g->lon_min; P.lon_span = g->lon_max - g->lon_min;
    P.order = order; P.nrow = nrow; P.ncol = ncol; P.row0 = row0; P.nrow_global = nrow_global;
    P.pos_x = pos_x; P.pos_y = pos_y; P.out = out;
    const long long n = (long long)nrow * ncol;
    map_coordinates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
    lcs_count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_map_coordinates");
    return LCS_OK;
}

extern "C" int lcs_gather_peak_smem(const void* pairs, int nlat, int nlon, int nrow, int ncol, int nwindows, int iters,
                                    double* sink, void* stream) {
    if (!pairs || !sink) return lcs_fail(LCS_E_INVALID, "lcs_gather_peak_smem: null argument");
    if (nlat < 8 || nlon < 136 || nrow < 1 || ncol < 1 || nwindows < 1 || nwindows > 65535 || iters < 1)
        return lcs_fail(LCS_E_INVALID, "lcs_gather_peak_smem: bad sizes (the tile needs nlat >= 8, nlon >= 136)");
    const dim3 grid((unsigned)((ncol + 127) / 128), (unsigned)((nrow + 1) / 2), (unsigned)nwindows);
    gather_peak_smem_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>((const d2*)pairs, nlat, nlon, nrow, ncol, iters, sink);
    lcs_count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_gather_peak_smem");
    return LCS_OK;
}

extern "C" int lcs_gather_peak(const void* pairs, int pair_dtype, int vec_width, int nlat, int nlon, int nrow, int ncol,
                               int nwindows, int taps, double jitter, int iters, double* sink, void* stream) {
    if (!pairs || !sink) return lcs_fail(LCS_E_INVALID, "lcs_gather_peak: null argument");
    if (nlat < 4 || nlon < 4 || nrow < 1 || ncol < 1 || nwindows < 1 || nwindows > 65535 || iters < 1 || jitter < 0)
        return lcs_fail(LCS_E_INVALID, "lcs_gather_peak: bad sizes");
    if ((taps != 2 && taps != 4) || (vec_width != 2 && vec_width != 4) || (pair_dtype != LCS_F64 && pair_dtype != LCS_F32))
        return lcs_fail(LCS_E_INVALID, "lcs_gather_peak: taps and vec_width must be 2 or 4");
    int bl = lcs_env_int("LCS_ADVECT_BAND_LOG2", 1);       // same thread->particle tiling as the integrator
    if (bl < 0) bl = 0;
    if (bl > 5) bl = 5;
    const int tw = 256 >> bl, th = 1 << bl;
    const dim3 grid((unsigned)((ncol + tw - 1) / tw), (unsigned)((nrow + th - 1) / th), (unsigned)nwindows);
    const int jit = (int)jitter;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LCS_GP(EL, TP) gather_peak_kernel<EL, TP><<<grid, 256, 0, st>>>((const EL::type*)pairs, nlat, nlon, nrow, ncol, jit, iters, bl, sink)
    if (pair_dtype == LCS_F64) {
        if (vec_width == 4) { if (taps == 4) LCS_GP(Pair4<double>, 4); else LCS_GP(Pair4<double>, 2); }
        else { if (taps == 4) LCS_GP(Vec2<double>, 4); else LCS_GP(Vec2<double>, 2); }
    } else {
        if (vec_width == 4) { if (taps == 4) LCS_GP(Pair4<float>, 4); else LCS_GP(Pair4<float>, 2); }
        else { if (taps == 4) LCS_GP(Vec2<float>, 4); else LCS_GP(Vec2<float>, 2); }
    }
#undef LCS_GP
    lcs_count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_gather_peak");
    return LCS_OK;
}
