// Wind staging: B-spline prefilter (once per level) and packing into the gather layout.
//
// scipy.ndimage.map_coordinates(order >= 2, mode='wrap') re-runs spline_filter over the whole field inside every
// call (tools.py:26-30: 18 calls per wind interval at SETTLS_order=4).  The coefficients depend only on the level,
// so here they are computed once per level.
//
// scipy's filter is, per axis and per pole z, the gain (1-z)(1-1/z) times a causal and an anticausal one-pole
// recursion with the *exact* initialisation for the mirror extension (d c b | a b c d | c b a).  The recursion is
// sequential along a line; its closed form is a symmetric two-sided exponential
//      c[i] = sum_k  h0 * z^|k| * s[mirror(i + k)],
// and |z| <= 0.43, so truncating at |k| <= log(1e-18)/log|z| leaves a remainder far below one f64 ulp: every run of
// outputs can start from truncated end sums.  Two kernels, chosen by lcs_prefilter: iir_column_kernel (series: a thread
// walks a whole column with the tiles it needs in registers, every input read once -- 66-70 % of the copy bandwidth) and
// iir_axis0_transpose_kernel (single fields and the two-pole orders: each thread a run of KQ consecutive outputs, the two
// one-sided sums at the run's ends by Horner over the mirror extension, the same one-pole recursions scipy uses inside
// the run).  Both passes filter along axis 0 of their input with threads along axis 1
// (coalesced) and write their result transposed, so two applications filter axis 0 (lat) then axis 1 (lon) --
// scipy's order -- and restore the layout.  Agreement with scipy.ndimage.spline_filter is ~1e-15 of the field
// magnitude (tests/test_gpu_engine.py), not bitwise: the summation order differs from the whole-line recursion.
#include <math.h>
#include "lcs_internal.h"
#include "lcs_device.cuh"

namespace lcs {

// Inside a thread's run of KQ consecutive outputs the two one-sided sums obey the one-pole recursions that scipy runs
// along the whole line,
//      a[i] = s[i] + z a[i-1]   (causal),      m[i] = s[i] + z m[i+1]   (anticausal),      c[i] = sqrt(3) (a[i] + z m[i+1]),
// and only their values at the ends of the run need the truncated (|k| <= kh, remainder < 1e-18) mirror sums.  Per run:
// 2*(kh+1) Horner steps for the two ends + 2*KQ recursion steps (the first version of this file was a 65-tap FIR per
// output: 3.5 ms against 1.4 ms per pass on the bench step).  ncu of the register form, round 2
// (profiles/r02_prof_staging_B1184*): 1.45 ms per pass over 2 x 1192 C2 levels = 2.35 TB/s of DRAM traffic (36 % of the
// copy bandwidth), L1 pipe 62 %, 27 % of the warp slots at 96 registers.
#ifndef LCS_PREFILTER_RUN
#define LCS_PREFILTER_RUN 32        // ncu launch lists, 1192 C2 levels x 2 components, lat + lon pass: runs of 32 at 96 registers
                                    // 1.31 + 1.37 / 1.45 + 1.50 ms on two boxes; runs of 24 capped at 64 registers 1.40 + 1.45 ms
#endif
constexpr int KQ = LCS_PREFILTER_RUN;    // outputs per thread in the recursive form

struct MirrorWalk {        // index into the mirror extension d c b | a b c d | c b a, stepped by +-1
    int i, dir, n;
    __device__ __forceinline__ MirrorWalk(int start, int n_, int step) : n(n_) {
        const int period = 2 * n - 2;
        int ii = start % period;
        if (ii < 0) ii += period;
        dir = step;
        if (ii >= n) { ii = period - ii; dir = -step; }
        i = ii;
        fix();
    }
    __device__ __forceinline__ void fix() { if (i == n - 1 && dir > 0) dir = -1; else if (i == 0 && dir < 0) dir = 1; }
    __device__ __forceinline__ void next() { i += dir; fix(); }
};

// Round 2b: the causal results of a run live in a shared-memory tile [column][j] instead of 64 registers, the outputs
// overwrite them in place, and the block then writes the tile out TRANSPOSED with coalesced stores (a warp per
// destination row: KQ consecutive doubles).  The first form stored o[j] straight from the thread that computed it --
// 32 lanes, 32 different destination rows, one 8-B piece of a sector each: 32 L1 wavefronts per store instruction,
// 1024 of the ~1300 wavefronts a warp spent on a run (ncu: L1 pipe 62 % busy at 36 % of the copy bandwidth).
constexpr int kIirThreads = 128;
#ifndef LCS_PREFILTER_BATCH
#define LCS_PREFILTER_BATCH 8
#endif
constexpr int kIirBatch = LCS_PREFILTER_BATCH;
static_assert(KQ % kIirBatch == 0, "the run length must be a multiple of the load batch");
template <typename Tin>
__global__ void __launch_bounds__(kIirThreads)
iir_axis0_transpose_kernel(const Tin* __restrict__ in_a, const Tin* __restrict__ in_b, int interleaved_planes,
                           double* __restrict__ out_a, double* __restrict__ out_b, int split_out,
                           int n0, int n1, double z, double h0, int kh) {
    __shared__ double tile[kIirThreads][KQ + 1];            // +1: a thread walks its own row, lanes are 33 doubles apart (no bank conflicts)
    const int c_raw = blockIdx.x * kIirThreads + threadIdx.x;
    const int c = c_raw < n1 ? c_raw : n1 - 1;              // threads past the edge redo the last column (they must reach the barrier)
    const int r0 = blockIdx.y * KQ;
    const int p = blockIdx.z;
    const size_t plane = (size_t)n0 * n1;
    const Tin* src = interleaved_planes ? ((p & 1) ? in_b : in_a) + (size_t)(p >> 1) * plane + c
                                        : in_a + (size_t)p * plane + c;
    double* dst = split_out ? ((p & 1) ? out_b : out_a) + (size_t)(p >> 1) * plane
                            : out_a + (size_t)p * plane;
    auto ld = [&](int i) { LCS_ASSERT(i >= 0 && i < n0); return (double)__ldg(src + (size_t)i * n1); };
    double* const mine = tile[threadIdx.x];
    // Loads are taken kIirBatch at a time ahead of the recursion steps that consume them: a rolled loop has ONE load in
    // flight per thread (load, dependent FMA, next load), i.e. 2 x (kh + KQ) serialised L2 round trips per run.
    // causal sum just before the run: a[r0-1] = sum_{k>=0} z^k s[r0-1-k], Horner from the far end
    double a;
    {
        MirrorWalk w(r0 - 1 - kh, n0, +1);
        a = ld(w.i);
        for (int k = 0; k < kh; k += kIirBatch) {
            double v[kIirBatch];
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) { w.next(); v[t] = ld(w.i); }
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) if (k + t < kh) a = fma(z, a, v[t]);
        }
    }
    {
        MirrorWalk w(r0, n0, +1);
#pragma unroll
        for (int j0 = 0; j0 < KQ; j0 += kIirBatch) {
            double v[kIirBatch];
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) { v[t] = ld(w.i); w.next(); }
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) { a = fma(z, a, v[t]); mine[j0 + t] = a; }
        }
    }
    // anticausal sum just after the run: m[r0+KQ] = sum_{k>=0} z^k s[r0+KQ+k]
    double m;
    {
        MirrorWalk w(r0 + KQ + kh, n0, -1);
        m = ld(w.i);
        for (int k = 0; k < kh; k += kIirBatch) {
            double v[kIirBatch];
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) { w.next(); v[t] = ld(w.i); }
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) if (k + t < kh) m = fma(z, m, v[t]);
        }
    }
    {
        MirrorWalk w(r0 + KQ - 1, n0, -1);
#pragma unroll
        for (int j0 = KQ - 1; j0 >= 0; j0 -= kIirBatch) {
            double v[kIirBatch];
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) { v[t] = ld(w.i); w.next(); }
#pragma unroll
            for (int t = 0; t < kIirBatch; ++t) {
                mine[j0 - t] = h0 * fma(z, m, mine[j0 - t]);
                m = fma(z, m, v[t]);
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kIirThreads;
    for (int cl = warp; cl < kIirThreads && c0 + cl < n1; cl += kIirThreads / 32) {
        double* o = dst + (size_t)(c0 + cl) * n0 + r0;
        for (int j = lane; j < KQ; j += 32)
            if (r0 + j < n0) o[j] = tile[cl][j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Column-streaming form (round 2b; the default for one-pole filters, i.e. orders 2 and 3).  A thread owns ONE column of
// one plane and walks down the whole axis in tiles of CQ rows, with the tiles it needs in REGISTERS: the one being
// written out, the two below it (kh <= 2 CQ rows restart the anticausal recursion) and one in flight.  Every input value
// is loaded exactly once, coalesced across the block's columns; the run form above reads each value 4.1 times through an
// L1 that its tile has squeezed to a few tens of KB and is bound by L2 -> SM bandwidth (2.1-2.4 ms for the two passes over
// 2 x 1192 C2 levels whatever the load batching).  The causal state is carried from tile to tile exactly -- scipy's own
// whole-line recursion, started 2 CQ rows into the mirror extension with the truncated sum (|z|^32 < 1e-18).  A first
// version kept the tiles in a shared-memory ring filled by cp.async: 1.5 KB of shared memory per thread left 4 warps per
// SM and every LDS latency exposed -- 4.0 ms.  Per output: 1 LDG, 5 FMAs, and the transposed copy-out through a small
// shared tile (coalesced stores).
constexpr int kColThreads = 64;
constexpr int CQ = 16;
#ifndef LCS_PREFILTER_AHEAD
#define LCS_PREFILTER_AHEAD 1
#endif
constexpr int kColAhead = LCS_PREFILTER_AHEAD;     // tiles in flight beyond the three being read
__device__ __forceinline__ int mirror_row(int r, int n) {       // d c b | a b c d | c b a, any r
    const int period = 2 * n - 2;
    r %= period;
    if (r < 0) r += period;
    return r < n ? r : period - r;
}

template <typename Tin>
__global__ void __launch_bounds__(kColThreads)
iir_column_kernel(const Tin* __restrict__ in_a, const Tin* __restrict__ in_b, int interleaved_planes,
                  double* __restrict__ out_a, double* __restrict__ out_b, int split_out,
                  int n0, int n1, double z, double h0) {
    __shared__ double outt[2][kColThreads][CQ + 1];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kColThreads;
    const int c = min(c0 + tid, n1 - 1);                    // threads past the edge redo the last column (they must reach the barriers)
    const int p = blockIdx.z;
    const size_t plane = (size_t)n0 * n1;
    const Tin* src = interleaved_planes ? ((p & 1) ? in_b : in_a) + (size_t)(p >> 1) * plane + c
                                        : in_a + (size_t)p * plane + c;
    double* dst = split_out ? ((p & 1) ? out_b : out_a) + (size_t)(p >> 1) * plane
                            : out_a + (size_t)p * plane;
    const int ntile = (n0 + CQ - 1) / CQ;
    auto load_tile = [&](int t, double (&S)[CQ]) {          // rows [t CQ, (t + 1) CQ) of the mirror extension
        const int r = t * CQ;
        if (r >= 0 && r + CQ <= n0) {                       // interior tile: no row is reflected
            const Tin* g = src + (size_t)r * n1;
#pragma unroll
            for (int j = 0; j < CQ; ++j) S[j] = (double)__ldg(g + (size_t)j * n1);
        } else {
#pragma unroll
            for (int j = 0; j < CQ; ++j) {
                const int rr = mirror_row(r + j, n0);
                LCS_ASSERT(rr >= 0 && rr < n0);
                S[j] = (double)__ldg(src + (size_t)rr * n1);
            }
        }
    };
    double S0[CQ], S1[CQ], S2[CQ], P[kColAhead][CQ], A[CQ];
    // causal start: two tiles of the mirror extension above row 0, then tile 0
    double a = 0.0;
    load_tile(-2, S0); load_tile(-1, S1); load_tile(0, S2);
#pragma unroll
    for (int j = 0; j < CQ; ++j) a = fma(z, a, S0[j]);
#pragma unroll
    for (int j = 0; j < CQ; ++j) a = fma(z, a, S1[j]);
#pragma unroll
    for (int j = 0; j < CQ; ++j) { a = fma(z, a, S2[j]); A[j] = a; S0[j] = S2[j]; }
    load_tile(1, S1); load_tile(2, S2);
#pragma unroll
    for (int k = 0; k + 1 < kColAhead; ++k) load_tile(3 + k, P[k]);
    const int half = (tid >> 4) & 1, l16 = tid & 15, warp = tid >> 5;
    for (int i = 0; i < ntile; ++i) {
        load_tile(i + 2 + kColAhead, P[kColAhead - 1]);     // in flight during the arithmetic of kColAhead tiles
        // anticausal restart below tile i: m = sum_{k < 2 CQ} z^k s[(i + 1) CQ + k]
        double m = 0.0;
#pragma unroll
        for (int j = CQ - 1; j >= 0; --j) m = fma(z, m, S2[j]);
#pragma unroll
        for (int j = CQ - 1; j >= 0; --j) m = fma(z, m, S1[j]);
        double (*ot)[CQ + 1] = outt[i & 1];
#pragma unroll
        for (int j = CQ - 1; j >= 0; --j) {
            ot[tid][j] = h0 * fma(z, m, A[j]);
            m = fma(z, m, S0[j]);
        }
        // causal pass of tile i + 1
#pragma unroll
        for (int j = 0; j < CQ; ++j) { a = fma(z, a, S1[j]); A[j] = a; }
        __syncthreads();                                    // out tile complete (double-buffered: one barrier per tile)
        // transposed copy-out: half a warp per destination row (CQ consecutive doubles)
        const int r0 = i * CQ;
        for (int cl = 2 * warp + half; cl < kColThreads && c0 + cl < n1; cl += kColThreads / 16) {
            LCS_ASSERT(c0 + cl < n1 && cl < kColThreads);
            if (r0 + l16 < n0) dst[(size_t)(c0 + cl) * n0 + r0 + l16] = ot[cl][l16];
        }
#pragma unroll
        for (int j = 0; j < CQ; ++j) {
            S0[j] = S1[j]; S1[j] = S2[j]; S2[j] = P[0][j];
#pragma unroll
            for (int k = 0; k + 1 < kColAhead; ++k) P[k][j] = P[k + 1][j];
        }
    }
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

// E[k] = (u_k, v_k), k < nlev;  S[k] = (2u_k - u_{k+1}, 2v_k - v_{k+1}), k < nlev-1 (f64 arithmetic), written in the
// halo layout of include/lcs_b200.h: one thread per padded cell, halo cells read their mirror image.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256)
pack_es_kernel(const Tin* __restrict__ u, const Tin* __restrict__ v,
               typename Vec2Of<Tout>::type* __restrict__ e_out, typename Vec2Of<Tout>::type* __restrict__ s_out,
               int nlat, int nlon, int nlev) {
    const int pitch = nlon + LCS_HALO_LO + LCS_HALO_HI;
    const int pc = blockIdx.x * blockDim.x + threadIdx.x;
    if (pc >= pitch) return;
    const int pr = blockIdx.y, k = blockIdx.z;
    const int r = mirror_near(pr - LCS_HALO_LO, nlat), c = mirror_near(pc - LCS_HALO_LO, nlon);
    const size_t plane = (size_t)nlat * nlon;
    const size_t src = (size_t)k * plane + (size_t)r * nlon + c;
    const size_t dst = ((size_t)k * (nlat + LCS_HALO_LO + LCS_HALO_HI) + pr) * pitch + pc;
    const double uk = (double)u[src], vk = (double)v[src];
    typename Vec2Of<Tout>::type e;
    e.x = (Tout)uk; e.y = (Tout)vk;
    e_out[dst] = e;
    if (k < nlev - 1) {
        typename Vec2Of<Tout>::type s;
        s.x = (Tout)__dsub_rn(__dmul_rn(2.0, uk), (double)u[src + plane]);
        s.y = (Tout)__dsub_rn(__dmul_rn(2.0, vk), (double)v[src + plane]);
        s_out[dst] = s;
    }
}

}  // namespace lcs

using namespace lcs;

extern "C" size_t lcs_prefilter_scratch_bytes(int nlev, int nlat, int nlon) {
    if (nlev < 1 || nlat < 1 || nlon < 1) return 0;
    return (size_t)2 * nlev * nlat * nlon * sizeof(double);
}

// Poles of the B-spline prefilter (scipy ni_splines.c:get_filter_poles)
static int spline_poles(int order, double* z) {
    switch (order) {
        case 2: z[0] = sqrt(8.0) - 3.0; return 1;
        case 3: z[0] = sqrt(3.0) - 2.0; return 1;
        case 4: z[0] = sqrt(664.0 - sqrt(438976.0)) + sqrt(304.0) - 19.0;
                z[1] = sqrt(664.0 + sqrt(438976.0)) - sqrt(304.0) - 19.0; return 2;
        case 5: z[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
                z[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5; return 2;
        default: return 0;
    }
}

extern "C" int lcs_prefilter(const void* u, const void* v, int in_dtype, double* coef_u, double* coef_v,
                             void* scratch, size_t scratch_bytes, int nlev, int nlat, int nlon, int order, void* stream) {
    if (!u || !v || !coef_u || !coef_v || !scratch) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: null argument");
    if (nlev < 1 || nlat < 2 || nlon < 2) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: bad sizes");
    if (u == coef_u || v == coef_v) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: outputs may not alias inputs");
    if (scratch_bytes < lcs_prefilter_scratch_bytes(nlev, nlat, nlon))
        return lcs_fail(LCS_E_WORKSPACE, "lcs_prefilter: scratch too small");
    if (2 * nlev > 65535) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: at most 32767 levels per call");
    if (in_dtype != LCS_F64 && in_dtype != LCS_F32) return lcs_fail(LCS_E_INVALID, "lcs_prefilter: bad in_dtype");
    double poles[2];
    const int npoles = spline_poles(order, poles);
    if (npoles == 0) return lcs_fail(LCS_E_UNSUPPORTED, "lcs_prefilter: order must be 2..5 (order 1 needs no prefilter)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* tmp = static_cast<double*>(scratch);
    const dim3 q1((nlon + kIirThreads - 1) / kIirThreads, (nlat + KQ - 1) / KQ, 2 * nlev), q2((nlat + kIirThreads - 1) / kIirThreads, (nlon + KQ - 1) / KQ, 2 * nlev);
    const dim3 g1((nlon + kColThreads - 1) / kColThreads, 1, 2 * nlev), g2((nlat + kColThreads - 1) / kColThreads, 1, 2 * nlev);
    // LCS_PREFILTER_FORM: 0 = column-streaming kernel where kh allows (orders 2, 3) and the launch has enough columns to
    // fill the machine (a column is walked sequentially: one C2 field is 108 blocks of 64 threads and takes 64 us against
    // 44 us in the run form), 1 = independent runs always, 2 = column kernel whenever kh allows
    int form = lcs_env_int("LCS_PREFILTER_FORM", 0);
    if (form == 0 && (long long)g1.x * g1.z < 4LL * lcs_sm_count()) form = 1;
    if (form == 2) form = 0;
    // One pole = one symmetric two-sided exponential h0 z^|k| with unit DC gain.  Per pole: a pass along latitude
    // ([plane][lat][lon] -> scratch [plane][lon][lat]) and a pass along longitude (scratch -> coef [lat][lon]); the
    // second pole of orders 4 and 5 re-reads the coefficient planes.  scipy runs all poles along an axis before the
    // next axis; the passes are linear and separable, so the order only moves the last bits.
    for (int ip = 0; ip < npoles; ++ip) {
        const double z = poles[ip];
        const double h0 = (1.0 - z) * (1.0 - 1.0 / z) * (-z) / (1.0 - z * z);
        int kh = (int)ceil(log(1e-18) / log(fabs(z)));                  // |z|^kh < 1e-18: below one f64 ulp of the sum
        if (kh < 4) kh = 4;
        const void* src_u = ip == 0 ? u : (const void*)coef_u;
        const void* src_v = ip == 0 ? v : (const void*)coef_v;
        const int src_dtype = ip == 0 ? in_dtype : LCS_F64;
        cudaError_t e;
        if (form == 0 && kh <= 2 * CQ) {
            if (src_dtype == LCS_F64)
                iir_column_kernel<double><<<g1, kColThreads, 0, st>>>((const double*)src_u, (const double*)src_v, 1, tmp, nullptr, 0, nlat, nlon, z, h0);
            else
                iir_column_kernel<float><<<g1, kColThreads, 0, st>>>((const float*)src_u, (const float*)src_v, 1, tmp, nullptr, 0, nlat, nlon, z, h0);
            e = cudaGetLastError();
            if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(lat pass)");
            iir_column_kernel<double><<<g2, kColThreads, 0, st>>>(tmp, nullptr, 0, coef_u, coef_v, 1, nlon, nlat, z, h0);
            e = cudaGetLastError();
            if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(lon pass)");
            lcs_count_launches(2);
            continue;
        }
        if (src_dtype == LCS_F64) {
            iir_axis0_transpose_kernel<double><<<q1, kIirThreads, 0, st>>>((const double*)src_u, (const double*)src_v, 1, tmp, nullptr, 0,
                                                                   nlat, nlon, z, h0, kh);
        } else {
            iir_axis0_transpose_kernel<float><<<q1, kIirThreads, 0, st>>>((const float*)src_u, (const float*)src_v, 1, tmp, nullptr, 0,
                                                                  nlat, nlon, z, h0, kh);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(lat pass)");
        iir_axis0_transpose_kernel<double><<<q2, kIirThreads, 0, st>>>(tmp, nullptr, 0, coef_u, coef_v, 1, nlon, nlat, z, h0, kh);
        e = cudaGetLastError();
        if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_prefilter(lon pass)");
        lcs_count_launches(2);
    }
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
    lcs_count_launches(1);
    return LCS_OK;
}

extern "C" int lcs_pack_es(const void* u, const void* v, int in_dtype, void* e_out, void* s_out, int es_dtype,
                           int nlev, int nlat, int nlon, void* stream) {
    if (!u || !v || !e_out || !s_out) return lcs_fail(LCS_E_INVALID, "lcs_pack_es: null argument");
    if (nlev < 2 || nlat < 1 || nlon < 1) return lcs_fail(LCS_E_INVALID, "lcs_pack_es: need at least two levels");
    if (nlat < 4 || nlon < 4) return lcs_fail(LCS_E_INVALID, "lcs_pack_es: grid must be at least 4x4 (single mirror reflection of the halo)");
    if (nlev > 65535 || nlat + LCS_HALO_LO + LCS_HALO_HI > 65535) return lcs_fail(LCS_E_INVALID, "lcs_pack_es: at most 65535 levels / 65530 rows per call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 gb((unsigned)((nlon + LCS_HALO_LO + LCS_HALO_HI + 255) / 256), (unsigned)(nlat + LCS_HALO_LO + LCS_HALO_HI), (unsigned)nlev);
    if (in_dtype == LCS_F64 && es_dtype == LCS_F64)
        pack_es_kernel<double, double><<<gb, 256, 0, st>>>((const double*)u, (const double*)v, (d2*)e_out, (d2*)s_out, nlat, nlon, nlev);
    else if (in_dtype == LCS_F64 && es_dtype == LCS_F32)
        pack_es_kernel<double, float><<<gb, 256, 0, st>>>((const double*)u, (const double*)v, (float2*)e_out, (float2*)s_out, nlat, nlon, nlev);
    else if (in_dtype == LCS_F32 && es_dtype == LCS_F64)
        pack_es_kernel<float, double><<<gb, 256, 0, st>>>((const float*)u, (const float*)v, (d2*)e_out, (d2*)s_out, nlat, nlon, nlev);
    else if (in_dtype == LCS_F32 && es_dtype == LCS_F32)
        pack_es_kernel<float, float><<<gb, 256, 0, st>>>((const float*)u, (const float*)v, (float2*)e_out, (float2*)s_out, nlat, nlon, nlev);
    else return lcs_fail(LCS_E_INVALID, "lcs_pack_es: bad dtype");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_pack_es");
    lcs_count_launches(1);
    return LCS_OK;
}
