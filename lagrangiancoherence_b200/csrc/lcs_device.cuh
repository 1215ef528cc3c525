// Device-side building blocks shared by the integrator, the seams and the microbenchmark.
//
// Everything here restates the *published* algorithm of scipy.ndimage.map_coordinates
// (scipy 1.18.1; the third-party call behind tools.py:26-30,35-39 of the reference) in the
// order scipy evaluates it.  `oracle/lcs_oracle.py` holds the same restatement in numpy and is
// checked bit-for-bit against scipy on the CPU (tests/test_oracle_scipy_spec.py).
//
// STRICT=true  : every product/sum is a separate IEEE operation in scipy's order
//                ((c*wy)*wx accumulated sequentially, true division by 6).
// STRICT=false : same tap order, but weights are pre-multiplied and accumulated with FMA and
//                the /6 becomes *(1/6).  Differences are O(1 ulp) per sample.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lcs_b200.h"

// -DLCS_BOUNDS_CHECK builds (tests/test_gpu_bounds.py, scripts/sanitize_case.py): device-side asserts on every index the
// gathers, the persistent kernel's state / candidate / flag tables and the prefilter form -- compute-sanitizer is closed on
// the measurement pool, so memory safety is checked with bounds checks of our own on ragged cases.  No code otherwise.
#ifdef LCS_BOUNDS_CHECK
#include <assert.h>
#define LCS_ASSERT(c) assert(c)
#else
#define LCS_ASSERT(c) ((void)0)
#endif

namespace lcs {

// ------------------------------------------------------------------ gather element policies
// A policy names the storage element of one grid point (`type`), how many values a tap yields
// (`NV`) and how to load them through the read-only path.
struct __align__(32) d4 { double x, y, z, w; };   // 32-B sector
struct __align__(16) d2 { double x, y; };

template <typename T> struct PairOf;                // 4-wide element (u_k, v_k, u_{k+1}, v_{k+1})
template <> struct PairOf<double> { using type = d4; };
template <> struct PairOf<float> { using type = float4; };
template <typename T> struct Vec2Of;                // 2-wide element
template <> struct Vec2Of<double> { using type = d2; };
template <> struct Vec2Of<float> { using type = float2; };

// all four SETTLS operands of a tap with one load (256-bit LDG.E.256.CONSTANT for f64)
template <typename T> struct Pair4;
template <> struct Pair4<double> {
    using type = d4; static constexpr int NV = 4; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[4]) {
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
    }
};
template <> struct Pair4<float> {
    using type = float4; static constexpr int NV = 4; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[4]) {
        const float4 t = __ldg(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
};
// first half (u_k, v_k) of a 4-wide element: the Euler stage samples one level (trajectory.py:82-84)
template <typename T> struct Pair4Lo;
template <> struct Pair4Lo<double> {
    using type = d4; static constexpr int NV = 2; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[2]) {
        asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "l"(p));
    }
};
template <> struct Pair4Lo<float> {
    using type = float4; static constexpr int NV = 2; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[2]) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        o[0] = t.x; o[1] = t.y;
    }
};
// 2-wide element (the E / S arrays of the fast layout): 16 B (f64) or 8 B (f32) per tap.  HALO: every level is
// stored with a mirror-filled halo (LCS_HALO_LO cells before, LCS_HALO_HI after each axis; lcs_pack_es), so a tap
// index never has to be reflected and every gather takes the one-base-address path.  That matters more than it
// sounds: the boundary clamps park a large share of the particles exactly on the domain edge, where the 4x4
// stencil sticks out of the grid -- ncu (round 2, C2 outer clamp) attributed 30 % of the integrator's executed
// instructions to the reflected-index path of the dense layout.
template <typename T> struct Vec2;
template <> struct Vec2<double> {
    using type = d2; static constexpr int NV = 2; static constexpr bool A32 = false, HALO = true;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[2]) {
        asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "l"(p));
    }
};
template <> struct Vec2<float> {
    using type = float2; static constexpr int NV = 2; static constexpr bool A32 = false, HALO = true;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[2]) {
        const float2 t = __ldg(p);
        o[0] = t.x; o[1] = t.y;
    }
};
// f32 elements whose cubic taps are also weighted and accumulated in f32 (LCS_ARITH_F32, the fast path):
// same storage and loads as Vec2<float>; the cubic gather is gather_cubic_wrap_f32 below
struct Vec2F32Arith {
    using type = float2; static constexpr int NV = 2; static constexpr bool A32 = true, HALO = true;
    static __device__ __forceinline__ void ld(const type* p, double (&o)[2]) {
        const float2 t = __ldg(p);
        o[0] = t.x; o[1] = t.y;
    }
};
// planar f64 field, one value per tap (the map_coordinates seam)
struct Scalar32 {
    using type = float; static constexpr int NV = 1; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const float* p, double (&o)[1]) { o[0] = (double)__ldg(p); }
};
struct Scalar64 {
    using type = double; static constexpr int NV = 1; static constexpr bool A32 = false, HALO = false;
    static __device__ __forceinline__ void ld(const double* p, double (&o)[1]) { o[0] = __ldg(p); }
};

// row pitch (elements) of a level of policy E over an nlat x nlon grid; HALO levels are addressed from their
// (0, 0) element, i.e. the caller passes base + LCS_HALO_LO * pitch + LCS_HALO_LO
template <typename E> __device__ __forceinline__ int level_pitch(int nlon) {
    return E::HALO ? nlon + LCS_HALO_LO + LCS_HALO_HI : nlon;
}

// ------------------------------------------------------------------ index arithmetic
// tools.py:21-22: n * (pos - cmin) / (cmax - cmin)   (n points, not n-1: quirk Q4)
__device__ __forceinline__ double index_map(double pos, double cmin, double span, double n) {
    return __ddiv_rn(__dmul_rn(n, __dsub_rn(pos, cmin)), span);
}
// fast variant: one multiply by the precomputed n/span (differs from the above by <= 1 ulp)
__device__ __forceinline__ double index_map_fast(double pos, double cmin, double n_over_span) {
    return (pos - cmin) * n_over_span;
}

// scipy mode='wrap' coordinate fold: period n-1, result in [0, n-1]
// Within one period of the range the truncated quotient scipy forms is 0 (below) or 1 (above), so the fold is one
// exact add; that is every coordinate a clamped particle can have (index n at x = x_max: quirk Q4), and it skips
// the f64 division (a ~30-instruction sequence) the general case needs.
__device__ __forceinline__ double fold_wrap(double c, int n) {
    const double szd = (double)(n - 1);
    if (c < 0.0) {
        if (-c < szd) return __dadd_rn(c, szd);
        const long long sz = n - 1;
        c += (double)(sz * ((long long)(__ddiv_rn(-c, szd)) + 1));
    } else if (c > szd) {
        if (c < 2.0 * szd) return __dsub_rn(c, szd);
        const long long sz = n - 1;
        c -= (double)(sz * (long long)(__ddiv_rn(c, szd)));
    }
    return c;
}

// scipy's spline_mode mirror for out-of-range tap indices is d c b | a b c d | c b a.
// The indices a gather can actually produce: coordinates are folded into [0, n-1] first (or rejected by
// the 'constant' branch), so taps lie in [-1, n+1] (orders 1, 3) or [-2, n+2] (orders 2, 4, 5); for n >= 4 a
// single reflection needs no division: -1 -> 1, n -> n-2, n+2 -> n-4 (scipy's general formula,
// oracle.mirror_index, reduces to this).
__device__ __forceinline__ int mirror_near(int i, int n) {
    return i < 0 ? -i : (i > n - 1 ? 2 * (n - 1) - i : i);
}

template <bool STRICT>
__device__ __forceinline__ void cubic_weights(double y, double (&w)[4]) {
    const double z = __dsub_rn(1.0, y);
    if (STRICT) {
        w[1] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(y, y), __dsub_rn(y, 2.0)), 3.0), 4.0), 6.0);
        w[2] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(z, z), __dsub_rn(z, 2.0)), 3.0), 4.0), 6.0);
        w[0] = __ddiv_rn(__dmul_rn(__dmul_rn(z, z), z), 6.0);
        w[3] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, w[0]), w[1]), w[2]);
    } else {
        // the same cubics in Horner form, 11 operations per axis instead of 17: w1 = (y/2 - 1) y^2 + 2/3 (likewise w2 in
        // z), w0 = z^3 / 6, w3 = y^3 / 6 (scipy forms w3 = 1 - w0 - w1 - w2; equal up to rounding)
        const double s = 1.0 / 6.0, c23 = 2.0 / 3.0;
        const double ty = y * y, tz = z * z;
        w[1] = fma(fma(0.5, y, -1.0), ty, c23);
        w[2] = fma(fma(0.5, z, -1.0), tz, c23);
        w[0] = (z * s) * tz;
        w[3] = (y * s) * ty;
    }
}

template <bool STRICT>
__device__ __forceinline__ double tap_acc(double t, double c, double wy, double wx, double wyx) {
    if (STRICT) return __dadd_rn(t, __dmul_rn(__dmul_rn(c, wy), wx));
    return fma(c, wyx, t);
}

// ------------------------------------------------------------------ gathers
// Cubic B-spline, mode='wrap' (fold period n-1, mirror taps), 4x4 taps in scipy's order
// (axis-0 index outer, axis-1 inner).  Interior positions (all 16 taps inside the grid, the
// overwhelmingly common case) take a path with one base address and immediate offsets; the
// general path reflects every tap index.
template <typename E, bool STRICT>
__device__ __forceinline__ void gather_cubic_wrap(const typename E::type* __restrict__ f,
                                                  int nlat, int nlon, double iy, double ix,
                                                  double (&out)[E::NV]) {
    constexpr int NV = E::NV;
    const double cy = fold_wrap(iy, nlat);
    const double cx = fold_wrap(ix, nlon);
    const double fy = floor(cy), fx = floor(cx);
    double wy[4], wx[4];
    cubic_weights<STRICT>(__dsub_rn(cy, fy), wy);
    cubic_weights<STRICT>(__dsub_rn(cx, fx), wx);
    const int sy = (int)fy - 1, sx = (int)fx - 1;
    const int pitch = level_pitch<E>(nlon);
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = 0.0;
    if (E::HALO || (sy >= 0 && sy + 3 < nlat && sx >= 0 && sx + 3 < nlon)) {
        LCS_ASSERT(!E::HALO || (sy >= -LCS_HALO_LO && sy + 3 <= nlat - 1 + LCS_HALO_HI && sx >= -LCS_HALO_LO && sx + 3 <= nlon - 1 + LCS_HALO_HI));
        const typename E::type* base = f + (sy * pitch + sx);               // a level has < 2^31 elements
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double c[4][NV];
#pragma unroll
            for (int j = 0; j < 4; ++j) E::ld(base + j, c[j]);
            if (STRICT) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) out[v] = tap_acc<true>(out[v], c[j][v], wy[i], wx[j], 0.0);
                }
            } else {
                // row sums first, then the latitude weight: 5 operations per row and value instead of 6 (no weight
                // products) and four short dependency chains instead of one of 16 FMAs
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    double r = c[0][v] * wx[0];
                    r = fma(c[1][v], wx[1], r);
                    r = fma(c[2][v], wx[2], r);
                    r = fma(c[3][v], wx[3], r);
                    out[v] = i == 0 ? r * wy[0] : fma(r, wy[i], out[v]);
                }
            }
            base += pitch;
        }
        return;
    }
    // edge path: every tap index reflected; fully unrolled so the weight arrays stay in registers
    // (a rolled loop indexed them dynamically and put 4 local-memory stores into every stage: ncu showed
    // local-memory wavefronts at 22-48 % of the global-load wavefronts on the L1 pipe that bounds the kernel)
    int col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) col[j] = mirror_near(sx + j, nlon);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const typename E::type* rowp = f + (size_t)mirror_near(sy + i, nlat) * nlon;
        double c[4][NV];
#pragma unroll
        for (int j = 0; j < 4; ++j) E::ld(rowp + col[j], c[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double wyx = wy[i] * wx[j];
#pragma unroll
            for (int v = 0; v < NV; ++v) out[v] = tap_acc<STRICT>(out[v], c[j][v], wy[i], wx[j], wyx);
        }
    }
}

// ------------------------------------------------------------------ spline orders 2, 4, 5
// `traj_interp_order` is free upstream (whatever scipy accepts); 3 is the default and 1 the other value any
// caller uses, so those two have hand-tuned gathers above.  Orders 2, 4 and 5 share this generic form:
// scipy's weights (ni_splines.c:get_spline_interpolation_weights) in scipy's order of operations, first tap
// floor(c) - order/2 (odd) or floor(c + 0.5) - order/2 (even), (order+1)^2 mirrored taps.
template <bool ST> __device__ __forceinline__ double op_mul(double a, double b) { return ST ? __dmul_rn(a, b) : a * b; }
template <bool ST> __device__ __forceinline__ double op_add(double a, double b) { return ST ? __dadd_rn(a, b) : a + b; }
template <bool ST> __device__ __forceinline__ double op_sub(double a, double b) { return ST ? __dsub_rn(a, b) : a - b; }
template <bool ST> __device__ __forceinline__ double op_div(double a, double b) { return ST ? __ddiv_rn(a, b) : a * (1.0 / b); }

template <int ORDER, bool ST>
__device__ __forceinline__ int spline_weights(double c, double (&w)[ORDER + 1]) {
    static_assert(ORDER == 2 || ORDER == 4 || ORDER == 5, "orders 1 and 3 have their own gathers");
    const double f = (ORDER & 1) ? floor(c) : floor(__dadd_rn(c, 0.5));
    const double x = __dsub_rn(c, f);
    double y = x, z = __dsub_rn(1.0, x), t;
    if (ORDER == 2) {
        w[1] = op_sub<ST>(0.75, op_mul<ST>(x, x));
        y = op_sub<ST>(0.5, x);
        w[0] = op_mul<ST>(op_mul<ST>(0.5, y), y);
    } else if (ORDER == 4) {
        t = op_mul<ST>(x, x);
        w[2] = op_add<ST>(op_mul<ST>(t, op_sub<ST>(op_mul<ST>(t, 0.25), 0.625)), 115.0 / 192.0);
        y = op_add<ST>(1.0, x);
        w[1] = op_add<ST>(op_mul<ST>(y, op_add<ST>(op_mul<ST>(y, op_sub<ST>(op_div<ST>(op_mul<ST>(y, op_sub<ST>(5.0, y)), 6.0), 1.25)), 5.0 / 24.0)), 55.0 / 96.0);
        w[3] = op_add<ST>(op_mul<ST>(z, op_add<ST>(op_mul<ST>(z, op_sub<ST>(op_div<ST>(op_mul<ST>(z, op_sub<ST>(5.0, z)), 6.0), 1.25)), 5.0 / 24.0)), 55.0 / 96.0);
        y = op_sub<ST>(0.5, x);
        t = op_mul<ST>(y, y);
        w[0] = op_div<ST>(op_mul<ST>(t, t), 24.0);
    } else {
        t = op_mul<ST>(y, y);
        w[2] = op_add<ST>(op_mul<ST>(t, op_sub<ST>(op_mul<ST>(t, op_sub<ST>(0.25, op_div<ST>(y, 12.0))), 0.5)), 0.55);
        t = op_mul<ST>(z, z);
        w[3] = op_add<ST>(op_mul<ST>(t, op_sub<ST>(op_mul<ST>(t, op_sub<ST>(0.25, op_div<ST>(z, 12.0))), 0.5)), 0.55);
        y = op_add<ST>(y, 1.0);
        w[1] = op_add<ST>(op_mul<ST>(y, op_add<ST>(op_mul<ST>(y, op_sub<ST>(op_mul<ST>(y, op_add<ST>(op_mul<ST>(y, op_sub<ST>(op_div<ST>(y, 24.0), 0.375)), 1.25)), 1.75)), 0.625)), 0.425);
        z = op_add<ST>(z, 1.0);
        w[4] = op_add<ST>(op_mul<ST>(z, op_add<ST>(op_mul<ST>(z, op_sub<ST>(op_mul<ST>(z, op_add<ST>(op_mul<ST>(z, op_sub<ST>(op_div<ST>(z, 24.0), 0.375)), 1.25)), 1.75)), 0.625)), 0.425);
        y = op_sub<ST>(1.0, x);
        t = op_mul<ST>(y, y);
        w[0] = op_div<ST>(op_mul<ST>(op_mul<ST>(y, t), t), 120.0);
    }
    double last = 1.0;
#pragma unroll
    for (int i = 0; i < ORDER; ++i) last = __dsub_rn(last, w[i]);
    w[ORDER] = last;
    return (int)f - ORDER / 2;
}

template <typename E, bool STRICT, int ORDER>
__device__ __forceinline__ void gather_spline_wrap(const typename E::type* __restrict__ f,
                                                   int nlat, int nlon, double iy, double ix,
                                                   double (&out)[E::NV]) {
    constexpr int NV = E::NV, NT = ORDER + 1;
    const double cy = fold_wrap(iy, nlat);
    const double cx = fold_wrap(ix, nlon);
    double wy[NT], wx[NT];
    const int sy = spline_weights<ORDER, STRICT>(cy, wy);
    const int sx = spline_weights<ORDER, STRICT>(cx, wx);
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = 0.0;
    const int pitch = level_pitch<E>(nlon);
    if (E::HALO || (sy >= 0 && sy + ORDER < nlat && sx >= 0 && sx + ORDER < nlon)) {
        LCS_ASSERT(!E::HALO || (sy >= -LCS_HALO_LO && sy + ORDER <= nlat - 1 + LCS_HALO_HI && sx >= -LCS_HALO_LO && sx + ORDER <= nlon - 1 + LCS_HALO_HI));
        const typename E::type* base = f + (sy * pitch + sx);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            double c[NT][NV];
#pragma unroll
            for (int j = 0; j < NT; ++j) E::ld(base + j, c[j]);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double wyx = wy[i] * wx[j];
#pragma unroll
                for (int v = 0; v < NV; ++v) out[v] = tap_acc<STRICT>(out[v], c[j][v], wy[i], wx[j], wyx);
            }
            base += pitch;
        }
        return;
    }
    int col[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) col[j] = mirror_near(sx + j, nlon);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const typename E::type* rowp = f + (size_t)mirror_near(sy + i, nlat) * nlon;
        double c[NT][NV];
#pragma unroll
        for (int j = 0; j < NT; ++j) E::ld(rowp + col[j], c[j]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double wyx = wy[i] * wx[j];
#pragma unroll
            for (int v = 0; v < NV; ++v) out[v] = tap_acc<STRICT>(out[v], c[j][v], wy[i], wx[j], wyx);
        }
    }
}

// Fast-path cubic gather (LCS_ARITH_F32): index map, fold and the fractional offsets stay in f64 (an f32 index
// of magnitude ~10^3 would carry 1e-4 cells of error into the weights); the 16 products and the accumulation are f32,
// two values (u, v) per packed FFMA2, row sums first, then the latitude weights.
// ANOMALY FORM (round 2): the taps are differenced against the stencil's central tap c11 before they are weighted,
//   sample = c11 + sum_ij w_ij (c_ij - c11)        (the weights sum to one),
// so every f32 rounding -- of a weight, a product, a partial sum -- is relative to the local VARIATION of the field
// over four cells instead of its magnitude, and the large term c11 enters once, exactly, in f64.  Numpy emulation of
// this arithmetic (scripts/proto_f32fast.py) on the tolerance test's case: FTLE within 1e-5 at 97.9 % of the points
// against 90.7 % for plain f32 sums and 98.2 % for f64 arithmetic on the same f32 coefficients (the storage rounding
// is what remains).  The weights are evaluated in f64 (Horner, as the f64 path) and rounded once: with the anomaly
// form their error no longer multiplies the field's magnitude, and f64 weights are worth another 0.8 % of the points.
// Cost over the plain form: 16 packed subtractions, 8 + 4 conversions per sample.
#ifndef LCS_F32_WEIGHTS64
#define LCS_F32_WEIGHTS64 1
#endif
__device__ __forceinline__ void cubic_weights_f32(float y, float2 (&w)[4]) {
    const float z = 1.0f - y;
    const float s = 1.0f / 6.0f;
    const float w1 = (y * y * (y - 2.0f) * 3.0f + 4.0f) * s;
    const float w2 = (z * z * (z - 2.0f) * 3.0f + 4.0f) * s;
    const float w0 = z * z * z * s;
    const float w3 = 1.0f - w0 - w1 - w2;
    w[0] = make_float2(w0, w0); w[1] = make_float2(w1, w1); w[2] = make_float2(w2, w2); w[3] = make_float2(w3, w3);
}
__device__ __forceinline__ void cubic_weights_f32_from_f64(double y, float2 (&w)[4]) {
    double wd[4];
    cubic_weights<false>(y, wd);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float t = __double2float_rn(wd[i]); w[i] = make_float2(t, t); }
}

__device__ __forceinline__ void gather_cubic_wrap_f32(const float2* __restrict__ f, int nlat, int nlon,
                                                      double iy, double ix, double (&out)[2]) {
    const double cy = fold_wrap(iy, nlat);
    const double cx = fold_wrap(ix, nlon);
    const double fy = floor(cy), fx = floor(cx);
    float2 wy[4], wx[4];
#if LCS_F32_WEIGHTS64
    cubic_weights_f32_from_f64(cy - fy, wy);
    cubic_weights_f32_from_f64(cx - fx, wx);
#else
    cubic_weights_f32((float)(cy - fy), wy);
    cubic_weights_f32((float)(cx - fx), wx);
#endif
    const int sy = (int)fy - 1, sx = (int)fx - 1;
    const int pitch = nlon + LCS_HALO_LO + LCS_HALO_HI;                      // halo layout: no tap index is ever reflected
    const float2* base = f + (sy * pitch + sx);
    (void)nlat;
    LCS_ASSERT(sy >= -LCS_HALO_LO && sy + 3 <= nlat - 1 + LCS_HALO_HI && sx >= -LCS_HALO_LO && sx + 3 <= nlon - 1 + LCS_HALO_HI);
    float2 c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = __ldg(base + i * pitch + j);
    }
    const float2 ref = c[1][1];
    const float2 nref = make_float2(-ref.x, -ref.y);
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 r = __fmul2_rn(__fadd2_rn(c[i][0], nref), wx[0]);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            if (i == 1 && j == 1) continue;                                   // c11 - c11 = 0
            r = __ffma2_rn(__fadd2_rn(c[i][j], nref), wx[j], r);
        }
        acc = i == 0 ? __fmul2_rn(r, wy[0]) : __ffma2_rn(r, wy[i], acc);
    }
    out[0] = (double)ref.x + (double)acc.x; out[1] = (double)ref.y + (double)acc.y;
}

// Shared 2x2 tap sum of the order-1 branches (weights (1-y, 1-(1-y)), mirror taps).
template <typename E, bool STRICT>
__device__ __forceinline__ void bilinear_taps(const typename E::type* __restrict__ f,
                                              int nlat, int nlon, double cy, double cx,
                                              double (&out)[E::NV]) {
    constexpr int NV = E::NV;
    const double fy = floor(cy), fx = floor(cx);
    const double y = __dsub_rn(cy, fy), x = __dsub_rn(cx, fx);
    // scipy: weights[0] = 1 - x; weights[order] = 1 - sum(others)
    const double wy0 = __dsub_rn(1.0, y), wx0 = __dsub_rn(1.0, x);
    const double wy[2] = {wy0, __dsub_rn(1.0, wy0)};
    const double wx[2] = {wx0, __dsub_rn(1.0, wx0)};
    const int iy0 = (int)fy, ix0 = (int)fx;
    LCS_ASSERT(iy0 >= (E::HALO ? -LCS_HALO_LO : 0) && iy0 + 1 <= nlat - 1 + (E::HALO ? LCS_HALO_HI : 1) && ix0 >= (E::HALO ? -LCS_HALO_LO : 0) && ix0 + 1 <= nlon - 1 + (E::HALO ? LCS_HALO_HI : 1));
    const int pitch = level_pitch<E>(nlon);
    const int col[2] = {E::HALO ? ix0 : mirror_near(ix0, nlon), E::HALO ? ix0 + 1 : mirror_near(ix0 + 1, nlon)};
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const typename E::type* rowp = f + (E::HALO ? iy0 + i : mirror_near(iy0 + i, nlat)) * pitch;
        double c[2][NV];
#pragma unroll
        for (int j = 0; j < 2; ++j) E::ld(rowp + col[j], c[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double wyx = wy[i] * wx[j];
#pragma unroll
            for (int v = 0; v < NV; ++v) out[v] = tap_acc<STRICT>(out[v], c[j][v], wy[i], wx[j], wyx);
        }
    }
}

// order=1, mode='wrap' (interior rows when interp_order == 1)
template <typename E, bool STRICT>
__device__ __forceinline__ void gather_linear_wrap(const typename E::type* __restrict__ f,
                                                   int nlat, int nlon, double iy, double ix,
                                                   double (&out)[E::NV]) {
    bilinear_taps<E, STRICT>(f, nlat, nlon, fold_wrap(iy, nlat), fold_wrap(ix, nlon), out);
}

// order=1, mode='constant', cval=0: coordinates outside [0, n-1] sample 0 (tools.py:35-39)
template <typename E, bool STRICT>
__device__ __forceinline__ void gather_linear_constant(const typename E::type* __restrict__ f,
                                                       int nlat, int nlon, double iy, double ix,
                                                       double (&out)[E::NV]) {
    // written so that NaN coordinates also take the constant branch
    if (!(iy >= 0.0 && iy <= (double)(nlat - 1) && ix >= 0.0 && ix <= (double)(nlon - 1))) {
#pragma unroll
        for (int v = 0; v < E::NV; ++v) out[v] = 0.0;
        return;
    }
    bilinear_taps<E, STRICT>(f, nlat, nlon, iy, ix, out);
}

// ------------------------------------------------------------------ boundaries (trajectory.py:89-97)
__device__ __forceinline__ double clamp_y(double y, double ymin, double ymax) {
    y = (y > ymin) ? y : ymin;      // .where(y > y_min, y_min): NaN -> y_min
    y = (y < ymax) ? y : ymax;
    return y;
}

// numpy float `%` (sign of the divisor)
__device__ __forceinline__ double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if ((b < 0.0) != (r < 0.0)) r = __dadd_rn(r, b); }
    else r = copysign(0.0, b);
    return r;
}

__device__ __forceinline__ double wrap_x_cyclic(double x) {
    x = (x > -180.0) ? x : pymod(x, 180.0);                       // trajectory.py:93
    x = (x < 180.0) ? x : __dadd_rn(-180.0, pymod(x, 180.0));     // trajectory.py:94
    return x;
}

__device__ __forceinline__ double clamp_x_pointwise(double x, double xmin, double xmax) {
    if (x < xmin) x = xmin;
    if (x > xmax) x = xmax;
    return x;
}

}  // namespace lcs
