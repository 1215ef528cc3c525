// Departure-point integrator: parcel_propagation's loop (trajectory.py:80-126) with the
// xr_map_coordinates calls inside it (tools.py:11-41) as CUDA kernels for sm_100a.
//
// Launch shapes, all built from the same stage functions (stage_euler / stage_settls):
//  * advect_fused_kernel        -- one thread per particle carries it across every wind interval and SETTLS
//    sub-iteration in registers (cyclic / pointwise x-boundary, where particles are independent);
//  * advect_outer_group_kernel  -- the as-executed outer-product x-clamp (quirk Q6) couples all particles of a
//    window after every sub-step: one cooperative grid of resident CTAs in groups, a group per window in flight,
//    group barriers in global memory between the sub-steps, optional exchange of the column flags with other GPUs;
//  * advect_phase_*             -- the same clamp with one launch pair per sub-step (kernel boundaries as barriers):
//    the independent implementation the tests compare the persistent kernel with.
//
// Data, two layouts (include/lcs_b200.h):
//  * PAIR4: pairs[k][lat][lon] = (u_k, v_k, u_{k+1}, v_{k+1}); one 32-B (f64) vector load per tap feeds the four
//    operands of a SETTLS stage, combined in the reference's order (bit-faithful `strict` evaluation);
//  * ES   : E[k] = (u_k, v_k) and S[k] = (2u_k - u_{k+1}, 2v_k - v_{k+1}), every level with a mirror-filled halo; a
//    SETTLS stage needs one 16-B load per tap and no gather ever reflects a tap index.
// The kernels are bound by L1 data-pipe wavefronts (ncu, round 2: 82-85 % of peak, issue slots 48 %, DRAM 41 % for
// the outer-clamp kernel): gathers go through the read-only L1 path, 16 B per tap.  A block-level shared-memory /
// TMA tile of the taps was built in round 1 and lost to its barriers; what paid in round 2 was
// removing instructions and wavefronts around the gathers (halo layout, fold fast path, row-sum taps).
#pragma once
#include <stdio.h>
#include <mutex>
#include "lcs_internal.h"
#include "lcs_device.cuh"

namespace lcs {

struct AdvectParams {
    // winds: PAIR4 layout -> raw_a/coef_a = packed pairs; ES layout -> *_a = E levels, *_b = S intervals
    const void* raw_a;
    const void* raw_b;
    int raw_planar;               // ES layout, orders >= 2: raw_a / raw_b are the planar u / v series themselves
    int raw_f32;                  //   ... stored as f32 (else f64)
    const void* coef_a;
    const void* coef_b;
    size_t plane;                 // nlat*nlon (elements per level of the dense arrays: PAIR4 pairs, planar raw winds)
    size_t plane_es;              // elements per level of the ES arrays (halo layout, include/lcs_b200.h)
    int halo_off;                 // offset of grid point (0, 0) inside an ES level
    int nlat, nlon;
    double nlat_d, nlon_d;
    double lat_min, lat_span, lon_min, lon_span, lat_max, lon_max;
    double nlat_over_span, nlon_over_span;
    // particles
    int nrow, ncol, row0, nrow_global;
    int np;                       // nrow*ncol (< 2^31)
    const double* lat;
    const double* lon;
    const double* kx;
    const double* hx;
    double ky, hy;
    float ky32, hy32;             // R32 with a weak (Python) timestep: the y increments are formed in f32 (NEP 50)
    int y_weak;
    int nsteps, S, xmode, level0, level_stride, band, band_log2;
    double* x_out;
    double* y_out;
    double* x_traj;
    double* y_traj;
    // phased (outer clamp) state, indexed by the particle's enumeration index p
    d2* spos;                     // [nwindows][np] (x, y)
    d2* swind;                    // [nwindows][np] (ua, va) of the interval's Euler stage
    int* cand;                    // [nwindows][np] particles with x > lon_max after the current sub-step
    int* cand_count;              // [nwindows][nsub]
    unsigned char* flags;         // [nwindows][nsub][2][nrow+ncol]
    int nsub;
    const lcs_xrank* xr_host;              // host pointer (launcher only): cross-rank exchange of the outer clamp, or NULL
    int nslots, ntc;                       // persistent kernel: slots per window (tiles of 2x16), tile columns
    unsigned ntc_magic; int ntc_shift;     // tile / ntc as a multiply-high, see slot_rc
};

constexpr int kPair4 = LCS_LAYOUT_PAIR4, kES = LCS_LAYOUT_ES;
constexpr int kES32 = 2;     // internal: ES layout of f32 elements with the cubic taps evaluated in f32 (LCS_ARITH_F32)
template <typename T, int LAYOUT> struct EsPolicy { using type = Vec2<T>; };
template <> struct EsPolicy<float, kES32> { using type = Vec2F32Arith; };

// Thread -> particle: a block is a (band x blockDim/band) tile of the particle grid, a warp a
// (band x 32/band) patch (smaller unique tap footprint than a 1x32 strip: better L1 hit rate);
// grid = (column tiles, row bands, windows), so no integer division is needed.
// STRIP: block = 8 rows x 32 columns, warp = one row of 32 particles with the block's warps stacked (they share tap
// rows).  Measured against the 2 x 16 patches: the same at C2 (14.3 ms), 4 % slower with f32 taps, 3 % faster with hourly
// winds (C4) and 18-21 % faster on the 721 x 1440 grid (C3: 3.17 vs 4.00 ms for four 12-interval windows), where the
// compact block keeps its taps in L1 while the 2 x 128 strip of the default mapping does not.  Chosen by launch_advect.
template <bool STRIP = false>
__device__ __forceinline__ bool particle_rc(const AdvectParams& P, int& row, int& col) {
    if (STRIP) {
        row = blockIdx.y * 8 + (threadIdx.x >> 5);
        col = blockIdx.x * 32 + (threadIdx.x & 31);
        return row < P.nrow && col < P.ncol;
    }
    const int r = threadIdx.x & (P.band - 1);
    const int c = threadIdx.x >> P.band_log2;
    row = blockIdx.y * P.band + r;
    col = blockIdx.x * (blockDim.x >> P.band_log2) + c;          // blocks of 256, 128 or 64 threads (launch_advect)
    return row < P.nrow && col < P.ncol;
}

// One xr_map_coordinates evaluation (tools.py:19-41) of the element policy E at (x, y):
// index map, then the order-1/'constant' branch on pole rows, else order ORDER/'wrap'.
template <typename E, bool STRICT, int ORDER>
__device__ __forceinline__ void sample(const AdvectParams& P, const void* raw, const void* coef, int level,
                                       bool pole, double x, double y, double (&out)[E::NV]) {
    using ET = typename E::type;
    double iy, ix;
    if (STRICT) {
        iy = index_map(y, P.lat_min, P.lat_span, P.nlat_d);
        ix = index_map(x, P.lon_min, P.lon_span, P.nlon_d);
    } else {
        iy = index_map_fast(y, P.lat_min, P.nlat_over_span);
        ix = index_map_fast(x, P.lon_min, P.nlon_over_span);
    }
    const size_t lvl = E::HALO ? (size_t)level * P.plane_es + P.halo_off : (size_t)level * P.plane;
    if (pole) {
        gather_linear_constant<E, STRICT>(reinterpret_cast<const ET*>(raw) + lvl, P.nlat, P.nlon, iy, ix, out);
    } else if (ORDER == 3) {
        if constexpr (E::A32) gather_cubic_wrap_f32(reinterpret_cast<const ET*>(coef) + lvl, P.nlat, P.nlon, iy, ix, out);
        else gather_cubic_wrap<E, STRICT>(reinterpret_cast<const ET*>(coef) + lvl, P.nlat, P.nlon, iy, ix, out);
    } else if (ORDER == 1) {
        gather_linear_wrap<E, STRICT>(reinterpret_cast<const ET*>(raw) + lvl, P.nlat, P.nlon, iy, ix, out);
    } else {
        if constexpr (ORDER == 2 || ORDER == 4 || ORDER == 5)
            gather_spline_wrap<E, STRICT, ORDER>(reinterpret_cast<const ET*>(coef) + lvl, P.nlat, P.nlon, iy, ix, out);
    }
}

// Pole rows of the spline orders (>= 2) sample the winds themselves with order 1 / 'constant' (tools.py:31-39).
// Those are 2*ORDER particle rows out of hundreds.  With `raw_planar` the ES layout does not pack a second copy of
// the series for them: they read the planar u, v input directly -- tap offsets and weights once, then four scalar
// taps per field.  SETTLS = true returns 2 f_k(pos) - f_{k+1}(pos) from two separate samples, which is the
// reference's own order (trajectory.py:105-112).  Saves a third of the staging traffic at no measurable integrator
// cost (a first version that called the generic bilinear gather four times cost 2.7 %).
template <typename TR, bool SETTLS, bool STRICT>
__device__ __forceinline__ void pole_taps_planar(const TR* __restrict__ u, const TR* __restrict__ v, size_t plane,
                                                 const int (&off)[4], const double (&wy)[2], const double (&wx)[2], double (&out)[4]) {
    double su = 0.0, sv = 0.0, su1 = 0.0, sv1 = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double a = wy[t >> 1], b = wx[t & 1], ab = a * b;          // scipy: t += (value * wy) * wx, axis-1 tap inner
        su = tap_acc<STRICT>(su, (double)__ldg(u + off[t]), a, b, ab);
        sv = tap_acc<STRICT>(sv, (double)__ldg(v + off[t]), a, b, ab);
        if (SETTLS) {
            su1 = tap_acc<STRICT>(su1, (double)__ldg(u + plane + off[t]), a, b, ab);
            sv1 = tap_acc<STRICT>(sv1, (double)__ldg(v + plane + off[t]), a, b, ab);
        }
    }
    out[0] = su; out[1] = sv; out[2] = su1; out[3] = sv1;
}

// out = (u_k, v_k, u_{k+1}, v_{k+1}) at (x, y); the level k+1 samples only with SETTLS
template <bool SETTLS, bool STRICT = false>
__device__ __forceinline__ void pole_sample_planar(const AdvectParams& P, int k, double x, double y, double (&out)[4]) {
    const double iy = STRICT ? index_map(y, P.lat_min, P.lat_span, P.nlat_d) : index_map_fast(y, P.lat_min, P.nlat_over_span);
    const double ix = STRICT ? index_map(x, P.lon_min, P.lon_span, P.nlon_d) : index_map_fast(x, P.lon_min, P.nlon_over_span);
    out[0] = 0.0; out[1] = 0.0; out[2] = 0.0; out[3] = 0.0;
    if (!(iy >= 0.0 && iy <= (double)(P.nlat - 1) && ix >= 0.0 && ix <= (double)(P.nlon - 1))) return;   // mode='constant', cval 0
    const double fy = floor(iy), fx = floor(ix);
    const double wy0 = __dsub_rn(1.0, __dsub_rn(iy, fy)), wx0 = __dsub_rn(1.0, __dsub_rn(ix, fx));
    const double wy[2] = {wy0, __dsub_rn(1.0, wy0)}, wx[2] = {wx0, __dsub_rn(1.0, wx0)};   // scipy: last weight = 1 - sum(others)
    const int r0 = (int)fy, c0 = (int)fx;
    const int r1 = mirror_near(r0 + 1, P.nlat), c1 = mirror_near(c0 + 1, P.nlon);
    LCS_ASSERT(r0 >= 0 && r0 < P.nlat && r1 >= 0 && r1 < P.nlat && c0 >= 0 && c0 < P.nlon && c1 >= 0 && c1 < P.nlon);
    const int off[4] = {r0 * P.nlon + c0, r0 * P.nlon + c1, r1 * P.nlon + c0, r1 * P.nlon + c1};
    const size_t o = (size_t)k * P.plane;
    if (P.raw_f32) pole_taps_planar<float, SETTLS, STRICT>(static_cast<const float*>(P.raw_a) + o, static_cast<const float*>(P.raw_b) + o, P.plane, off, wy, wx, out);
    else pole_taps_planar<double, SETTLS, STRICT>(static_cast<const double*>(P.raw_a) + o, static_cast<const double*>(P.raw_b) + o, P.plane, off, wy, wx, out);
}

// R32 -- the reference's dtype propagation for f32 winds on f64 coordinates (how ERA5 is stored).  scipy allocates
// map_coordinates' output with the INPUT dtype (tools.py:26-30,35-39): every sample is computed in f64 and rounded to
// f32.  numpy (NEP 50) then forms `timestep * conversion_y * va` and `0.5 * timestep * conversion_y * (va + 2 v_t - v_t1)`
// in f32 when `timestep` is a Python scalar (trajectory.py:86,110; f64 when it is a numpy scalar, e.g. after
// resample=), while the x increments multiply the f64 array conversion_x and are f64 (trajectory.py:87,112); the
// bracket of the SETTLS increment is f32 arithmetic on the separately rounded samples either way.
// R32 kernels are instantiated with STRICT = true: order-1 samples of f32 data at positions whose index fractions are
// small rationals (every particle starts on a grid node: fractions j/(n-1), quirk Q4) land EXACTLY on f32 rounding ties
// a few per cent of the time, and the tie is then broken by the last bit of the f64 sum -- so the taps must be
// accumulated in scipy's own order ((value * wy) * wx, sequentially) for the rounded sample to agree.
__device__ __forceinline__ double r32(double v) { return (double)__double2float_rn(v); }
__device__ __forceinline__ float settls_bracket32(double a, double ft, double ft1) {          // a + 2 f_t - f_t1 in f32
    return __fsub_rn(__fadd_rn((float)a, __fmul_rn(2.0f, __double2float_rn(ft))), __double2float_rn(ft1));
}
__device__ __forceinline__ double y_increment32(const AdvectParams& P, bool half, float v) {
    if (P.y_weak) return (double)__fmul_rn(half ? P.hy32 : P.ky32, v);
    return __dmul_rn(half ? P.hy : P.ky, (double)v);
}

// Euler stage, trajectory.py:82-87 (samples level k only), boundaries excluded.
template <typename T, bool STRICT, int ORDER, int LAYOUT, bool R32 = false>
__device__ __forceinline__ void stage_euler(const AdvectParams& P, int k, bool pole, double kx,
                                            double& x, double& y, double& ua, double& va) {
    double s[4];
    if (LAYOUT == kPair4) sample<Pair4Lo<T>, STRICT, ORDER>(P, P.raw_a, P.coef_a, k, pole, x, y, reinterpret_cast<double(&)[2]>(s));
    else if (ORDER >= 2 && pole && P.raw_planar) pole_sample_planar<false, STRICT>(P, k, x, y, s);
    else sample<typename EsPolicy<T, LAYOUT>::type, STRICT, ORDER>(P, P.raw_a, P.coef_a, k, pole, x, y, reinterpret_cast<double(&)[2]>(s));
    if (R32) {
        ua = r32(s[0]); va = r32(s[1]);
        y = __dadd_rn(y, y_increment32(P, false, (float)va));
    } else {
        ua = s[0]; va = s[1];
        y = __dadd_rn(y, __dmul_rn(P.ky, va));
    }
    x = __dadd_rn(x, __dmul_rn(kx, ua));
}

// SETTLS stage, trajectory.py:105-112: pos += 0.5*dt*conv*(va + 2*v_k(pos) - v_{k+1}(pos)).
// PAIR4: the four samples are taken separately and combined in the reference's order.
// ES   : interpolation is linear in the field, so S_k = 2*c_k - c_{k+1} is combined once per grid
//        point at staging time and a single 2-value sample yields 2*v_k(pos) - v_{k+1}(pos):
//        half the gather bytes and half the FMAs of the stage (results differ by rounding only).
// R32  : (ES only) levels k and k+1 are sampled separately from E, because each sample is rounded to f32 before the
//        bracket is formed in f32 -- twice the gathers of the f64 case, the price of the reference's rounding.
template <typename T, bool STRICT, int ORDER, int LAYOUT, bool R32 = false>
__device__ __forceinline__ void stage_settls(const AdvectParams& P, int k, bool pole, double hx,
                                             double ua, double va, double& x, double& y) {
    if (LAYOUT == kPair4) {
        double s[4];
        sample<Pair4<T>, STRICT, ORDER>(P, P.raw_a, P.coef_a, k, pole, x, y, s);
        y = __dadd_rn(y, __dmul_rn(P.hy, __dsub_rn(__dadd_rn(va, __dmul_rn(2.0, s[1])), s[3])));
        x = __dadd_rn(x, __dmul_rn(hx, __dsub_rn(__dadd_rn(ua, __dmul_rn(2.0, s[0])), s[2])));
    } else if (R32) {
        double s[4];
        if (ORDER >= 2 && pole && P.raw_planar) pole_sample_planar<true, STRICT>(P, k, x, y, s);
        else {
            using EP = typename EsPolicy<T, LAYOUT>::type;
            double a[2], b[2];
            sample<EP, STRICT, ORDER>(P, P.raw_a, P.coef_a, k, pole, x, y, a);
            sample<EP, STRICT, ORDER>(P, P.raw_a, P.coef_a, k + 1, pole, x, y, b);
            s[0] = a[0]; s[1] = a[1]; s[2] = b[0]; s[3] = b[1];
        }
        y = __dadd_rn(y, y_increment32(P, true, settls_bracket32(va, s[1], s[3])));
        x = __dadd_rn(x, __dmul_rn(hx, (double)settls_bracket32(ua, s[0], s[2])));
    } else {
        double s[4];
        if (ORDER >= 2 && pole && P.raw_planar) {
            pole_sample_planar<true>(P, k, x, y, s);
            s[0] = 2.0 * s[0] - s[2]; s[1] = 2.0 * s[1] - s[3];       // the reference's own order (trajectory.py:105-112)
        } else sample<typename EsPolicy<T, LAYOUT>::type, STRICT, ORDER>(P, P.raw_b, P.coef_b, k, pole, x, y, reinterpret_cast<double(&)[2]>(s));
        y = __dadd_rn(y, __dmul_rn(P.hy, __dadd_rn(va, s[1])));
        x = __dadd_rn(x, __dmul_rn(hx, __dadd_rn(ua, s[0])));
    }
}

__device__ __forceinline__ void bounds_local(const AdvectParams& P, double& x, double& y) {
    y = clamp_y(y, P.lat_min, P.lat_max);
    if (P.xmode == LCS_X_CYCLIC) x = wrap_x_cyclic(x);
    else x = clamp_x_pointwise(x, P.lon_min, P.lon_max);
}

// ---------------------------------------------------------------------------------------------
#ifndef LCS_FUSED_MINBLOCKS
#define LCS_FUSED_MINBLOCKS 4       // 64 registers, 1024 threads per SM (measured against 3 and 2 blocks: see DESIGN.md)
#endif
template <typename T, bool STRICT, int ORDER, int LAYOUT, bool STRIP = false, bool R32 = false>
__global__ void __launch_bounds__(256, LCS_FUSED_MINBLOCKS)
advect_fused_kernel(const AdvectParams P) {
    const int w = blockIdx.z;
    int row, col;
    if (!particle_rc<STRIP>(P, row, col)) return;
    const int grow = P.row0 + row;
    const bool pole = (grow < ORDER) || (grow >= P.nrow_global - ORDER);   // tools.py:31-33
    const double kx = __ldg(P.kx + row), hx = __ldg(P.hx + row);
    double x = __ldg(P.lon + col), y = __ldg(P.lat + row);                 // trajectory.py:68-70
    const size_t o = (size_t)row * P.ncol + col;
    const size_t wnp = (size_t)w * P.np;
    double* xt = P.x_traj ? P.x_traj + wnp * (P.nsteps + 1) + o : nullptr;
    double* yt = P.y_traj ? P.y_traj + wnp * (P.nsteps + 1) + o : nullptr;
    if (xt) { xt[0] = x; yt[0] = y; }
    const int pair0 = P.level0 + w * P.level_stride;
    for (int t = 0; t < P.nsteps; ++t) {
        double ua, va;
        stage_euler<T, STRICT, ORDER, LAYOUT, R32>(P, pair0 + t, pole, kx, x, y, ua, va);
        bounds_local(P, x, y);
        for (int k = 0; k < P.S; ++k) {
            stage_settls<T, STRICT, ORDER, LAYOUT, R32>(P, pair0 + t, pole, hx, ua, va, x, y);
            bounds_local(P, x, y);
        }
        if (xt) { xt[(size_t)(t + 1) * P.np] = x; yt[(size_t)(t + 1) * P.np] = y; }
    }
    P.x_out[wnp + o] = x;
    P.y_out[wnp + o] = y;
}

// ---------------------------------------------------------------------------------------------
// Outer-product clamp (trajectory.py:96-97 / 122-123).  `px[np.where(px < x_min)] = x_min` on a
// DataArray is an ORTHOGONAL assignment: every (row, col) with row in {rows holding an exit} and
// col in {columns holding an exit} is set, then the same for x_max on the updated array.  This
// couples all particles of a window after every sub-step, so a sub-step q is
//   move(q)  : apply the pending x_min / x_max passes of q-1 (flags are complete by then), run the
//              stage, clamp y, store the state, raise the (row, col) "<x_min" flags of q and append
//              particles with x > x_max to the window's candidate list;
//   gtpass(q): tiny launch over the candidates: x' = x_min if rowflag&colflag, and if still
//              x' > x_max raise the ">x_max" flags of q.
// Flags of sub-step q, window w: flags + ((w*nsub + q)*2 + which)*(nrow+ncol); bytes [0,nrow) rows,
// [nrow,nrow+ncol) columns.  Flags and candidate counters start at zero (cleared by lcs_advect).
__device__ __forceinline__ unsigned char* flag_slot(const AdvectParams& P, int w, int q, int which) {
    return P.flags + ((size_t)((size_t)w * P.nsub + q) * 2 + which) * (size_t)(P.nrow + P.ncol);
}

__device__ __forceinline__ double apply_pending(const AdvectParams& P, int w, int q_prev, int row, int col, double x) {
    const unsigned char* lt = flag_slot(P, w, q_prev, 0);
    const unsigned char* gt = flag_slot(P, w, q_prev, 1);
    if (lt[row] && lt[P.nrow + col]) x = P.lon_min;           // trajectory.py:96
    if (gt[row] && gt[P.nrow + col]) x = P.lon_max;           // trajectory.py:97
    return x;
}

template <typename T, bool STRICT, int ORDER, int LAYOUT, bool STRIP = false, bool R32 = false>
__global__ void __launch_bounds__(256)
advect_phase_move(const AdvectParams P, int q /* global sub-step index */, int t /* interval */, int k /* 0: Euler */) {
    const int w = blockIdx.z;
    int row, col;
    if (!particle_rc<STRIP>(P, row, col)) return;
    const int p = row * P.ncol + col;
    const int grow = P.row0 + row;
    const bool pole = (grow < ORDER) || (grow >= P.nrow_global - ORDER);
    const size_t o = (size_t)w * P.np + p;
    double x, y;
    if (q == 0) {
        x = __ldg(P.lon + col); y = __ldg(P.lat + row);
    } else {
        const d2 s = P.spos[o];
        x = apply_pending(P, w, q - 1, row, col, s.x); y = s.y;
    }
    if (k == 0 && P.x_traj) {                         // level t is final once the pending passes ran
        const size_t to = ((size_t)w * (P.nsteps + 1) + t) * P.np + p;
        P.x_traj[to] = x; P.y_traj[to] = y;
    }
    const int pair = P.level0 + w * P.level_stride + t;
    if (k == 0) {
        double ua, va;
        stage_euler<T, STRICT, ORDER, LAYOUT, R32>(P, pair, pole, __ldg(P.kx + row), x, y, ua, va);
        d2 e; e.x = ua; e.y = va;
        P.swind[o] = e;
    } else {
        const d2 e = P.swind[o];
        stage_settls<T, STRICT, ORDER, LAYOUT, R32>(P, pair, pole, __ldg(P.hx + row), e.x, e.y, x, y);
    }
    y = clamp_y(y, P.lat_min, P.lat_max);
    d2 s; s.x = x; s.y = y;
    P.spos[o] = s;
    if (x < P.lon_min) {
        unsigned char* f = flag_slot(P, w, q, 0);
        f[row] = 1; f[P.nrow + col] = 1;
    } else if (x > P.lon_max) {
        const int slot = atomicAdd(P.cand_count + (size_t)w * P.nsub + q, 1);
        LCS_ASSERT(slot >= 0 && slot < P.np);
        P.cand[(size_t)w * P.np + slot] = p;
    }
}

static __global__ void __launch_bounds__(256)
advect_phase_gtpass(const AdvectParams P, int q) {
    const int w = blockIdx.y;
    const int n = P.cand_count[(size_t)w * P.nsub + q];
    const unsigned char* lt = flag_slot(P, w, q, 0);
    unsigned char* gt = flag_slot(P, w, q, 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = P.cand[(size_t)w * P.np + i];
        const int row = p / P.ncol, col = p - row * P.ncol;
        // x > lon_max here; it survives the x_min pass unless its row and column both hold an exit
        if (!(lt[row] && lt[P.nrow + col])) { gt[row] = 1; gt[P.nrow + col] = 1; }
    }
}

static __global__ void __launch_bounds__(256)
advect_phase_final(const AdvectParams P) {
    const int w = blockIdx.z;
    int row, col;
    if (!particle_rc(P, row, col)) return;
    const int p = row * P.ncol + col;
    double x, y;
    if (P.nsub == 0) { x = __ldg(P.lon + col); y = __ldg(P.lat + row); }
    else {
        const d2 s = P.spos[(size_t)w * P.np + p];
        x = apply_pending(P, w, P.nsub - 1, row, col, s.x); y = s.y;
    }
    const size_t o = (size_t)p;
    P.x_out[(size_t)w * P.np + o] = x; P.y_out[(size_t)w * P.np + o] = y;
    if (P.x_traj) {
        const size_t to = ((size_t)w * (P.nsteps + 1) + P.nsteps) * P.np + o;
        P.x_traj[to] = x; P.y_traj[to] = y;
    }
}

// ---------------------------------------------------------------------------------------------
// Slot enumeration of a window for the persistent kernel: warp-sized tiles of 2 rows x 16 columns
// (6 L1 wavefronts per 16-B gather request instead of 8 for a 4x8 patch, better hit rate than a 1x32
// strip); slot e -> tile e>>5, lane e&31.  Slots past the grid edge are idle.
__device__ __forceinline__ bool slot_rc(const AdvectParams& P, int e, int& row, int& col) {
    const int tile = e >> 5, lane = e & 31;
    const int tr = P.ntc_magic ? (int)(__umulhi((unsigned)tile, P.ntc_magic) >> P.ntc_shift) : (tile >> P.ntc_shift);
    const int tc = tile - tr * P.ntc;
    row = tr * 2 + (lane >> 4);
    col = tc * 16 + (lane & 15);
    return row < P.nrow && col < P.ncol;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {     // p 16-B aligned, bytes a multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Outer-product clamp, group-persistent form (the default).  Round 1's cluster kernel kept one window per CTA
// (or per hardware cluster) resident, so 296 windows x 3 MB of position / Euler-sample state were in flight at once
// and every sub-step streamed it through DRAM (ncu, round 1: 241 GB per 1184 windows, 42 % of the DRAM bandwidth,
// L2 hit rate 37 %).  Here the launch is ONE cooperative grid of `ncta` resident CTAs split into `ngroups` groups; a
// group integrates one window at a time and draws the next one from a global counter, so only `ngroups` windows
// are in flight and their state (ngroups x 3 MB at C2) stays in L2 -- or, STATE = 1, when a group is large enough that
// every thread owns at most ONE slot (few windows on the whole machine), position and Euler sample never leave the
// thread's registers: no state loads, and no stores for the barrier's release fence to drain.  The two
// window-wide dependencies of a sub-step are resolved with a group barrier in global memory (arrive counter +
// spin by one thread per CTA, the cooperative launch guarantees co-residency) and -- when few particles left
// through x_max, the common case -- the second barrier is replaced by every CTA scanning the whole candidate
// list itself.  A group of one CTA degenerates to the per-CTA kernel; a single window is integrated by the
// whole machine (one group), which is the low-latency single-field path.
struct GroupHdr { unsigned next_window; unsigned error; unsigned pad[30]; };           // 128 B
struct GroupCtl {                                     // 256 B per group: the arrival counter and the words the pollers read sit in different L2 lines
    unsigned bar; unsigned pad0[31];
    unsigned released; int window; unsigned pad1[30];
};
struct GroupParams {
    GroupHdr* hdr;
    GroupCtl* ctl;                // [ngroups]
    double2* pos;                 // [ngroups][nslots]  positions (STATE = 0)
    double2* wind;                // [ngroups][nslots]  Euler-stage samples of the current interval
    int* cand;                    // [ngroups][2][nslots]  slots with x > lon_max after sub-step q: buffer q & 1 (a fast CTA appends
                                  //                       those of q+1 while a slow one still scans those of q)
    unsigned char* wbase;         // [ngroups][2 window parities][wstride]: { int count[2][nsub] | u8 flags[nsub][2][nflag_pad] }
    size_t wstride, flags_off;
    int ngroups, ncta, nwindows, nflag_pad, sflag_bytes, redundant_max, prefetch, per_sm;
    // Row-band sharding across GPUs under the outer clamp: the row flags of a band are local (rows are whole), the column
    // flags are the OR over all bands.  xr_world > 1: after the local barrier of a sub-step the group's first CTA posts
    // its band's column flags (and candidate count) into every rank's mailbox over NVLink, waits for everybody's, and
    // writes the OR back into the local flag page (include/lcs_b200.h: lcs_xrank).
    long long* timing;                    // LCS_OUTER_TIMING variant builds: [ncta][3] cycles in phase A / barrier / phase B
    int xr_world, xr_rank;
    unsigned char* const* xr_mail;        // [world] mailbox base of every rank as mapped into this process
    size_t xr_group_stride, xr_msg_stride, xr_hdr_bytes;
};

struct XrMsg { unsigned seq, ncand, pad0, pad1; };         // followed by the column flags (bytes)
constexpr size_t kXrHdrBytes = 1024;                      // unsigned seq[ngroups <= 256]: running exchange number per group


#ifndef LCS_GROUP_THREADS
#define LCS_GROUP_THREADS 512
#endif
#ifndef LCS_GROUP_MINBLOCKS
#define LCS_GROUP_MINBLOCKS 2
#endif
constexpr int kGroupThreads = LCS_GROUP_THREADS;
constexpr unsigned kSpinLimit = 1u << 24;             // ~10 s of polling: a lost arrival flags an error instead of hanging the GPU

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Barrier over the CTAs of one group (all resident: cooperative launch).  `target` = arrivals expected since launch.
// A group of ONE CTA (the batched case: a window per CTA) needs bar.sync only.  Otherwise one thread per CTA:
//   fence.release.gpu  -- MEMBAR.ALL.GPU: this CTA's writes (ordered before it by bar.sync) are at L2 before ...
//   red.add            -- ... its arrival is counted (no return value to wait for),
//   poll the counter   -- ld.relaxed.gpu (L2) until every CTA of the group has arrived.
// No acquire fence follows: on sm_100 `fence.acquire.gpu` IS `CCTL.IVALL` -- it throws away the SM's whole L1, i.e. the
// wind taps the next sub-step is about to gather again -- and nothing here needs it: every datum another CTA wrote is read
// with ld.cg (served by L2, the point of coherence), the wind levels and the coordinate tables are read-only for the
// whole launch.  Round 2a's barrier (fence.acq_rel before AND after, a separate release word bumped by the last arriver
// behind a third fence) cost 7 400 cycles of a 16 400-cycle sub-step with the whole machine on one C2 window (clock64
// instrumentation, -DLCS_OUTER_TIMING): three MEMBAR + ERRBAR drains and two L2 round trips on the critical path; this one
// has one drain and one round trip.  -DLCS_GROUP_BARRIER_V1 restores the old sequence for A/B runs.
#ifndef LCS_GROUP_BARRIER_MODE
#define LCS_GROUP_BARRIER_MODE 0
#endif
#ifndef LCS_GROUP_BARRIER_SLEEP
#define LCS_GROUP_BARRIER_SLEEP 64
#endif
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Measured and dropped (one C2 window, 257-266 us either way): skipping the fence in CTAs that wrote nothing other CTAs read,
// arrivals and polls on separate words, back-off between polls (LCS_GROUP_BARRIER_MODE 2 / 3) -- what is left of the barrier
// is L2 latency on a large die (one-way arrival + 1-2 polling round trips) and the skew between CTAs.
__device__ __forceinline__ void group_barrier(GroupCtl* ctl, unsigned target, unsigned* err, int gsize) {
    __syncthreads();
    if (gsize == 1) return;                       // uniform over the CTA
#ifdef LCS_GROUP_BARRIER_V1
    if (threadIdx.x == 0) {
        fence_acq_rel_gpu();
        if (atomicAdd(&ctl->bar, 1u) + 1u == target) {
            fence_acq_rel_gpu();          // acquire the other CTAs' arrivals before releasing everybody (cumulativity)
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(&ctl->released), "r"(target) : "memory");
        } else {
            unsigned spins = 0;
            while ((int)(ld_volatile_u32(&ctl->released) - target) < 0) {
                if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_volatile_u32(err))) { atomicExch(err, 1u); break; }
            }
        }
        fence_acq_rel_gpu();
    }
#elif LCS_GROUP_BARRIER_MODE == 2
    // arrivals and polls on different words: the last arriver (atom with return) bumps `released`
    if (threadIdx.x == 0) {
        asm volatile("fence.release.gpu;" ::: "memory");
        unsigned old;
        asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(&ctl->bar) : "memory");
        if (old + 1u == target) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(&ctl->released), "r"(target) : "memory");
        } else {
            unsigned spins = 0;
            while ((int)(ld_relaxed_gpu_u32(&ctl->released) - target) < 0) {
                if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_volatile_u32(err))) { atomicExch(err, 1u); break; }
            }
        }
    }
#else
    if (threadIdx.x == 0) {
        asm volatile("fence.release.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(&ctl->bar) : "memory");
        unsigned spins = 0;
        while ((int)(ld_relaxed_gpu_u32(&ctl->bar) - target) < 0) {
#if LCS_GROUP_BARRIER_MODE == 3
            __nanosleep(LCS_GROUP_BARRIER_SLEEP);
#endif
            if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_volatile_u32(err))) { atomicExch(err, 1u); break; }
        }
    }
#endif
    __syncthreads();
}

// thread-private state in global memory.  LCS_GROUP_STATE_POLICY: 0 = ld.cg / st.cg (L2 only); 1 = ld.cs / st.cs
// (evict-first); 2 = plain ld (L1-allocating) / st.cg; 3 = ld with L1 evict-first / st.cg.
// The loads of policy 0 are LDG.STRONG.GPU: they bypass L1, so the `prefetch.global.L1` of the next slot issued one iteration
// earlier never serves them and every iteration exposes an L2 round trip -- ncu's source view (round 2b) put 16 % of the
// kernel's stall samples on the position load and its first use.  A slot's state is written and read by ONE thread, and a
// thread always observes its own stores (the SM's L1 is kept coherent with the SM's own global stores), so an
// L1-allocating load is safe here; the prefetch then lands the line in L1 before it is needed.
#ifndef LCS_GROUP_STATE_POLICY
#define LCS_GROUP_STATE_POLICY 0
#endif
__device__ __forceinline__ double2 gld_state(const double2* p) {
#if LCS_GROUP_STATE_POLICY == 1
    return __ldcs(p);
#elif LCS_GROUP_STATE_POLICY == 2
    return *p;
#elif LCS_GROUP_STATE_POLICY == 3
    double2 v;
    asm volatile("ld.global.L1::evict_first.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
#else
    return __ldcg(p);
#endif
}
__device__ __forceinline__ void gst_state(double2* p, double2 v) {
#if LCS_GROUP_STATE_POLICY == 1
    __stcs(p, v);
#else
    __stcg(p, v);
#endif
}

template <typename T, bool STRICT, int ORDER, int LAYOUT, bool EULER, int STATE, bool R32>
__device__ __forceinline__ void group_phase_a(const AdvectParams& P, const GroupParams& G, const int g, const int w,
                                              const int q, const int t, const int e_begin, const int e_end,
                                              const unsigned char* s_f, double2& rpos, double2& rwind, const int wsel) {
    const unsigned sbase = (unsigned)g * (unsigned)P.nslots;        // host guarantees ngroups*nslots < 2^32
    constexpr int nthr_w = kGroupThreads;                           // a CTA owns a contiguous block of slots (see the kernel)
    int it = 0;
    for (int e = e_begin; e < e_end; e += nthr_w, ++it) {
        double2* const ps = G.pos + (sbase + (unsigned)e);             // STATE = 1: unused (one slot per thread, in registers)
        double2* const pw = G.wind + (sbase + (unsigned)e);
        if (STATE == 1) {
        } else if (G.prefetch == 3) {
            // one bulk L2 prefetch per CTA and iteration (TMA unit: no LSU wavefronts, unlike 2 x 16 prefetch instructions)
            if (threadIdx.x == 0 && e + nthr_w < e_end) {
                const unsigned bytes = (unsigned)min(nthr_w, e_end - (e + nthr_w)) * (unsigned)sizeof(double2);
                if (q != 0) bulk_prefetch_l2(ps + nthr_w, bytes);
                if (!EULER) bulk_prefetch_l2(pw + nthr_w, bytes);
            }
        } else if (G.prefetch && e + nthr_w < e_end) {               // next slot's state: on its way by the time it is needed
            if (G.prefetch == 1) {
                if (q != 0) prefetch_l1(ps + nthr_w);
                if (!EULER) prefetch_l1(pw + nthr_w);
            } else {
                if (q != 0) prefetch_l2(ps + nthr_w);
                if (!EULER) prefetch_l2(pw + nthr_w);
            }
        }
        int row, col;
        LCS_ASSERT(e >= 0 && e < P.nslots && g >= 0 && g < G.ngroups);
        if (!slot_rc(P, e, row, col)) continue;
        const int grow = P.row0 + row;
        const bool pole = (grow < ORDER) || (grow >= P.nrow_global - ORDER);
        double x, y;
        if (q == 0) { x = __ldg(P.lon + col); y = __ldg(P.lat + row); }
        else {
            const double2 s = (STATE == 1) ? rpos : gld_state(ps);
            x = s.x; y = s.y;
            const unsigned m = (unsigned)s_f[row] & (unsigned)s_f[P.nrow + col];   // bit 0: "< x_min" pass, bit 1: "> x_max" pass
            if (m & 1u) x = P.lon_min;                               // trajectory.py:96
            if (m & 2u) x = P.lon_max;                               // trajectory.py:97
        }
        const int pair = P.level0 + w * P.level_stride + t;
        if (EULER) {
            if (P.x_traj) {                                          // level t is final once the pending passes ran
                const size_t to = ((size_t)w * (P.nsteps + 1) + t) * P.np + (size_t)row * P.ncol + col;
                P.x_traj[to] = x; P.y_traj[to] = y;
            }
            double ua, va;
            stage_euler<T, STRICT, ORDER, LAYOUT, R32>(P, pair, pole, __ldg(P.kx + row), x, y, ua, va);
            if (STATE == 1) rwind = make_double2(ua, va);
            else gst_state(pw, make_double2(ua, va));
        } else {
            const double2 wv = (STATE == 1) ? rwind : gld_state(pw);
            stage_settls<T, STRICT, ORDER, LAYOUT, R32>(P, pair, pole, __ldg(P.hx + row), wv.x, wv.y, x, y);
        }
        y = clamp_y(y, P.lat_min, P.lat_max);
        if (STATE == 1) rpos = make_double2(x, y);
        else gst_state(ps, make_double2(x, y));
        if (x < P.lon_min) {
            // rare: the window's flag page and the slot's row / column are re-derived rather than kept live
            unsigned char* g_lt = G.wbase + (size_t)wsel * G.wstride + G.flags_off + (size_t)(2 * q) * G.nflag_pad;
            int r2, c2;
            slot_rc(P, e, r2, c2);
            g_lt[r2] = 1; g_lt[P.nrow + c2] = 1;
        } else if (x > P.lon_max) {
            const int slot = atomicAdd(reinterpret_cast<int*>(G.wbase + (size_t)wsel * G.wstride) + q, 1);
            LCS_ASSERT(slot >= 0 && slot < P.nslots && q >= 0 && q < P.nsub);
            G.cand[((size_t)(2 * g + (q & 1))) * P.nslots + slot] = e;
        }
    }
}

__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Exchange number `seq` of group g, run by every thread of the group's first CTA: OR the `ncol` column-flag bytes at
// `cols` (local global memory, complete: the group barrier ran) over all ranks, in place, and return the sum of
// `my_ncand`.  Messages are double-buffered by seq & 1: a rank posts exchange e+2 only after it finished e+1, which
// every rank posted only after reading all of e.
__device__ __forceinline__ unsigned xr_exchange(const GroupParams& G, int g, unsigned seq, unsigned char* cols, int ncol,
                                                unsigned my_ncand, unsigned* err, unsigned* s_total) {
    const size_t slot = (size_t)g * G.xr_group_stride + (size_t)(seq & 1u) * G.xr_world * G.xr_msg_stride;
    const size_t mine = G.xr_hdr_bytes + slot + (size_t)G.xr_rank * G.xr_msg_stride;
    for (int d = 0; d < G.xr_world; ++d) {                                   // post: plain stores into peer memory
        unsigned char* m = G.xr_mail[d] + mine + sizeof(XrMsg);
        for (int i = threadIdx.x; i < ncol; i += kGroupThreads) m[i] = __ldcg(cols + i);
    }
    __syncthreads();
    if ((int)threadIdx.x < G.xr_world) {
        fence_acq_rel_sys();                                                 // the CTA's stores (ordered by bar.sync) before the flag
        XrMsg* m = reinterpret_cast<XrMsg*>(G.xr_mail[threadIdx.x] + mine);
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(&m->ncand), "r"(my_ncand) : "memory");
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&m->seq), "r"(seq) : "memory");
        // wait for the message of rank threadIdx.x in the local mailbox
        const XrMsg* in = reinterpret_cast<const XrMsg*>(G.xr_mail[G.xr_rank] + G.xr_hdr_bytes + slot + (size_t)threadIdx.x * G.xr_msg_stride);
        unsigned spins = 0, v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&in->seq) : "memory");
            if (v == seq) break;
            if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_volatile_u32(err))) { atomicExch(err, 1u); break; }
        }
        unsigned nc;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(nc) : "l"(&in->ncand) : "memory");
        atomicAdd(s_total, nc);
        fence_acq_rel_sys();
    }
    __syncthreads();
    const unsigned char* base = G.xr_mail[G.xr_rank] + G.xr_hdr_bytes + slot + sizeof(XrMsg);
    for (int i = threadIdx.x; i < ncol; i += kGroupThreads) {
        unsigned char o = 0;
        for (int r = 0; r < G.xr_world; ++r) o |= __ldcg(base + (size_t)r * G.xr_msg_stride + i);
        __stcg(cols + i, o);
    }
    const unsigned total = *s_total;
    __syncthreads();
    if (threadIdx.x == 0) *s_total = 0;
    return total;
}

template <typename T, bool STRICT, int ORDER, int LAYOUT, int STATE, bool R32 = false>
__global__ void __launch_bounds__(kGroupThreads, LCS_GROUP_MINBLOCKS)
advect_outer_group_kernel(const AdvectParams P, const GroupParams G) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned char* const s_f = s_raw;                    // [rows | cols] bit 0 = "< x_min" flag, bit 1 = "> x_max" flag
    double2 rpos = make_double2(0.0, 0.0), rwind = make_double2(0.0, 0.0);          // STATE = 1: this thread's one slot
    // G.per_sm > 1: CTAs b, b + nsm, ... (the ones that share an SM on an idle device) are neighbours in `bid`, so they
    // land in the same group and an SM works on one window at a time; 1: consecutive CTAs, an SM hosts several groups
    const int nsm_ = G.ncta / G.per_sm;
    const int bid = G.per_sm > 1 ? (int)(blockIdx.x % nsm_) * G.per_sm + (int)(blockIdx.x / nsm_) : (int)blockIdx.x;
    const int g = (int)(((long long)(bid + 1) * G.ngroups - 1) / G.ncta);          // CTAs [g*ncta/ngroups, (g+1)*ncta/ngroups)
    const int c0 = (int)((long long)g * G.ncta / G.ngroups);
    const int gsize = (int)((long long)(g + 1) * G.ncta / G.ngroups) - c0;
    const int tid_w = (bid - c0) * kGroupThreads + threadIdx.x;
    const int nthr_w = gsize * kGroupThreads;
    // slots of this CTA: a contiguous block (whole 2x16 tiles), swept in order -- consecutive iterations of a warp then
    // touch neighbouring particle rows, whose 4x4 stencils share three of five coefficient rows in L1 (a group-strided
    // assignment lost that: ncu showed the L1 hit rate falling from 64 % to 59 % between groups of 1 and 8 CTAs)
    const int chunk = ((P.nslots >> 5) + gsize - 1) / gsize * 32;      // whole 2x16 tiles, the same number for every CTA of the group
    const int e_begin = (bid - c0) * chunk + threadIdx.x;
    const int e_end = min((bid - c0 + 1) * chunk, P.nslots);
    GroupCtl* const ctl = G.ctl + g;
    unsigned* const err = &G.hdr->error;
    unsigned arrivals = 0;                                // arrivals of this group expected at its next barrier
    int parity = 0;
    const bool xr = G.xr_world > 1;
    const bool leader = bid == c0;                        // the group's first CTA runs the cross-rank exchanges
    __shared__ unsigned s_xr_total;
    unsigned xr_seq = 0;                                  // running exchange number of this group (persists across calls)
    if (xr && leader) {
        if (threadIdx.x == 0) s_xr_total = 0;
        xr_seq = __ldcg(reinterpret_cast<const unsigned*>(G.xr_mail[G.xr_rank]) + g);
        __syncthreads();
    }
    int wstatic = g;                                      // cross-rank mode: window -> group assignment must agree on all ranks
#ifdef LCS_OUTER_TIMING
    long long tA = 0, tB = 0, tC = 0, t0_ = 0, t1_ = 0;    // variant builds only: where a sub-step's cycles go (scripts/probe_r2b.py timing)
#define LCS_TICK(acc) do { t1_ = clock64(); acc += t1_ - t0_; t0_ = t1_; } while (0)
#else
#define LCS_TICK(acc) do { } while (0)
#endif
    for (;;) {
        unsigned char* const wb = G.wbase + ((size_t)g * 2 + parity) * G.wstride;
        {   // this window's candidate counters and exit flags start cleared (the other parity may still be read)
            uint4* c = reinterpret_cast<uint4*>(wb);
            const int n16 = (int)(G.wstride >> 4);
            for (int i = tid_w; i < n16; i += nthr_w) __stcg(c + i, make_uint4(0u, 0u, 0u, 0u));
        }
        if (tid_w == 0) { ctl->window = xr ? wstatic : (int)atomicAdd(&G.hdr->next_window, 1u); }
        wstatic += G.ngroups;
        group_barrier(ctl, arrivals += gsize, err, gsize);
        const int w = __ldcg(&ctl->window);
        if (w >= G.nwindows) break;
        int q = 0;
#ifdef LCS_OUTER_TIMING
        t0_ = clock64();
#endif
        for (int t = 0; t < P.nsteps; ++t) {
            for (int k = 0; k <= P.S; ++k, ++q) {
                // ---- phase A: pending clamps of q-1 (mirrored in smem), stage, y clamp, raise "< x_min" flags
                if (k == 0) group_phase_a<T, STRICT, ORDER, LAYOUT, true, STATE, R32>(P, G, g, w, q, t, e_begin, e_end, s_f, rpos, rwind, 2 * g + parity);
                else group_phase_a<T, STRICT, ORDER, LAYOUT, false, STATE, R32>(P, G, g, w, q, t, e_begin, e_end, s_f, rpos, rwind, 2 * g + parity);
                LCS_TICK(tA);
                group_barrier(ctl, arrivals += gsize, err, gsize);
                LCS_TICK(tB);
                unsigned char* const lt_page = wb + G.flags_off + (size_t)(2 * q) * G.nflag_pad;
                int* const counts = reinterpret_cast<int*>(wb);
                if (xr) {
                    // ---- cross-rank: OR the "< x_min" column flags of all bands, sum the candidate counts
                    if (leader) {
                        const unsigned total = xr_exchange(G, g, ++xr_seq, lt_page + P.nrow, P.ncol, (unsigned)__ldcg(counts + q), err, &s_xr_total);
                        if (threadIdx.x == 0) __stcg(counts + P.nsub + q, (int)total);
                    }
                    group_barrier(ctl, arrivals += gsize, err, gsize);
                }
                // ---- phase B: mirror the "< x_min" flags; candidates that survive that pass raise the "> x_max" flags
                const unsigned* g_lt = reinterpret_cast<const unsigned*>(lt_page);
                unsigned* const s_f32 = reinterpret_cast<unsigned*>(s_f);
                const int ncand = __ldcg(counts + q);                                                              // in flight with the mirror loads
                const int ncand_all = xr ? __ldcg(counts + P.nsub + q) : ncand;                                    // over all ranks
                for (int i = threadIdx.x; i < (G.nflag_pad >> 2); i += kGroupThreads) s_f32[i] = __ldcg(g_lt + i);   // bytes 0 / 1
                const int* cand = G.cand + (size_t)(2 * g + (q & 1)) * P.nslots;
                // the first candidate of every thread is fetched before its count is known (one L2 round trip less on the
                // sub-step's critical path; entries past the count are stale and ignored)
                const int cand_first = (int)threadIdx.x < P.nslots ? __ldcg(cand + threadIdx.x) : 0;
                __syncthreads();
                if (ncand_all > 0) {                                 // uniform over the group (and over the ranks)
                    if (!xr && ncand <= G.redundant_max) {
                        // few candidates: every CTA scans them all and sets the "> x_max" bits itself -- no second barrier
                        for (int i = threadIdx.x; i < ncand; i += kGroupThreads) {
                            int row, col;
                            const int ce = i == (int)threadIdx.x ? cand_first : __ldcg(cand + i);
                            LCS_ASSERT(ce >= 0 && ce < P.nslots);
                            const bool cin = slot_rc(P, ce, row, col);
                            LCS_ASSERT(cin); (void)cin;
                            // bit 0 is stable in this phase and bit 1 only ever goes 0 -> 1: plain byte read-modify-writes are safe
                            if (!(s_f[row] & s_f[P.nrow + col] & 1)) { s_f[row] |= 2; s_f[P.nrow + col] |= 2; }
                        }
                        __syncthreads();
                    } else {
                        unsigned char* g_gt = lt_page + G.nflag_pad;
                        for (int i = tid_w; i < ncand; i += nthr_w) {
                            int row, col;
                            slot_rc(P, __ldcg(cand + i), row, col);
                            if (!(s_f[row] & s_f[P.nrow + col] & 1)) { g_gt[row] = 1; g_gt[P.nrow + col] = 1; }
                        }
                        group_barrier(ctl, arrivals += gsize, err, gsize);
                        if (xr) {
                            if (leader) xr_exchange(G, g, ++xr_seq, g_gt + P.nrow, P.ncol, 0u, err, &s_xr_total);
                            group_barrier(ctl, arrivals += gsize, err, gsize);
                        }
                        const unsigned* g_gt32 = reinterpret_cast<const unsigned*>(g_gt);
                        for (int i = threadIdx.x; i < (G.nflag_pad >> 2); i += kGroupThreads) s_f32[i] |= __ldcg(g_gt32 + i) << 1;
                        __syncthreads();
                    }
                }
                LCS_TICK(tC);
            }
        }
        // ---- final: pending clamps of the last sub-step, outputs
        const unsigned sbase = (unsigned)g * (unsigned)P.nslots;
        int it = 0;
        for (int e = e_begin; e < e_end; e += kGroupThreads, ++it) {
            int row, col;
            if (!slot_rc(P, e, row, col)) continue;
            const double2 s = (STATE == 1) ? rpos : gld_state(G.pos + (sbase + (unsigned)e));
            double x = s.x;
            const unsigned m = (unsigned)s_f[row] & (unsigned)s_f[P.nrow + col];
            if (m & 1u) x = P.lon_min;
            if (m & 2u) x = P.lon_max;
            const size_t o = (size_t)row * P.ncol + col;
            P.x_out[(size_t)w * P.np + o] = x; P.y_out[(size_t)w * P.np + o] = s.y;
            if (P.x_traj) {
                const size_t to = ((size_t)w * (P.nsteps + 1) + P.nsteps) * P.np + o;
                P.x_traj[to] = x; P.y_traj[to] = s.y;
            }
        }
        parity ^= 1;
    }
    if (xr && leader && threadIdx.x == 0) __stcg(reinterpret_cast<unsigned*>(G.xr_mail[G.xr_rank]) + g, xr_seq);
#ifdef LCS_OUTER_TIMING
    if (threadIdx.x == 0 && G.timing) { G.timing[3 * blockIdx.x] = tA; G.timing[3 * blockIdx.x + 1] = tB; G.timing[3 * blockIdx.x + 2] = tC; }
#endif
}

// Occupancy of a kernel on the current device, cached per (device, kernel, block, shared memory); thread-safe.
static int lcs_blocks_per_sm(const void* func, int threads, size_t smem) {
    struct Key { int dev; const void* f; int threads; size_t smem; int value; };
    static Key cache[64];
    static int ncache = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < ncache; ++i)
        if (cache[i].dev == dev && cache[i].f == func && cache[i].threads == threads && cache[i].smem == smem) return cache[i].value;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, func, threads, smem) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    if (ncache < 64) cache[ncache++] = Key{dev, func, threads, smem, n};
    return n;
}

// 0: group-persistent kernel (default), 1: phased launches
static int lcs_outer_mode() { return lcs_env_int("LCS_OUTER_MODE", 0) == 1 ? 1 : 0; }

// Sizes of the group path that do not depend on the kernel instantiation (lcs_advect_workspace_bytes needs them
// before the winds are known): windows in flight and the workspace layout.
struct GroupLayout {
    int ngroups_max;                       // upper bound of windows in flight (the launch may use fewer: ncta)
    int nflag_pad;
    size_t wstride, flags_off;
    size_t hdr, ctl, pos, wind, cand, wbase, head_bytes, total;
};
static size_t lcs_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t xr_msg_stride(int ncol) { return lcs_align_up(sizeof(XrMsg) + (size_t)ncol, 16); }

static GroupLayout group_layout(int nrow, int ncol, int nslots, int nsub, int nwindows, int xr_groups = 0) {
    GroupLayout L;
    // Windows in flight.  Measured (C2, 1184 windows, B200): one CTA per window is fastest -- 71.8 ms against 74.6 / 79.9 ms
    // with 2 / 8 CTAs per window, although the latter keep the state of the windows in flight inside L2: the kernel is
    // bound by L1 data-pipe wavefronts and issue slots, not by the DRAM round trip of the state (42 % of the DRAM
    // bandwidth, latency hidden by the prefetch).  So: as many groups as windows, capped by the resident CTAs; fewer
    // windows than CTAs are spread over the whole machine.  LCS_OUTER_L2_MB / LCS_OUTER_GROUPS override for experiments.
    long long n = nwindows;
    const int l2mb = lcs_env_int("LCS_OUTER_L2_MB", 0);
    if (l2mb > 0) n = ((long long)l2mb << 20) / ((long long)nslots * 32);
    const int forced = lcs_env_int("LCS_OUTER_GROUPS", 0);
    if (forced > 0) n = forced;
    if (xr_groups > 0) n = xr_groups;                    // cross-rank mode: the caller fixes the windows in flight for all ranks
    if (n < 1) n = 1;
    if (n > nwindows) n = nwindows;
    if (n > 4 * lcs_sm_count()) n = 4 * lcs_sm_count();
    L.ngroups_max = (int)n;
    L.nflag_pad = (int)lcs_align_up((size_t)(nrow + ncol), 16);
    L.flags_off = lcs_align_up((size_t)nsub * 2 * sizeof(int), 16);      // candidate counts: local, and summed over the ranks
    L.wstride = L.flags_off + (size_t)nsub * 2 * L.nflag_pad;
    L.hdr = 0;
    L.ctl = L.hdr + sizeof(GroupHdr);
    L.head_bytes = L.ctl + (size_t)n * sizeof(GroupCtl);
    L.pos = lcs_align_up(L.head_bytes, 256);
    L.wind = L.pos + lcs_align_up((size_t)n * nslots * sizeof(double2), 256);
    L.cand = L.wind + lcs_align_up((size_t)n * nslots * sizeof(double2), 256);
    L.wbase = L.cand + lcs_align_up((size_t)n * 2 * nslots * sizeof(int), 256);
    L.total = L.wbase + (size_t)n * 2 * L.wstride;
    return L;
}

template <typename T, bool STRICT, int ORDER, int LAYOUT, bool R32>
static cudaError_t launch_outer_group(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st, bool* launched) {
    *launched = false;
    const lcs_xrank* xr = (P.xr_host && P.xr_host->world > 1) ? P.xr_host : nullptr;
    const GroupLayout L = group_layout(P.nrow, P.ncol, P.nslots, P.nsub, nwindows, xr ? xr->ngroups : 0);
    // state placement: registers when every thread of the smallest group owns at most one slot (LCS_OUTER_STATE: -1 = that
    // rule, 0 = always global memory, 1 = registers whenever possible -- the same thing, kept for the tests' sake)
    const int state_req = lcs_env_int("LCS_OUTER_STATE", -1);
    const int nsm = lcs_sm_count();
    const int sflag = (int)lcs_align_up((size_t)L.nflag_pad, 16);
    int ncta = 0, ngroups = 0, per_sm_used = 1, state = 0;
    const size_t smem = (size_t)sflag;
    const void* fn = nullptr;
    const int cap = lcs_env_int("LCS_OUTER_CTAS", 0);
    for (int pass = 0; pass < 2; ++pass) {
        fn = state == 1 ? (const void*)advect_outer_group_kernel<T, STRICT, ORDER, LAYOUT, 1, R32>
                        : (const void*)advect_outer_group_kernel<T, STRICT, ORDER, LAYOUT, 0, R32>;
        per_sm_used = lcs_blocks_per_sm(fn, kGroupThreads, smem);
        ncta = per_sm_used * nsm;
        if (cap > 0 && cap < ncta) { ncta = cap; per_sm_used = 1; }
        if (pass == 1 || state_req == 0 || ncta < 1) break;
        const int n = L.ngroups_max < ncta ? L.ngroups_max : ncta;
        const int gmin = ncta / n;                                    // smallest group
        const int chunk = ((P.nslots >> 5) + gmin - 1) / gmin * 32;
        if (chunk > kGroupThreads) break;                             // several slots per thread: state in global memory
        state = 1;                                                    // occupancy of that instantiation: second pass
    }
    if (ncta < 1) return cudaSuccess;                                 // caller reports the failure
    ngroups = L.ngroups_max < ncta ? L.ngroups_max : ncta;
    if (xr && ngroups != xr->ngroups) return cudaErrorInvalidValue;    // every rank must run the same groups
    if ((unsigned long long)ngroups * (unsigned long long)P.nslots >= (1ULL << 32)) return cudaSuccess;
    char* wsb = static_cast<char*>(workspace);
    GroupParams G{};
    G.hdr = reinterpret_cast<GroupHdr*>(wsb + L.hdr);
    G.ctl = reinterpret_cast<GroupCtl*>(wsb + L.ctl);
    G.pos = reinterpret_cast<double2*>(wsb + L.pos);
    G.wind = reinterpret_cast<double2*>(wsb + L.wind);
    G.cand = reinterpret_cast<int*>(wsb + L.cand);
    G.wbase = reinterpret_cast<unsigned char*>(wsb + L.wbase);
    G.wstride = L.wstride; G.flags_off = L.flags_off;
    G.ngroups = ngroups; G.ncta = ncta; G.nwindows = nwindows; G.nflag_pad = L.nflag_pad; G.sflag_bytes = sflag;
    G.redundant_max = lcs_env_int("LCS_OUTER_REDUNDANT", 2048);
    G.prefetch = lcs_env_int("LCS_OUTER_PREFETCH", 1);
    G.per_sm = lcs_env_int("LCS_OUTER_SAMESM", 0) ? per_sm_used : 1;
    if (xr) {
        G.xr_world = xr->world; G.xr_rank = xr->rank;
        G.xr_mail = static_cast<unsigned char* const*>(xr->mailboxes);
        G.xr_msg_stride = xr_msg_stride(P.ncol);
        G.xr_group_stride = 2 * (size_t)xr->world * G.xr_msg_stride;
        G.xr_hdr_bytes = kXrHdrBytes;
    }
    cudaError_t e = cudaMemsetAsync(wsb, 0, L.head_bytes, st);        // window counter, error word, barrier counters
    if (e != cudaSuccess) return e;
    if (lcs_env_int("LCS_DEBUG_CLUSTER", 0))
        fprintf(stderr, "[lcs] outer group kernel: %d windows, %d CTAs in %d groups, state %d, %zu B smem\n",
                nwindows, ncta, ngroups, state, smem);
    AdvectParams Pc = P;
    void* args[2] = {&Pc, &G};
#ifdef LCS_OUTER_TIMING
    static long long* d_timing = nullptr;
    if (!d_timing) cudaMalloc(&d_timing, 4096 * 3 * sizeof(long long));
    G.timing = d_timing;
#endif
    e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)ncta), dim3(kGroupThreads), args, smem, st);
    if (e == cudaSuccess) { *launched = true; lcs_count_launches(1); }
#ifdef LCS_OUTER_TIMING
    if (e == cudaSuccess && lcs_env_int("LCS_OUTER_TIMING_PRINT", 0)) {
        static long long h[4096 * 3];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_timing, (size_t)ncta * 3 * sizeof(long long), cudaMemcpyDeviceToHost);
        double a = 0, b = 0, c = 0, amax = 0, amin = 1e30;
        for (int i = 0; i < ncta; ++i) { a += h[3 * i]; b += h[3 * i + 1]; c += h[3 * i + 2]; if (h[3 * i] > amax) amax = h[3 * i]; if (h[3 * i] < amin) amin = h[3 * i]; }
        const double nq = (double)P.nsub * ((nwindows + ngroups - 1) / ngroups);
        fprintf(stderr, "[lcs timing] %d windows, %d CTAs, %d groups: cycles per sub-step (mean over CTAs): phase A %.0f (min %.0f max %.0f), barrier %.0f, phase B %.0f\n",
                nwindows, ncta, ngroups, a / ncta / nq, amin / nq, amax / nq, b / ncta / nq, c / ncta / nq);
    }
#endif
    return e;
}

// ---------------------------------------------------------------------------------------------
template <typename T, bool STRICT, int ORDER, int LAYOUT, bool R32 = false>
static cudaError_t launch_advect(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st) {
    // Block size: 256 threads, except for launches of fewer than two blocks per SM (the 89 x 180 grid of the example: 63
    // blocks of 256 threads per window), which are cut into blocks of 128 so that every SM gets a share.  Measured (B200,
    // r2b_stage.log): C1 x 4 windows 0.105 -> 0.088 ms; a C2 window (423 blocks) is better left at 256 (0.097 vs 0.102 ms).
    // The warp -> particle patch is the same for every block size (LCS_ADVECT_BLOCK = 64 / 128 / 256 forces one).
    int bt = 256;
    {
        const int forced = lcs_env_int("LCS_ADVECT_BLOCK", 0);
        const long long rows_b = (P.nrow + P.band - 1) / P.band;
        const int tw256 = 256 >> P.band_log2;
        if (forced == 64 || forced == 128 || forced == 256) bt = forced;
        else if ((long long)((P.ncol + tw256 - 1) / tw256) * rows_b * nwindows < 2LL * lcs_sm_count()) bt = 128;
        if ((bt >> P.band_log2) < 1) bt = 256;
    }
    const dim3 block((unsigned)bt);
    const int tw = bt >> P.band_log2;
    const dim3 grid((unsigned)((P.ncol + tw - 1) / tw), (unsigned)((P.nrow + P.band - 1) / P.band), (unsigned)nwindows);
    // wide grids: 8 x 32 blocks of one-row warps (see particle_rc); LCS_ADVECT_STRIP=0/1 forces the choice
    constexpr bool kHasStrip = sizeof(T) == 8 && !STRICT && ORDER == 3 && LAYOUT == kES && !R32;
    const dim3 sgrid((unsigned)((P.ncol + 31) / 32), (unsigned)((P.nrow + 7) / 8), (unsigned)nwindows);
    bool strip = false;
    if (kHasStrip) {
        const int sv = lcs_env_int("LCS_ADVECT_STRIP", -1);
        strip = (sv == 1 || (sv < 0 && P.ncol >= 640)) && sgrid.y <= 65535;
    }
    if (P.xmode != LCS_X_CLAMP_OUTER) {
        if constexpr (kHasStrip) {
            if (strip) {
                advect_fused_kernel<T, STRICT, ORDER, LAYOUT, true, R32><<<sgrid, dim3(256), 0, st>>>(P);
                lcs_count_launches(1);
                return cudaGetLastError();
            }
        }
        advect_fused_kernel<T, STRICT, ORDER, LAYOUT, false, R32><<<grid, block, 0, st>>>(P);
        lcs_count_launches(1);
        return cudaGetLastError();
    }
    // Outer-product clamp.  Default: the group-persistent kernel.  LCS_OUTER_MODE=1: one launch pair per sub-step
    // (kernel boundaries as barriers) -- the independent implementation the tests compare the persistent kernel with.
    if (P.nsub > 0 && lcs_outer_mode() == 0) {
        if (P.nrow + P.ncol > 48 * 1024) return cudaErrorInvalidValue;
        bool launched = false;
        const cudaError_t e = launch_outer_group<T, STRICT, ORDER, LAYOUT, R32>(P, nwindows, workspace, st, &launched);
        return (e == cudaSuccess && !launched) ? cudaErrorLaunchOutOfResources : e;
    }
    const dim3 ggrid(4, (unsigned)nwindows);
    const dim3 gblock(256);
    for (int q = 0; q < P.nsub; ++q) {
        if constexpr (kHasStrip) {
            if (strip) advect_phase_move<T, STRICT, ORDER, LAYOUT, true, R32><<<sgrid, dim3(256), 0, st>>>(P, q, q / (1 + P.S), q % (1 + P.S));
        }
        if (!strip) advect_phase_move<T, STRICT, ORDER, LAYOUT, false, R32><<<grid, block, 0, st>>>(P, q, q / (1 + P.S), q % (1 + P.S));
        advect_phase_gtpass<<<ggrid, gblock, 0, st>>>(P, q);
    }
    advect_phase_final<<<grid, block, 0, st>>>(P);
    lcs_count_launches(2 * P.nsub + 1);
    return cudaGetLastError();
}

}  // namespace lcs
