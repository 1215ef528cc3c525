// Departure-point integrator: parcel_propagation's loop (trajectory.py:80-126) with the
// xr_map_coordinates calls inside it (tools.py:11-41) as CUDA kernels for sm_100a.
//
// Two launch shapes share the same stage functions:
//  * advect_fused_kernel  -- one thread per particle carries it across every wind interval and
//    SETTLS sub-iteration (cyclic / pointwise x-boundary, where particles are independent);
//  * advect_phase_*       -- the as-executed outer-product x-clamp (quirk Q6) couples all
//    particles of a window after every sub-step, so each sub-step is a pair of launches that
//    meet through per-substep row/column exit flags.
//
// Data: packed pairs[k][lat][lon] = (u_k, v_k, u_{k+1}, v_{k+1}); one 32-B (f64) vector load per
// tap feeds the four operands of a SETTLS stage.  Positions stay in registers in the fused
// kernel.  Gathers go through the read-only L1 path; the wind pairs of the two active levels
// are L2-resident (2.9 MB at 281x321, 33 MB at 721x1440).
#include "lcs_internal.h"
#include "lcs_device.cuh"

namespace lcs {

struct AdvectParams {
    const void* raw;
    const void* coef;
    size_t plane;                 // nlat*nlon (elements per packed level)
    int nlat, nlon;
    double nlat_d, nlon_d;
    double lat_min, lat_span, lon_min, lon_span, lat_max, lon_max;
    // particles
    int nrow, ncol, row0, nrow_global;
    long long np;                 // nrow*ncol
    const double* lat;
    const double* lon;
    const double* kx;
    const double* hx;
    double ky, hy;
    int nsteps, S, xmode, level0, level_stride, band;
    double* x_out;
    double* y_out;
    double* x_traj;
    double* y_traj;
    // phased (outer clamp) state
    double* sx; double* sy; double* sua; double* sva;   // [nwindows][np]
    unsigned char* flags;                                // [nwindows][nsub][2][nrow+ncol]
    int nsub;
};

// Particle enumeration: bands of `band` rows, column-major inside a band, so that a warp covers
// a (band x 32/band) patch -- smaller unique tap footprint than a 1x32 strip, no idle tail lanes.
__device__ __forceinline__ void particle_rc(const AdvectParams& P, long long p, int& row, int& col) {
    const long long per_band = (long long)P.band * P.ncol;
    const int nbands = (P.nrow + P.band - 1) / P.band;
    int b = (int)(p / per_band);
    if (b > nbands - 1) b = nbands - 1;
    const long long q = p - (long long)b * per_band;
    const int h = (b == nbands - 1) ? (P.nrow - b * P.band) : P.band;
    col = (int)(q / h);
    row = b * P.band + (int)(q - (long long)col * h);
}

template <typename T, bool STRICT, int ORDER, int NV>
__device__ __forceinline__ void sample(const AdvectParams& P, int pair_idx, bool pole, double x, double y,
                                       double (&out)[NV]) {
    using PT = typename PairOf<T>::type;
    const double iy = index_map(y, P.lat_min, P.lat_span, P.nlat_d);
    const double ix = index_map(x, P.lon_min, P.lon_span, P.nlon_d);
    if (pole) {
        gather_linear_constant<T, STRICT, NV>(reinterpret_cast<const PT*>(P.raw) + (size_t)pair_idx * P.plane,
                                              P.nlat, P.nlon, iy, ix, out);
    } else if (ORDER == 3) {
        gather_cubic_wrap<T, STRICT, NV>(reinterpret_cast<const PT*>(P.coef) + (size_t)pair_idx * P.plane,
                                         P.nlat, P.nlon, iy, ix, out);
    } else {
        gather_linear_wrap<T, STRICT, NV>(reinterpret_cast<const PT*>(P.raw) + (size_t)pair_idx * P.plane,
                                          P.nlat, P.nlon, iy, ix, out);
    }
}

// Euler stage, trajectory.py:82-87 (sample level k only), boundaries excluded.
template <typename T, bool STRICT, int ORDER>
__device__ __forceinline__ void stage_euler(const AdvectParams& P, int pair_idx, bool pole, double kx,
                                            double& x, double& y, double& ua, double& va) {
    double s[2];
    sample<T, STRICT, ORDER, 2>(P, pair_idx, pole, x, y, s);
    ua = s[0]; va = s[1];
    y = __dadd_rn(y, __dmul_rn(P.ky, va));
    x = __dadd_rn(x, __dmul_rn(kx, ua));
}

// SETTLS stage, trajectory.py:105-112: pos += 0.5*dt*conv*(va + 2*v_k(pos) - v_{k+1}(pos))
template <typename T, bool STRICT, int ORDER>
__device__ __forceinline__ void stage_settls(const AdvectParams& P, int pair_idx, bool pole, double hx,
                                             double ua, double va, double& x, double& y) {
    double s[4];
    sample<T, STRICT, ORDER, 4>(P, pair_idx, pole, x, y, s);
    y = __dadd_rn(y, __dmul_rn(P.hy, __dsub_rn(__dadd_rn(va, __dmul_rn(2.0, s[1])), s[3])));
    x = __dadd_rn(x, __dmul_rn(hx, __dsub_rn(__dadd_rn(ua, __dmul_rn(2.0, s[0])), s[2])));
}

__device__ __forceinline__ void bounds_local(const AdvectParams& P, double& x, double& y) {
    y = clamp_y(y, P.lat_min, P.lat_max);
    if (P.xmode == LCS_X_CYCLIC) x = wrap_x_cyclic(x);
    else x = clamp_x_pointwise(x, P.lon_min, P.lon_max);
}

// ---------------------------------------------------------------------------------------------
template <typename T, bool STRICT, int ORDER>
__global__ void __launch_bounds__(256)
advect_fused_kernel(const AdvectParams P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    int row, col;
    particle_rc(P, p, row, col);
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
        stage_euler<T, STRICT, ORDER>(P, pair0 + t, pole, kx, x, y, ua, va);
        bounds_local(P, x, y);
        for (int k = 0; k < P.S; ++k) {
            stage_settls<T, STRICT, ORDER>(P, pair0 + t, pole, hx, ua, va, x, y);
            bounds_local(P, x, y);
        }
        if (xt) { xt[(size_t)(t + 1) * P.np] = x; yt[(size_t)(t + 1) * P.np] = y; }
    }
    P.x_out[wnp + o] = x;
    P.y_out[wnp + o] = y;
}

// ---------------------------------------------------------------------------------------------
// Outer-product clamp (trajectory.py:96-97): after each sub-step
//   A: x_new computed, y clamped, flag rows/cols with x_new < x_min         (advect_phase_move)
//   B: x' = x_min where rowflag&colflag; flag rows/cols with x' > x_max      (advect_phase_lt)
//   next A (or the final launch) first applies x'' = x_max where rowflag&colflag of B.
// Flag slot for sub-step q of window w: flags + ((w*nsub + q)*2 + which) * (nrow+ncol);
// bytes [0,nrow) are row flags, [nrow, nrow+ncol) column flags.  All slots start at zero.
__device__ __forceinline__ unsigned char* flag_slot(const AdvectParams& P, int w, int q, int which) {
    return P.flags + ((size_t)((size_t)w * P.nsub + q) * 2 + which) * (size_t)(P.nrow + P.ncol);
}

template <typename T, bool STRICT, int ORDER>
__global__ void __launch_bounds__(256)
advect_phase_move(const AdvectParams P, int q /* global sub-step index */) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    int row, col;
    particle_rc(P, p, row, col);
    const int per = 1 + P.S;
    const int t = q / per, k = q - t * per;          // k == 0: Euler stage
    const int grow = P.row0 + row;
    const bool pole = (grow < ORDER) || (grow >= P.nrow_global - ORDER);
    const size_t o = (size_t)w * P.np + (size_t)row * P.ncol + col;
    double x, y;
    if (q == 0) {
        x = __ldg(P.lon + col); y = __ldg(P.lat + row);
        if (P.x_traj) {
            const size_t to = (size_t)w * P.np * (P.nsteps + 1) + (size_t)row * P.ncol + col;
            P.x_traj[to] = x; P.y_traj[to] = y;
        }
    } else {
        x = P.sx[o]; y = P.sy[o];
        const unsigned char* g = flag_slot(P, w, q - 1, 1);           // pending x_max pass
        if (g[row] && g[P.nrow + col]) x = P.lon_max;
        if (k == 0 && P.x_traj) {                                      // level t is now final
            const size_t to = ((size_t)w * (P.nsteps + 1) + t) * P.np + (size_t)row * P.ncol + col;
            P.x_traj[to] = x; P.y_traj[to] = y;
        }
    }
    const int pair = P.level0 + w * P.level_stride + t;
    if (k == 0) {
        double ua, va;
        stage_euler<T, STRICT, ORDER>(P, pair, pole, __ldg(P.kx + row), x, y, ua, va);
        P.sua[o] = ua; P.sva[o] = va;
    } else {
        stage_settls<T, STRICT, ORDER>(P, pair, pole, __ldg(P.hx + row), P.sua[o], P.sva[o], x, y);
    }
    y = clamp_y(y, P.lat_min, P.lat_max);
    P.sx[o] = x; P.sy[o] = y;
    if (x < P.lon_min) {
        unsigned char* f = flag_slot(P, w, q, 0);
        f[row] = 1; f[P.nrow + col] = 1;
    }
}

__global__ void __launch_bounds__(256)
advect_phase_lt(const AdvectParams P, int q) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    const int row = (int)(p / P.ncol), col = (int)(p - (long long)row * P.ncol);
    const size_t o = (size_t)w * P.np + (size_t)p;
    const unsigned char* f = flag_slot(P, w, q, 0);
    double x = P.sx[o];
    if (f[row] && f[P.nrow + col]) { x = P.lon_min; P.sx[o] = x; }
    if (x > P.lon_max) {
        unsigned char* g = flag_slot(P, w, q, 1);
        g[row] = 1; g[P.nrow + col] = 1;
    }
}

__global__ void __launch_bounds__(256)
advect_phase_final(const AdvectParams P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    const int row = (int)(p / P.ncol), col = (int)(p - (long long)row * P.ncol);
    const size_t o = (size_t)w * P.np + (size_t)p;
    double x, y;
    if (P.nsub == 0) { x = __ldg(P.lon + col); y = __ldg(P.lat + row); }
    else {
        x = P.sx[o]; y = P.sy[o];
        const unsigned char* g = flag_slot(P, w, P.nsub - 1, 1);
        if (g[row] && g[P.nrow + col]) x = P.lon_max;
    }
    P.x_out[o] = x; P.y_out[o] = y;
    if (P.x_traj) {
        const size_t to = ((size_t)w * (P.nsteps + 1) + P.nsteps) * P.np + (size_t)p;
        P.x_traj[to] = x; P.y_traj[to] = y;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T, bool STRICT, int ORDER>
static cudaError_t launch_advect(const AdvectParams& P, int nwindows, cudaStream_t st) {
    const dim3 block(256);
    const dim3 grid((unsigned)((P.np + 255) / 256), (unsigned)nwindows);
    if (P.xmode != LCS_X_CLAMP_OUTER) {
        advect_fused_kernel<T, STRICT, ORDER><<<grid, block, 0, st>>>(P);
        return cudaGetLastError();
    }
    for (int q = 0; q < P.nsub; ++q) {
        advect_phase_move<T, STRICT, ORDER><<<grid, block, 0, st>>>(P, q);
        advect_phase_lt<<<grid, block, 0, st>>>(P, q);
    }
    advect_phase_final<<<grid, block, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace lcs

using namespace lcs;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t lcs_advect_workspace_bytes(const lcs_particles* p, const lcs_advect_opts* o) {
    if (!p || !o || o->xmode != LCS_X_CLAMP_OUTER) return 0;
    const size_t np = (size_t)p->nrow * p->ncol;
    const size_t nsub = (size_t)o->nsteps * (1 + o->settls_order);
    const size_t state = align_up((size_t)o->nwindows * np * sizeof(double), 256);
    const size_t flags = align_up((size_t)o->nwindows * nsub * 2 * (size_t)(p->nrow + p->ncol), 256);
    return 4 * state + flags;
}

extern "C" int lcs_advect(const lcs_grid* g, const lcs_particles* p, const lcs_advect_opts* o,
                          const void* raw_pairs, const void* coef_pairs,
                          double* x_out, double* y_out, double* x_traj, double* y_traj,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !p || !o || !raw_pairs || !x_out || !y_out) return lcs_fail(LCS_E_INVALID, "lcs_advect: null argument");
    if (o->interp_order != 1 && o->interp_order != 3)
        return lcs_fail(LCS_E_UNSUPPORTED, "lcs_advect: interp_order must be 1 or 3");
    if (o->interp_order == 3 && !coef_pairs) return lcs_fail(LCS_E_INVALID, "lcs_advect: coef_pairs required for order 3");
    if (g->nlat < 4 || g->nlon < 4) return lcs_fail(LCS_E_INVALID, "lcs_advect: grid must be at least 4x4");
    if (p->nrow < 1 || p->ncol < 1 || o->nwindows < 1 || o->nsteps < 0 || o->settls_order < 0)
        return lcs_fail(LCS_E_INVALID, "lcs_advect: bad sizes");
    if (o->xmode < LCS_X_CYCLIC || o->xmode > LCS_X_CLAMP_OUTER) return lcs_fail(LCS_E_INVALID, "lcs_advect: bad xmode");
    if ((x_traj == nullptr) != (y_traj == nullptr)) return lcs_fail(LCS_E_INVALID, "lcs_advect: x_traj/y_traj must both be set");
    const size_t need = lcs_advect_workspace_bytes(p, o);
    if (need > workspace_bytes || (need && !workspace)) return lcs_fail(LCS_E_WORKSPACE, "lcs_advect: workspace too small");

    AdvectParams P{};
    P.raw = raw_pairs; P.coef = coef_pairs;
    P.plane = (size_t)g->nlat * g->nlon;
    P.nlat = g->nlat; P.nlon = g->nlon;
    P.nlat_d = (double)g->nlat; P.nlon_d = (double)g->nlon;
    P.lat_min = g->lat_min; P.lat_max = g->lat_max; P.lat_span = g->lat_max - g->lat_min;
    P.lon_min = g->lon_min; P.lon_max = g->lon_max; P.lon_span = g->lon_max - g->lon_min;
    P.nrow = p->nrow; P.ncol = p->ncol; P.row0 = p->row0; P.nrow_global = p->nrow_global;
    P.np = (long long)p->nrow * p->ncol;
    P.lat = p->lat; P.lon = p->lon; P.kx = p->kx; P.hx = p->hx; P.ky = p->ky; P.hy = p->hy;
    P.nsteps = o->nsteps; P.S = o->settls_order; P.xmode = o->xmode;
    P.level0 = o->level0; P.level_stride = o->level_stride;
    P.band = lcs_env_int("LCS_ADVECT_BAND", 4);
    if (P.band < 1) P.band = 1;
    if (P.band > 32) P.band = 32;
    P.x_out = x_out; P.y_out = y_out; P.x_traj = x_traj; P.y_traj = y_traj;
    P.nsub = o->nsteps * (1 + o->settls_order);
    if (o->xmode == LCS_X_CLAMP_OUTER) {
        const size_t state = align_up((size_t)o->nwindows * (size_t)P.np * sizeof(double), 256);
        char* w = static_cast<char*>(workspace);
        P.sx = reinterpret_cast<double*>(w);
        P.sy = reinterpret_cast<double*>(w + state);
        P.sua = reinterpret_cast<double*>(w + 2 * state);
        P.sva = reinterpret_cast<double*>(w + 3 * state);
        P.flags = reinterpret_cast<unsigned char*>(w + 4 * state);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (o->xmode == LCS_X_CLAMP_OUTER && P.nsub > 0) {       // exit flags start cleared
        e = cudaMemsetAsync(P.flags, 0, (size_t)o->nwindows * P.nsub * 2 * (size_t)(P.nrow + P.ncol), st);
        if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect(memset)");
    }
    const bool strict = o->strict != 0;
    const int ord = o->interp_order;
#define LCS_DISPATCH(TT)                                                                      \
    do {                                                                                      \
        if (ord == 3) e = strict ? launch_advect<TT, true, 3>(P, o->nwindows, st)             \
                                 : launch_advect<TT, false, 3>(P, o->nwindows, st);           \
        else e = strict ? launch_advect<TT, true, 1>(P, o->nwindows, st)                      \
                        : launch_advect<TT, false, 1>(P, o->nwindows, st);                    \
    } while (0)
    if (o->pair_dtype == LCS_F64) LCS_DISPATCH(double);
    else if (o->pair_dtype == LCS_F32) LCS_DISPATCH(float);
    else return lcs_fail(LCS_E_INVALID, "lcs_advect: bad pair_dtype");
#undef LCS_DISPATCH
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect");
    return LCS_OK;
}
