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
    int np;                       // nrow*ncol (< 2^31)
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
    // phased (outer clamp) state, indexed by the particle's enumeration index p
    d2* spos;                     // [nwindows][np] (x, y)
    d2* swind;                    // [nwindows][np] (ua, va) of the interval's Euler stage
    int* cand;                    // [nwindows][np] particles with x > lon_max after the current sub-step
    int* cand_count;              // [nwindows][nsub]
    unsigned char* flags;         // [nwindows][nsub][2][nrow+ncol]
    int nsub;
};

// Particle enumeration: bands of `band` rows, column-major inside a band, so that a warp covers
// a (band x 32/band) patch -- smaller unique tap footprint than a 1x32 strip, no idle tail lanes.
__device__ __forceinline__ void particle_rc(const AdvectParams& P, int p, int& row, int& col) {
    const int per_band = P.band * P.ncol;
    const int nbands = (P.nrow + P.band - 1) / P.band;
    int b = p / per_band;
    if (b > nbands - 1) b = nbands - 1;
    const int q = p - b * per_band;
    const int h = (b == nbands - 1) ? (P.nrow - b * P.band) : P.band;
    col = q / h;
    row = b * P.band + (q - col * h);
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
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
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

template <typename T, bool STRICT, int ORDER>
__global__ void __launch_bounds__(256)
advect_phase_move(const AdvectParams P, int q /* global sub-step index */) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    int row, col;
    particle_rc(P, p, row, col);
    const int per = 1 + P.S;
    const int t = q / per, k = q - t * per;          // k == 0: Euler stage
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
        const size_t to = ((size_t)w * (P.nsteps + 1) + t) * P.np + (size_t)row * P.ncol + col;
        P.x_traj[to] = x; P.y_traj[to] = y;
    }
    const int pair = P.level0 + w * P.level_stride + t;
    if (k == 0) {
        double ua, va;
        stage_euler<T, STRICT, ORDER>(P, pair, pole, __ldg(P.kx + row), x, y, ua, va);
        d2 e; e.x = ua; e.y = va;
        P.swind[o] = e;
    } else {
        const d2 e = P.swind[o];
        stage_settls<T, STRICT, ORDER>(P, pair, pole, __ldg(P.hx + row), e.x, e.y, x, y);
    }
    y = clamp_y(y, P.lat_min, P.lat_max);
    d2 s; s.x = x; s.y = y;
    P.spos[o] = s;
    if (x < P.lon_min) {
        unsigned char* f = flag_slot(P, w, q, 0);
        f[row] = 1; f[P.nrow + col] = 1;
    } else if (x > P.lon_max) {
        const int slot = atomicAdd(P.cand_count + (size_t)w * P.nsub + q, 1);
        P.cand[(size_t)w * P.np + slot] = p;
    }
}

__global__ void __launch_bounds__(256)
advect_phase_gtpass(const AdvectParams P, int q) {
    const int w = blockIdx.y;
    const int n = P.cand_count[(size_t)w * P.nsub + q];
    const unsigned char* lt = flag_slot(P, w, q, 0);
    unsigned char* gt = flag_slot(P, w, q, 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = P.cand[(size_t)w * P.np + i];
        int row, col;
        particle_rc(P, p, row, col);
        // x > lon_max here; it survives the x_min pass unless its row and column both hold an exit
        if (!(lt[row] && lt[P.nrow + col])) { gt[row] = 1; gt[P.nrow + col] = 1; }
    }
}

__global__ void __launch_bounds__(256)
advect_phase_final(const AdvectParams P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (p >= P.np) return;
    int row, col;
    particle_rc(P, p, row, col);
    double x, y;
    if (P.nsub == 0) { x = __ldg(P.lon + col); y = __ldg(P.lat + row); }
    else {
        const d2 s = P.spos[(size_t)w * P.np + p];
        x = apply_pending(P, w, P.nsub - 1, row, col, s.x); y = s.y;
    }
    const size_t o = (size_t)row * P.ncol + col;
    P.x_out[(size_t)w * P.np + o] = x; P.y_out[(size_t)w * P.np + o] = y;
    if (P.x_traj) {
        const size_t to = ((size_t)w * (P.nsteps + 1) + P.nsteps) * P.np + o;
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
    const dim3 ggrid(4, (unsigned)nwindows);
    for (int q = 0; q < P.nsub; ++q) {
        advect_phase_move<T, STRICT, ORDER><<<grid, block, 0, st>>>(P, q);
        advect_phase_gtpass<<<ggrid, block, 0, st>>>(P, q);
    }
    advect_phase_final<<<grid, block, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace lcs

using namespace lcs;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace of the outer-clamp path: [pos | wind | cand | cleared: cand_count, flags]
struct WsLayout { size_t pos, wind, cand, count, flags, clear_bytes, total; };
static WsLayout ws_layout(const lcs_particles* p, const lcs_advect_opts* o) {
    const size_t np = (size_t)p->nrow * p->ncol;
    const size_t nsub = (size_t)o->nsteps * (1 + o->settls_order);
    WsLayout L;
    L.pos = 0;
    L.wind = L.pos + align_up((size_t)o->nwindows * np * sizeof(d2), 256);
    L.cand = L.wind + align_up((size_t)o->nwindows * np * sizeof(d2), 256);
    L.count = L.cand + align_up((size_t)o->nwindows * np * sizeof(int), 256);
    L.flags = L.count + align_up((size_t)o->nwindows * nsub * sizeof(int), 256);
    L.total = L.flags + align_up((size_t)o->nwindows * nsub * 2 * (size_t)(p->nrow + p->ncol), 256);
    L.clear_bytes = L.total - L.count;
    return L;
}

extern "C" size_t lcs_advect_workspace_bytes(const lcs_particles* p, const lcs_advect_opts* o) {
    if (!p || !o || o->xmode != LCS_X_CLAMP_OUTER) return 0;
    const WsLayout L = ws_layout(p, o);
    return L.total;
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
    if ((long long)p->nrow * p->ncol >= (1LL << 31)) return lcs_fail(LCS_E_INVALID, "lcs_advect: too many particles per window");
    P.np = p->nrow * p->ncol;
    P.lat = p->lat; P.lon = p->lon; P.kx = p->kx; P.hx = p->hx; P.ky = p->ky; P.hy = p->hy;
    P.nsteps = o->nsteps; P.S = o->settls_order; P.xmode = o->xmode;
    P.level0 = o->level0; P.level_stride = o->level_stride;
    P.band = lcs_env_int("LCS_ADVECT_BAND", 4);
    if (P.band < 1) P.band = 1;
    if (P.band > 32) P.band = 32;
    P.x_out = x_out; P.y_out = y_out; P.x_traj = x_traj; P.y_traj = y_traj;
    P.nsub = o->nsteps * (1 + o->settls_order);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (o->xmode == LCS_X_CLAMP_OUTER) {
        const WsLayout L = ws_layout(p, o);
        char* w = static_cast<char*>(workspace);
        P.spos = reinterpret_cast<d2*>(w + L.pos);
        P.swind = reinterpret_cast<d2*>(w + L.wind);
        P.cand = reinterpret_cast<int*>(w + L.cand);
        P.cand_count = reinterpret_cast<int*>(w + L.count);
        P.flags = reinterpret_cast<unsigned char*>(w + L.flags);
        if (P.nsub > 0) {                                     // exit flags and candidate counters start cleared
            e = cudaMemsetAsync(w + L.count, 0, L.clear_bytes, st);
            if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect(memset)");
        }
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
