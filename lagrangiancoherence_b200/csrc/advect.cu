// lcs_advect: argument validation, workspace layout and dispatch to the kernel instantiations
// (advect_kernels.cuh; compiled in advect_inst_*.cu so that the translation units build in parallel).
#include "advect_kernels.cuh"

namespace lcs {
#define LCS_DECL(name) cudaError_t name(const AdvectParams& P, int nwindows, void* workspace, cudaStream_t st)
LCS_DECL(lcs_launch_f64_es3); LCS_DECL(lcs_launch_f64_es1); LCS_DECL(lcs_launch_f64_es2); LCS_DECL(lcs_launch_f64_es4);
LCS_DECL(lcs_launch_f64_es5); LCS_DECL(lcs_launch_f64_p3); LCS_DECL(lcs_launch_f64_p3s); LCS_DECL(lcs_launch_f64_p1);
LCS_DECL(lcs_launch_f64_p1s); LCS_DECL(lcs_launch_f32_es3); LCS_DECL(lcs_launch_f32_es1); LCS_DECL(lcs_launch_f32_p3);
LCS_DECL(lcs_launch_f32_p3s); LCS_DECL(lcs_launch_f32_p1); LCS_DECL(lcs_launch_f32_p1s); LCS_DECL(lcs_launch_f32_fast3);
LCS_DECL(lcs_launch_r32_es1); LCS_DECL(lcs_launch_r32_es2); LCS_DECL(lcs_launch_r32_es3); LCS_DECL(lcs_launch_r32_es4); LCS_DECL(lcs_launch_r32_es5);
#undef LCS_DECL
}  // namespace lcs

using namespace lcs;

static size_t align_up(size_t v, size_t a) { return lcs_align_up(v, a); }

// workspace of the outer-clamp path: [pos | wind | cand | cleared: cand_count, flags]
struct WsLayout { size_t pos, wind, cand, count, flags, clear_bytes, total; };
static size_t lcs_slots_per_window(int nrow, int ncol) {
    return (size_t)((nrow + 1) / 2) * (size_t)((ncol + 15) / 16) * 32;      // >= nrow*ncol
}
static WsLayout ws_layout(const lcs_particles* p, const lcs_advect_opts* o) {
    const size_t np = lcs_slots_per_window(p->nrow, p->ncol);                // both outer paths index state by slot or p < slots
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
    const size_t nsub = (size_t)o->nsteps * (1 + o->settls_order);
    if (nsub == 0) return 0;
    if (lcs_outer_mode() != 0) return ws_layout(p, o).total;       // phased launches / hardware clusters: state per window
    const int xg = (o->xrank && o->xrank->world > 1) ? o->xrank->ngroups : 0;
    const GroupLayout L = group_layout(p->nrow, p->ncol, (int)lcs_slots_per_window(p->nrow, p->ncol), (int)nsub, o->nwindows, xg);
    return L.total;
}

extern "C" size_t lcs_xrank_mailbox_bytes(int world, int ngroups, int ncol) {
    if (world < 1 || ngroups < 1 || ncol < 1) return 0;
    return kXrHdrBytes + (size_t)ngroups * 2 * (size_t)world * xr_msg_stride(ncol);
}

extern "C" int lcs_advect_check(const void* workspace, void* stream) {
    if (!workspace || lcs_outer_mode() != 0) return LCS_OK;          // only the group kernel keeps a status word
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GroupHdr h{};
    cudaError_t e = cudaMemcpyAsync(&h, workspace, sizeof(unsigned) * 2, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect_check");
    if (h.error) return lcs_fail(LCS_E_TIMEOUT, "lcs_advect: a group barrier of the outer-clamp kernel timed out (results invalid)");
    return LCS_OK;
}

extern "C" int lcs_advect(const lcs_grid* g, const lcs_particles* p, const lcs_advect_opts* o,
                          const lcs_winds* w,
                          double* x_out, double* y_out, double* x_traj, double* y_traj,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !p || !o || !w || !x_out || !y_out) return lcs_fail(LCS_E_INVALID, "lcs_advect: null argument");
    if (o->interp_order < 1 || o->interp_order > 5)
        return lcs_fail(LCS_E_UNSUPPORTED, "lcs_advect: interp_order must be 1..5 (0 is broken upstream: empty row slices)");
    const bool other_order = o->interp_order == 2 || o->interp_order == 4 || o->interp_order == 5;
    if (other_order && (w->dtype != LCS_F64 || w->layout != LCS_LAYOUT_ES || o->strict))
        return lcs_fail(LCS_E_UNSUPPORTED, "lcs_advect: interp_order 2, 4, 5 need f64 winds in the ES layout, strict 0");
    if (w->layout != LCS_LAYOUT_PAIR4 && w->layout != LCS_LAYOUT_ES) return lcs_fail(LCS_E_INVALID, "lcs_advect: bad wind layout");
    if (o->strict && w->layout != LCS_LAYOUT_PAIR4)
        return lcs_fail(LCS_E_INVALID, "lcs_advect: strict evaluation needs the PAIR4 layout");
    if (o->nsteps > 0) {
        if (!w->raw_a || (w->layout == LCS_LAYOUT_ES && !w->raw_b)) return lcs_fail(LCS_E_INVALID, "lcs_advect: raw winds missing");
        if (o->interp_order >= 2 && (!w->coef_a || (w->layout == LCS_LAYOUT_ES && !w->coef_b)))
            return lcs_fail(LCS_E_INVALID, "lcs_advect: spline coefficients required for orders >= 2");
    }
    if (o->arith != LCS_ARITH_F64 && o->arith != LCS_ARITH_F32) return lcs_fail(LCS_E_INVALID, "lcs_advect: bad arith");
    if (o->arith == LCS_ARITH_F32 && (w->dtype != LCS_F32 || w->layout != LCS_LAYOUT_ES || o->interp_order != 3 || o->strict))
        return lcs_fail(LCS_E_INVALID, "lcs_advect: f32 arithmetic needs f32 winds in the ES layout, interp_order 3, strict 0");
    if (w->raw_planar && (w->layout != LCS_LAYOUT_ES || o->interp_order < 2 || (w->raw_dtype != LCS_F64 && w->raw_dtype != LCS_F32)))
        return lcs_fail(LCS_E_INVALID, "lcs_advect: planar raw winds are for the ES layout with interp_order >= 2 (raw_dtype f64/f32)");
    if (o->round32 < 0 || o->round32 > 2) return lcs_fail(LCS_E_INVALID, "lcs_advect: round32 must be 0, 1 or 2");
    if (o->round32 && (w->layout != LCS_LAYOUT_ES || w->dtype != LCS_F64 || o->strict || o->arith != LCS_ARITH_F64))
        return lcs_fail(LCS_E_UNSUPPORTED, "lcs_advect: round32 (f32 dtype propagation) needs f64 ES levels, strict 0, f64 arithmetic");
    if (o->xrank && o->xrank->world > 1) {
        const lcs_xrank* xr = o->xrank;
        if (o->xmode != LCS_X_CLAMP_OUTER) return lcs_fail(LCS_E_INVALID, "lcs_advect: xrank is for LCS_X_CLAMP_OUTER (the other x-boundaries need no exchange)");
        if (lcs_outer_mode() != 0) return lcs_fail(LCS_E_UNSUPPORTED, "lcs_advect: xrank needs the group-persistent kernel (LCS_OUTER_MODE=0)");
        if (xr->rank < 0 || xr->rank >= xr->world || xr->world > kGroupThreads || xr->ngroups < 1 || xr->ngroups > 256 || xr->ngroups > o->nwindows || !xr->mailboxes)
            return lcs_fail(LCS_E_INVALID, "lcs_advect: bad xrank (need 0 <= rank < world, 1 <= ngroups <= min(nwindows, 256), mailboxes)");
        if (xr->mailbox_bytes < lcs_xrank_mailbox_bytes(xr->world, xr->ngroups, p->ncol))
            return lcs_fail(LCS_E_WORKSPACE, "lcs_advect: xrank mailboxes too small (lcs_xrank_mailbox_bytes)");
    }
    if (g->nlat < 4 || g->nlon < 4) return lcs_fail(LCS_E_INVALID, "lcs_advect: grid must be at least 4x4");
    if (p->nrow < 1 || p->ncol < 1 || o->nwindows < 1 || o->nsteps < 0 || o->settls_order < 0)
        return lcs_fail(LCS_E_INVALID, "lcs_advect: bad sizes");
    if (o->nwindows > 65535) return lcs_fail(LCS_E_INVALID, "lcs_advect: at most 65535 windows per call");
    if (o->xmode < LCS_X_CYCLIC || o->xmode > LCS_X_CLAMP_OUTER) return lcs_fail(LCS_E_INVALID, "lcs_advect: bad xmode");
    if ((x_traj == nullptr) != (y_traj == nullptr)) return lcs_fail(LCS_E_INVALID, "lcs_advect: x_traj/y_traj must both be set");
    const size_t need = lcs_advect_workspace_bytes(p, o);
    if (need > workspace_bytes || (need && !workspace)) return lcs_fail(LCS_E_WORKSPACE, "lcs_advect: workspace too small");

    AdvectParams P{};
    P.raw_a = w->raw_a; P.raw_b = w->raw_b; P.coef_a = w->coef_a; P.coef_b = w->coef_b;
    P.raw_planar = w->raw_planar != 0; P.raw_f32 = w->raw_dtype == LCS_F32;
    P.plane = (size_t)g->nlat * g->nlon;
    P.plane_es = LCS_ES_LEVEL_ELEMS(g->nlat, g->nlon);
    P.halo_off = LCS_HALO_LO * (g->nlon + LCS_HALO_LO + LCS_HALO_HI) + LCS_HALO_LO;
    if (P.plane_es >= (1ULL << 31)) return lcs_fail(LCS_E_INVALID, "lcs_advect: grid too large (a level must stay below 2^31 elements)");
    P.nlat = g->nlat; P.nlon = g->nlon;
    P.nlat_d = (double)g->nlat; P.nlon_d = (double)g->nlon;
    P.lat_min = g->lat_min; P.lat_max = g->lat_max; P.lat_span = g->lat_max - g->lat_min;
    P.lon_min = g->lon_min; P.lon_max = g->lon_max; P.lon_span = g->lon_max - g->lon_min;
    P.nlat_over_span = P.nlat_d / P.lat_span; P.nlon_over_span = P.nlon_d / P.lon_span;
    P.nrow = p->nrow; P.ncol = p->ncol; P.row0 = p->row0; P.nrow_global = p->nrow_global;
    if ((long long)p->nrow * p->ncol >= (1LL << 31)) return lcs_fail(LCS_E_INVALID, "lcs_advect: too many particles per window");
    P.np = p->nrow * p->ncol;
    P.lat = p->lat; P.lon = p->lon; P.kx = p->kx; P.hx = p->hx; P.ky = p->ky; P.hy = p->hy;
    P.ky32 = (float)p->ky; P.hy32 = (float)p->hy; P.y_weak = o->round32 == 1;       // NEP 50: a weak Python scalar becomes f32
    P.nsteps = o->nsteps; P.S = o->settls_order; P.xmode = o->xmode;
    P.level0 = o->level0; P.level_stride = o->level_stride;
    P.band_log2 = lcs_env_int("LCS_ADVECT_BAND_LOG2", 1);      // tile = 2 rows x 128 columns, warp = 2 x 16
    if (P.band_log2 < 0) P.band_log2 = 0;
    if (P.band_log2 > 5) P.band_log2 = 5;
    P.band = 1 << P.band_log2;
    if ((P.nrow + P.band - 1) / P.band > 65535) return lcs_fail(LCS_E_INVALID, "lcs_advect: too many row bands");
    P.x_out = x_out; P.y_out = y_out; P.x_traj = x_traj; P.y_traj = y_traj;
    P.nsub = o->nsteps * (1 + o->settls_order);
    P.xr_host = (o->xrank && o->xrank->world > 1) ? o->xrank : nullptr;
    {   // persistent-kernel slot enumeration; tile / ntc for 0 <= tile < 2^31 as a multiply-high:
        // k = floor(log2 ntc), magic = ceil(2^(32+k) / ntc) < 2^32; the rounding error stays below 1/ntc
        P.ntc = (P.ncol + 15) / 16;
        P.nslots = (int)lcs_slots_per_window(P.nrow, P.ncol);
        int k = 0;
        while ((2LL << k) <= P.ntc) ++k;
        P.ntc_shift = k;
        P.ntc_magic = ((1LL << k) == P.ntc) ? 0u
            : (unsigned)(((1ULL << (32 + k)) + (unsigned long long)P.ntc - 1) / (unsigned long long)P.ntc);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (o->xmode == LCS_X_CLAMP_OUTER && P.nsub > 0 && lcs_outer_mode() != 0) {
        if ((unsigned long long)o->nwindows * (unsigned long long)P.nslots >= (1ULL << 32))
            return lcs_fail(LCS_E_INVALID, "lcs_advect: outer clamp: nwindows * slots per window must stay below 2^32 (split the call)");
        const WsLayout L = ws_layout(p, o);
        char* wsb = static_cast<char*>(workspace);
        P.spos = reinterpret_cast<d2*>(wsb + L.pos);
        P.swind = reinterpret_cast<d2*>(wsb + L.wind);
        P.cand = reinterpret_cast<int*>(wsb + L.cand);
        P.cand_count = reinterpret_cast<int*>(wsb + L.count);
        P.flags = reinterpret_cast<unsigned char*>(wsb + L.flags);
        e = cudaMemsetAsync(wsb + L.count, 0, L.clear_bytes, st);   // exit flags and candidate counters start cleared
        if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect(memset)");
    }
    const bool strict = o->strict != 0;
    const int ord = o->interp_order;
    const int nw = o->nwindows;
    const bool es = w->layout == LCS_LAYOUT_ES;
    if (o->round32) {
        e = ord == 1 ? lcs_launch_r32_es1(P, nw, workspace, st) : ord == 2 ? lcs_launch_r32_es2(P, nw, workspace, st)
          : ord == 3 ? lcs_launch_r32_es3(P, nw, workspace, st) : ord == 4 ? lcs_launch_r32_es4(P, nw, workspace, st)
          : lcs_launch_r32_es5(P, nw, workspace, st);
    } else if (o->arith == LCS_ARITH_F32) e = lcs_launch_f32_fast3(P, nw, workspace, st);          // validated above
    else if (ord == 2) e = lcs_launch_f64_es2(P, nw, workspace, st);
    else if (ord == 4) e = lcs_launch_f64_es4(P, nw, workspace, st);
    else if (ord == 5) e = lcs_launch_f64_es5(P, nw, workspace, st);
    else if (w->dtype == LCS_F64) {
        if (es) e = ord == 3 ? lcs_launch_f64_es3(P, nw, workspace, st) : lcs_launch_f64_es1(P, nw, workspace, st);
        else if (ord == 3) e = strict ? lcs_launch_f64_p3s(P, nw, workspace, st) : lcs_launch_f64_p3(P, nw, workspace, st);
        else e = strict ? lcs_launch_f64_p1s(P, nw, workspace, st) : lcs_launch_f64_p1(P, nw, workspace, st);
    } else if (w->dtype == LCS_F32) {
        if (es) e = ord == 3 ? lcs_launch_f32_es3(P, nw, workspace, st) : lcs_launch_f32_es1(P, nw, workspace, st);
        else if (ord == 3) e = strict ? lcs_launch_f32_p3s(P, nw, workspace, st) : lcs_launch_f32_p3(P, nw, workspace, st);
        else e = strict ? lcs_launch_f32_p1s(P, nw, workspace, st) : lcs_launch_f32_p1(P, nw, workspace, st);
    } else return lcs_fail(LCS_E_INVALID, "lcs_advect: bad wind dtype");
    if (e != cudaSuccess) return lcs_fail_cuda(e, "lcs_advect");
    return LCS_OK;
}
