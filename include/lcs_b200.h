/*
 * lcs_b200.h -- C ABI of the B200-native FTLE engine (liblcs_b200.so, sm_100a only).
 *
 * The reference (gabrielmpp/LagrangianCoherence) has no FFI layer: its hot path is three
 * Python callables whose arithmetic lives in scipy.ndimage.map_coordinates, a numba stencil
 * and scipy.linalg.norm.  Each entry point below names the reference code it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller; the library
 *     never allocates device memory (workspaces are sized by lcs_*_workspace_bytes);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and are
 *     thread-safe for distinct streams;
 *   - return value: 0 on success, a negative LCS_E_* code otherwise; lcs_last_error() returns
 *     a thread-local message;
 *   - grids are ascending and row-major [lat][lon]; wind series are [level][lat][lon].
 */
#ifndef LCS_B200_H_
#define LCS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCS_ABI_VERSION 3

enum { LCS_OK = 0, LCS_E_INVALID = -1, LCS_E_CUDA = -2, LCS_E_WORKSPACE = -3, LCS_E_UNSUPPORTED = -4,
       LCS_E_TIMEOUT = -5 };
enum { LCS_F64 = 0, LCS_F32 = 1 };
/* device layouts of a staged wind series, see lcs_pack_pairs / lcs_pack_es */
enum { LCS_LAYOUT_PAIR4 = 0, LCS_LAYOUT_ES = 1 };
/* arithmetic of the cubic taps in the integrator */
enum { LCS_ARITH_F64 = 0,         /* weights, products and sums in f64 whatever the storage type (parity path)      */
       LCS_ARITH_F32 = 1 };       /* fast path: f32 weights and packed f32 FMAs on f32 winds; index map, positions
                                     and the SETTLS update stay f64.  FTLE agrees with the f64 path to ~1e-5 relative
                                     away from ridge-singular points (tolerance-tested, not bit parity)               */
/* x-boundary of the integrator: trajectory.py:92-97 / 118-123 */
enum { LCS_X_CYCLIC = 0,          /* cyclic_xboundary=True: the two `where` with Python-sign % 180 */
       LCS_X_CLAMP_POINTWISE = 1, /* per-particle clamp to [lon_min, lon_max] */
       LCS_X_CLAMP_OUTER = 2 };   /* as executed: xarray orthogonal assignment (rows x cols having any exit) */

/* Halo of the LCS_LAYOUT_ES arrays: every level is stored as (nlat + LO + HI) x (nlon + LO + HI) elements, grid point
 * (i, j) at row i + LO, column j + LO, and the halo cells hold the mirror-reflected values (scipy's spline boundary
 * d c b | a b c d | c b a), so a gather never reflects a tap index.  2 + 3 covers the tap range [-2, n+2] of every
 * spline order up to 5 after the coordinate fold. */
#define LCS_HALO_LO 2
#define LCS_HALO_HI 3
#define LCS_ES_LEVEL_ELEMS(nlat, nlon) ((size_t)((nlat) + LCS_HALO_LO + LCS_HALO_HI) * (size_t)((nlon) + LCS_HALO_LO + LCS_HALO_HI))

/* Wind grid; min/max are coord.min()/coord.max() used by the index map of tools.py:19-22. */
typedef struct lcs_grid {
    int32_t nlat, nlon;
    double lat_min, lat_max, lon_min, lon_max;
} lcs_grid;

/* Particle set = a band of rows of the arrival grid (the reference seeds one particle per
 * wind grid point, trajectory.py:68-70; a finer particle grid is an extension).
 * kx/hx carry timestep*conversion_x(row) and 0.5*timestep*conversion_x(row) evaluated by the
 * host exactly as trajectory.py:56,87,112 does, so the device reproduces numpy's rounding. */
typedef struct lcs_particles {
    int32_t nrow, ncol;          /* rows held by this call (band incl. halo), columns          */
    int32_t row0, nrow_global;   /* global index of the first row, global row count: the
                                    order-1 "pole row" rule of tools.py:31-39 uses global rows */
    const double* lat;           /* device [nrow]  start latitude of each row                   */
    const double* lon;           /* device [ncol]  start longitude of each column               */
    const double* kx;            /* device [nrow]  timestep * conversion_x(row)                 */
    const double* hx;            /* device [nrow]  0.5 * timestep * conversion_x(row)           */
    double ky, hy;               /* timestep * conversion_y ; 0.5 * timestep * conversion_y     */
} lcs_particles;

/* Row-band sharding across GPUs under LCS_X_CLAMP_OUTER: the as-executed clamp sets every (row, column)
 * with the row AND the column holding an exit (trajectory.py:96-97).  Rows are whole inside a band, so the row flags are
 * local; the column flags are the OR over all bands.  With `xrank` set, the integrator of every rank posts its band's
 * column flags after each sub-step into every rank's mailbox over NVLink (plain stores into peer-mapped memory, then a
 * release flag), waits for the others inside the persistent kernel, and continues with the OR -- twice per sub-step
 * when particles also left through x_max.  All ranks must call lcs_advect with the same windows, sub-steps and ngroups.
 *   mailboxes : DEVICE pointer to an array[world] of device pointers, entry d = rank d's mailbox buffer as mapped into
 *               this process (e.g. torch symmetric memory: buffer_ptrs_dev).  The buffers must be zero before the first
 *               call and are reused by later calls (they carry a running exchange number); mailbox_bytes >=
 *               lcs_xrank_mailbox_bytes(world, ngroups, ncol) each.
 *   ngroups   : windows in flight (1 = the whole GPU on one window at a time), identical on every rank. */
typedef struct lcs_xrank {
    int32_t world, rank;
    int32_t ngroups, reserved;
    const void* mailboxes;
    size_t mailbox_bytes;
} lcs_xrank;
size_t lcs_xrank_mailbox_bytes(int world, int ngroups, int ncol);

typedef struct lcs_advect_opts {
    int32_t nsteps;         /* wind intervals to cross = nt-1 (trajectory.py:80)                 */
    int32_t settls_order;   /* accumulating sub-iterations per interval (trajectory.py:100)      */
    int32_t interp_order;   /* 1..5 (tools.py:11 `order`); 2, 4, 5: f64 ES layout only            */
    int32_t xmode;          /* LCS_X_*                                                            */
    int32_t strict;         /* 1: accumulate taps as scipy does ((c*wy)*wx, no fused multiply-add,
                               true divisions); needs LCS_LAYOUT_PAIR4                            */
    int32_t nwindows;       /* independent start times integrated by this call (rolling series)  */
    int32_t level0, level_stride; /* window b starts at packed pair index level0 + b*level_stride  */
    int32_t arith;          /* LCS_ARITH_*; LCS_ARITH_F32 needs dtype LCS_F32, LCS_LAYOUT_ES, order 3   */
    int32_t round32;        /* the reference's dtype propagation for f32 WINDS on f64 coordinates: scipy returns
                               map_coordinates samples in the input dtype (tools.py:26-30) and numpy's promotion
                               rules (NEP 50) then fix the precision of every increment (trajectory.py:86-87,110-112).
                               0: off (f64 winds); 1: every sample rounded to f32, the SETTLS bracket and the y
                               increments in f32 (timestep is a Python scalar); 2: as 1 but y increments in f64
                               (timestep is a numpy f64 scalar, e.g. after resample=).  Needs LCS_LAYOUT_ES of f64
                               coefficients computed from the f32 winds; levels k and k+1 are then sampled
                               separately, so a SETTLS stage costs two gathers                                  */
    const lcs_xrank* xrank; /* NULL, or the cross-rank exchange of the outer clamp for row-band sharding (above)      */
} lcs_advect_opts;

/* Staged winds handed to the integrator.
 *   LCS_LAYOUT_PAIR4: raw_a / coef_a = pairs[k][lat][lon] = (u_k, v_k, u_{k+1}, v_{k+1}), k < nlev-1;
 *                     raw_b / coef_b unused.
 *   LCS_LAYOUT_ES   : raw_a / coef_a = E[k] = (u_k, v_k), k < nlev;
 *                     raw_b / coef_b = S[k] = (2u_k - u_{k+1}, 2v_k - v_{k+1}), k < nlev-1;
 *                     each level in the halo layout above (LCS_ES_LEVEL_ELEMS elements per level).
 * raw_* hold the winds themselves (order-1 pole rows, or interp_order == 1); coef_* their cubic
 * B-spline coefficients (interp_order == 3, else NULL).  dtype is the element storage type. */
typedef struct lcs_winds {
    int32_t layout, dtype;
    const void* raw_a;
    const void* raw_b;
    const void* coef_a;
    const void* coef_b;
    int32_t raw_planar;     /* 1 (LCS_LAYOUT_ES, interp_order >= 2 only): raw_a / raw_b are the planar series u / v
                               [nlev][nlat][nlon] of raw_dtype instead of packed E / S.  Only the 2*order pole rows
                               sample the raw winds at those orders, so no second packed copy is staged for them */
    int32_t raw_dtype;      /* LCS_F64 / LCS_F32 storage of the planar raw series                                  */
} lcs_winds;

/* ---------------------------------------------------------------- housekeeping */
int lcs_abi_version(void);
const char* lcs_last_error(void);
/* kernels launched by this library since it was loaded (all threads); evidence for bench.py's gpu_launches */
unsigned long long lcs_kernel_launches(void);

/* ---------------------------------------------------------------- wind staging
 * lcs_prefilter: B-spline coefficients (order 2..5; 3 = the reference default) of every level, mirror
 * boundary, latitude axis then longitude axis, f64 -- what scipy.ndimage.map_coordinates(order, mode='wrap')
 * recomputes inside every call at tools.py:26-30; here it runs once per level.  u, v: device [nlev][nlat][nlon] of
 * `in_dtype`; coef_u, coef_v: device f64 planes of the same shape (may not alias the inputs);
 * scratch: device buffer of lcs_prefilter_scratch_bytes() bytes. */
size_t lcs_prefilter_scratch_bytes(int nlev, int nlat, int nlon);
int lcs_prefilter(const void* u, const void* v, int in_dtype, double* coef_u, double* coef_v,
                  void* scratch, size_t scratch_bytes, int nlev, int nlat, int nlon, int order, void* stream);

/* lcs_pack_pairs: interleave two planar series into the gather layout
 * pairs[k][lat][lon] = (u_k, v_k, u_{k+1}, v_{k+1}), k = 0..nlev-2, one 32-byte (f64) or 16-byte
 * (f32) element per grid point, so that one vector load feeds all four operands of a SETTLS
 * sub-iteration (trajectory.py:105-108).  Pure data movement, no reference counterpart. */
int lcs_pack_pairs(const void* u, const void* v, int in_dtype, void* pairs, int pair_dtype,
                   int nlev, int nlat, int nlon, void* stream);

/* lcs_pack_es: the fast gather layout.  The SETTLS increment needs 2*f_k(pos) - f_{k+1}(pos)
 * (trajectory.py:110-112); interpolation is linear in the field, so the combination is formed once
 * per grid point: E[k] = (u_k, v_k) for k < nlev and S[k] = (2u_k - u_{k+1}, 2v_k - v_{k+1}) for
 * k < nlev-1 (evaluated in f64, then stored as es_dtype).  A SETTLS stage then gathers 16 B (f64)
 * per tap instead of 32 B.  Levels are written in the halo layout: e_out holds nlev and s_out nlev-1 levels of
 * LCS_ES_LEVEL_ELEMS(nlat, nlon) two-value elements. */
int lcs_pack_es(const void* u, const void* v, int in_dtype, void* e_out, void* s_out, int es_dtype,
                int nlev, int nlat, int nlon, void* stream);

/* lcs_time_lerp: `u.resample({timedim: freq}).interpolate('linear')` of LCS.py:88-91 on the device.  New level k
 * = w_hi[k] * in[lo[k]+1] + w_lo[k] * in[lo[k]], the form scipy 1.18.1's interp1d(kind='linear') evaluates
 * (w_hi = (x-x_lo)/(x_hi-x_lo), w_lo = (x_hi-x)/(x_hi-x_lo), computed by the host).  Needs at least two levels.
 * in: device [nlev][plane] of in_dtype; lo/w_hi/w_lo: device [nnew]; out: device f64 [nnew][plane]. */
int lcs_time_lerp(const void* in, int in_dtype, const int32_t* lo, const double* w_hi, const double* w_lo,
                  int nnew, int64_t plane, double* out, void* stream);

/* lcs_regrid_linear_nearest: the regrid of the global path, LCS.py:105-114: `u.interp(latitude=, longitude=,
 * method='linear')` with its NaNs (targets outside the source coordinates) filled from `u.reindex(..., method='nearest')`.
 * Per axis the host supplies, for every target coordinate, the bracket `lo`, the weights of scipy's interp1d
 * (new = w_hi*y[lo+1] + w_lo*y[lo]), `valid` (inside the source range) and the `nearest` source index
 * (lagrangiancoherence_b200/regrid.py).  in: device [nlev][nlat_src][nlon_src] of in_dtype; out: device f64
 * [nlev][nlat_dst][nlon_dst]; plan arrays: device, length nlat_dst / nlon_dst. */
int lcs_regrid_linear_nearest(const void* in, int in_dtype, int nlev, int nlat_src, int nlon_src,
                              const int32_t* lat_lo, const double* lat_w_hi, const double* lat_w_lo,
                              const uint8_t* lat_valid, const int32_t* lat_nearest,
                              const int32_t* lon_lo, const double* lon_w_hi, const double* lon_w_lo,
                              const uint8_t* lon_valid, const int32_t* lon_nearest,
                              int nlat_dst, int nlon_dst, double* out, void* stream);

/* lcs_spectral_truncate: the triangular spectral truncation of the global path, LCS.py:115-118:
 * `VectorWind(u, v).truncate(field, truncation=T)` = windspharm -> pyspharm grdtospec / spectogrd -> SPHEREPACK shaes /
 * shses on the "regular" grid (rows at colatitudes i*pi/(nlat-1), columns at 2*pi*j/nlon).  The operation is linear
 * and separable; the host builds its tables (lagrangiancoherence_b200/spectral.py): At [T+1][nlat][nlat] = the
 * colatitude operator of every zonal wavenumber, TRANSPOSED (At[m][k][i] = A_m[i][k]); Fc [nlon][2T+1] and
 * Fi [2T+1][nlon] = the forward / inverse longitude transform (columns: mean, cos 1, sin 1, ..., cos T, sin T).
 * in: device [nfields][nlat][nlon] of in_dtype; out: device f64, same shape, not aliasing in; scratch: device,
 * lcs_spectral_truncate_scratch_bytes().  Parity against the reference is UNPINNED for this step (SPHEREPACK is not
 * available in this image); the kernel is tested against the CPU restatement oracle/spectral_oracle.py. */
size_t lcs_spectral_truncate_scratch_bytes(int nfields, int nlat, int ntrunc);
int lcs_spectral_truncate(const void* in, int in_dtype, int nfields, int nlat, int nlon, int ntrunc,
                          const double* At, const double* Fc, const double* Fi,
                          void* scratch, size_t scratch_bytes, double* out, void* stream);

/* ---------------------------------------------------------------- integrator
 * lcs_advect replaces parcel_propagation's loop, trajectory.py:80-126, together with the
 * xr_map_coordinates calls inside it (tools.py:11-41).
 *   w         : staged winds (lcs_pack_pairs or lcs_pack_es, raw and -- for order 3 -- coefficients)
 *   x_out,y_out: device f64 [nwindows][nrow][ncol] final positions
 *   x_traj,y_traj: NULL, or device f64 [nwindows][nsteps+1][nrow][ncol] (level 0 = start grid,
 *                  trajectory.py:76-77,125-126)
 *   workspace : device scratch of lcs_advect_workspace_bytes() bytes (only LCS_X_CLAMP_OUTER
 *               needs any: positions, Euler samples and per-sub-step row/column exit flags of the
 *               windows in flight -- a few dozen at most, not of every window of the call) */
size_t lcs_advect_workspace_bytes(const lcs_particles* p, const lcs_advect_opts* o);
int lcs_advect(const lcs_grid* g, const lcs_particles* p, const lcs_advect_opts* o,
               const lcs_winds* w,
               double* x_out, double* y_out, double* x_traj, double* y_traj,
               void* workspace, size_t workspace_bytes, void* stream);

/* The outer-clamp kernel synchronises the CTAs that share a window through barriers in global memory; a barrier
 * that is not completed within ~10 s gives up instead of hanging the device and records that in the workspace.
 * lcs_advect_check synchronises `stream` and returns LCS_E_TIMEOUT if that happened in the last lcs_advect call
 * that used `workspace` (LCS_OK otherwise, and for the other x-boundary modes). */
int lcs_advect_check(const void* workspace, void* stream);

/* ---------------------------------------------------------------- fused epilogue
 * lcs_ftle_epilogue replaces flowmap_gradient (LCS.py:171-225), the six
 * derivative_spherical_coords/fourth_order_derivative calls (tools.py:190-267) and the batched
 * spectral norm of LCS.py:145-157 with one kernel.
 *   x_dep,y_dep: device f64 [nfields][nrow_in][nlon]; rows cover global rows
 *                [in_row0, in_row0+nrow_in) and must include the +-2 halo of the output band
 *   out_row0,nrow_out: global rows written;  sigma: device f64 [nfields][nrow_out][nlon]
 *   jac: NULL or device f64 [nfields][6][nrow_out][nlon] = dXdx,dXdy,dYdx,dYdy,dZdx,dZdy
 *   dx: device [nlat_global] metric spacing per global row, dy scalar (tools.py:255-256)
 *   mask: NULL or device u8 [nrow_out][nlon] (subdomain crop, LCS.py:143-144): 0 => NaN
 *   log_scale: 0 => sigma_max (what LCS.__call__ returns); 1 => 0.5*log(sigma_max), the scaling
 *              the reference's callers apply (examples/ideal_vortex.py:282)
 *   status: NULL or device int32[1], OR-ed with 1 if any derivative is +-inf (the reference's
 *           scipy.linalg.norm raises ValueError there, LCS.py:154) */
int lcs_ftle_epilogue(const double* x_dep, const double* y_dep, int nfields,
                      int nlat_global, int nlon, int in_row0, int nrow_in,
                      int out_row0, int nrow_out, const double* dx, double dy,
                      const uint8_t* mask, int log_scale,
                      double* sigma, double* jac, int32_t* status, void* stream);

/* lcs_gaussian_filter2d: scipy.ndimage.gaussian_filter(x_departure, sigma) of LCS.py:187-190 (mode='reflect',
 * truncate=4, axis 0 then axis 1, scipy's accumulation order: bit-identical results).  in/out/scratch: device f64
 * [nfields][n0][n1], distinct; weights: device [2*radius+1] normalised kernel, radius = int(4*sigma + 0.5). */
int lcs_gaussian_filter2d(const double* in, double* out, double* scratch, int nfields, int n0, int n1,
                          const double* weights, int radius, void* stream);

/* ---------------------------------------------------------------- array-level seams
 * The three third-party kernels the reference calls, as stand-alone device operations. */

/* xr_map_coordinates body (tools.py:19-41) for one field: positions in degrees,
 * [nrow][ncol]; rows < order or >= nrow_global-order (global row index) take the
 * order-1/'constant' branch.  field: device f64 [nlat][nlon] raw values; coef: its B-spline
 * coefficients of that order (order >= 2) or NULL; out: device f64 [nrow][ncol]. */
int lcs_map_coordinates(const lcs_grid* g, const double* field, const double* coef, int order,
                        const double* pos_x, const double* pos_y, int nrow, int ncol,
                        int row0, int nrow_global, double* out, void* stream);

/* fourth_order_derivative (tools.py:190-245) on an f32 array [n0][n1], dim 0 or 1. */
int lcs_fourth_order_derivative(const float* arr, int n0, int n1, int dim, int isglobal,
                                float* out, void* stream);

/* scipy.linalg.norm(vals[3,3,N], axis=(0,1), ord=2) of LCS.py:154 for the layout the reference
 * feeds it: vals = nine stacked planes of N points (plane k = entry k of the row-major 3x3). */
int lcs_spectral_norm_3x3(const double* vals, int64_t n, double* out, void* stream);

/* lcs_ridge_classify: the per-point loop of find_ridges_spherical_hessian (tools.py:93-136): clean the Hessian
 * (inf, NaN -> 0), eigen-decompose [[hxx, hxy], [hxy, hyy]] with LAPACK's dgeev/dlanv2 conventions (order of the
 * eigenvalues, signs of the eigenvectors), dt = dot(ROW argmin(eigvals) of the eigenvector matrix, gradient) as
 * executed upstream, eigmin = eigenvalue of largest magnitude, dt_prod = 1 where not(|dt| > tolerance) and
 * eigmin < 0 else 0.  Optional (NULL to skip) diagnostics of return_eigvectors=True (tools.py:148-152): dt_raw = the
 * unthresholded dt, evec0/evec1 = the two entries of that eigenvector row.  All arrays: device f64 [n]. */
int lcs_ridge_classify(const double* hxx, const double* hxy, const double* hyy, const double* gx, const double* gy,
                       int64_t n, double tolerance, double* dt_prod, double* eigmin,
                       double* dt_raw, double* evec0, double* evec1, void* stream);

/* ---------------------------------------------------------------- roofline microbenchmark
 * Same taps x taps vector-gather pattern and thread tiling as the integrator with nothing else in the
 * loop (integer positions, one add per loaded value, all rounds independent): the measured upper bound
 * for the L1/L2 gather roofline.  vec_width = values per tap (4: PAIR4 elements, 32 B in f64 -- the
 * reference formulation's bytes; 2: ES elements).  jitter: per-particle pseudo-random displacement of
 * up to +-jitter cells on top of a coherent per-round shift (0 = neighbours stay neighbours). */
int lcs_gather_peak(const void* pairs, int pair_dtype, int vec_width, int nlat, int nlon, int nrow, int ncol,
                    int nwindows, int taps, double jitter, int iters, double* sink, void* stream);

/* The same pattern with the block's tap bounding box staged through shared memory every round (f64 2-value elements,
 * 4x4 taps, zero jitter): the measured ceiling of a shared-memory-tile design in its best case, reported beside the
 * direct-gather ceiling in DESIGN.md.  Bytes are counted as for lcs_gather_peak (taps only). */
int lcs_gather_peak_smem(const void* pairs, int nlat, int nlon, int nrow, int ncol, int nwindows, int iters,
                         double* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCS_B200_H_ */
