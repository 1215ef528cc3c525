"""Device-level FTLE engine: torch tensors for memory/streams, liblcs_b200.so for every kernel.

This is the layer between the reference-shaped Python API (``LCS``, ``parcel_propagation``,
``flowmap_gradient``) and the C ABI of ``include/lcs_b200.h``.  Host-side scalars and per-row
factors are evaluated with numpy in the reference's order of operations
(trajectory.py:54-57,86-87,110-112; tools.py:253-256) so that the device reproduces numpy's
rounding; everything per-particle or per-grid-point runs in the CUDA kernels.

No CPU fallback: a missing shared object or a non-CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

EARTH_R = 6371000          # trajectory.py:54
XMODES = {'cyclic': _lib.LCS_X_CYCLIC, 'pointwise': _lib.LCS_X_CLAMP_POINTWISE, 'outer': _lib.LCS_X_CLAMP_OUTER}
DTYPES = {'f64': _lib.LCS_F64, 'f32': _lib.LCS_F32}


def precision_args(precision):
    """``precision=`` of the public API -> FtleEngine keywords: 'f64' (parity path), 'f32' (f32 packed winds, f64
    tap arithmetic), 'f32fast' (f32 winds and f32 tap arithmetic, cubic interpolation only)."""
    try:
        return {'f64': dict(pair_dtype='f64', arith='f64'), 'f32': dict(pair_dtype='f32', arith='f64'),
                'f32fast': dict(pair_dtype='f32', arith='f32')}[precision]
    except KeyError:
        raise ValueError("precision must be 'f64', 'f32' or 'f32fast'") from None


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dtype_code(t):
    if t.dtype == torch.float64:
        return _lib.LCS_F64
    if t.dtype == torch.float32:
        return _lib.LCS_F32
    raise TypeError(f'winds must be float32 or float64, got {t.dtype}')


@dataclass
class StagedWinds:
    """Gather layouts of a wind series on the device (lcs_pack_pairs / lcs_pack_es).

    PAIR4: ``raw_a``/``coef_a`` = ``[nlev-1, nlat, nlon, 4]`` pairs.
    ES   : ``raw_a``/``coef_a`` = E ``[nlev, nlat+5, nlon+5, 2]``, ``raw_b``/``coef_b`` = S ``[nlev-1, nlat+5, nlon+5, 2]``
           (halo layout: 2 mirror-filled cells before and 3 after each axis).
    """
    layout: int
    dtype: int
    nlev: int
    raw_a: torch.Tensor | None = None
    raw_b: torch.Tensor | None = None
    coef_a: torch.Tensor | None = None
    coef_b: torch.Tensor | None = None
    raw_planar: bool = False      # ES, orders >= 2: raw_a / raw_b are the planar u / v series (kept alive here)
    round32: int = 0              # lcs_advect_opts.round32: the reference's dtype propagation for f32 winds

    def struct(self):
        p = lambda t: t.data_ptr() if t is not None else None
        raw_dtype = _dtype_code(self.raw_a) if self.raw_planar else _lib.LCS_F64
        return _lib.Winds(self.layout, self.dtype, p(self.raw_a), p(self.raw_b), p(self.coef_a), p(self.coef_b),
                          int(self.raw_planar), raw_dtype)


class FtleEngine:
    """One wind grid + one integration recipe; reusable across calls (buffers are cached).

    Parameters mirror the reference: ``timestep`` [s, sign = direction], ``SETTLS_order``,
    ``interp_order`` (= ``traj_interp_order``), and the x-boundary: ``'cyclic'``
    (cyclic_xboundary=True), ``'outer'`` (as-executed non-cyclic clamp, quirk Q6) or
    ``'pointwise'``.  ``pair_dtype`` selects the storage of the packed winds: ``'f64'`` (parity
    path) or ``'f32'`` (fast path); positions and the epilogue stay f64 either way.
    ``part_lat``/``part_lon`` seed a particle grid different from the wind grid (extension, C5).
    ``f32_propagation``: f32 WINDS (on f64 coordinates) are integrated the way the reference's dtypes propagate --
    scipy returns samples in the input dtype, numpy (NEP 50) forms the y increments in f32 when ``timestep`` is a Python
    scalar (tools.py:26-30, trajectory.py:86-87,110-112) -- which costs two gathers per SETTLS stage; ``False`` promotes
    f32 winds to f64 instead (faster, but then only f64 inputs match the reference to 1e-10).
    """

    def __init__(self, lat, lon, timestep, SETTLS_order=0, interp_order=3, xmode='outer',
                 pair_dtype='f64', strict=False, device='cuda:0', part_lat=None, part_lon=None, layout='es',
                 arith='f64', f32_propagation=True):
        if not torch.cuda.is_available():
            raise _lib.LcsError('lagrangiancoherence_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.device = torch.device(device)
        # f32 coordinates make the reference integrate entirely in f32 (np.meshgrid of f32 coordinates, trajectory.py:68-70);
        # that case is promoted to f64 here and agrees to the reference's own f32 noise only (tests/test_gpu_edges.py)
        self.coords_f64 = np.asarray(lat).dtype == np.float64 and np.asarray(lon).dtype == np.float64
        self.lat = np.ascontiguousarray(lat, dtype=np.float64)
        self.lon = np.ascontiguousarray(lon, dtype=np.float64)
        if np.any(np.diff(self.lat) <= 0) or np.any(np.diff(self.lon) <= 0):
            raise ValueError('latitude/longitude must be ascending (the reference sorts first, LCS.py:101-104)')
        self.nlat, self.nlon = self.lat.size, self.lon.size
        self.timestep = timestep
        self.timestep_weak = not isinstance(timestep, np.generic)      # Python scalar: weak under NEP 50
        self.f32_propagation = bool(f32_propagation)
        self.S = int(SETTLS_order)
        self.order = int(interp_order)
        if self.order not in (1, 2, 3, 4, 5):
            # scipy raises RuntimeError('spline order not supported') outside 0..5; order 0 yields empty row slices
            # upstream (tools.py:25,31-33) and is refused here
            raise RuntimeError('spline order not supported' if self.order != 0 else
                               'interp_order=0 is broken upstream (empty slices in xr_map_coordinates); use 1..5')
        self.xmode = XMODES[xmode]
        self.pair_dtype = DTYPES[pair_dtype]
        self.strict = int(bool(strict))
        self.layout = _lib.LCS_LAYOUT_PAIR4 if self.strict or layout == 'pair4' else _lib.LCS_LAYOUT_ES
        if self.order in (2, 4, 5) and (self.pair_dtype != _lib.LCS_F64 or self.layout != _lib.LCS_LAYOUT_ES):
            raise ValueError('interp_order 2, 4 and 5 run on the f64 ES layout only (no strict / pair4 / f32 variants)')
        if arith not in ('f64', 'f32'):
            raise ValueError("arith must be 'f64' or 'f32'")
        if arith == 'f32' and (self.pair_dtype != _lib.LCS_F32 or self.layout != _lib.LCS_LAYOUT_ES or self.order != 3):
            raise ValueError("arith='f32' needs pair_dtype='f32', the ES layout and interp_order=3")
        self.arith = _lib.LCS_ARITH_F32 if arith == 'f32' else _lib.LCS_ARITH_F64
        self.grid = _lib.Grid(self.nlat, self.nlon, self.lat.min(), self.lat.max(), self.lon.min(), self.lon.max())
        self.part_lat = self.lat if part_lat is None else np.ascontiguousarray(part_lat, dtype=np.float64)
        self.part_lon = self.lon if part_lon is None else np.ascontiguousarray(part_lon, dtype=np.float64)
        # trajectory.py:54-57 (arrival-grid latitude, quirk Q5) and the products of :86-87,:110-112
        conversion_y = 180 / (EARTH_R * np.pi)
        conversion_x = 180 / (np.pi * EARTH_R * np.abs(np.cos(self.part_lat * np.pi / 180)))
        self.ky = float(timestep * conversion_y)
        self.hy = float(0.5 * timestep * conversion_y)
        kx = timestep * conversion_x
        hx = 0.5 * timestep * conversion_x
        # tools.py:253-256 (uniform spacing taken from the first two coordinates)
        y = self.lat * np.pi / 180
        dx = (np.pi / 180) * (self.lon[1] - self.lon[0]) * EARTH_R * np.cos(y)
        self.dy = float((np.pi / 180) * (self.lat[1] - self.lat[0]) * EARTH_R)
        with torch.cuda.device(self.device):
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
            self.d_plat, self.d_plon = dev(self.part_lat), dev(self.part_lon)
            self.d_kx, self.d_hx, self.d_dx = dev(kx), dev(hx), dev(dx)
            self.d_status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ws = None
        self._tmp = {}

    def _buf(self, name, shape, dtype):
        """A temporary of the staging path that never leaves the engine (coefficient planes, filter scratch), or -- with
        ``stage(reuse=True)`` -- the E / S levels: one cached allocation per name, grown when a larger one is asked for.
        All engine work is stream-ordered on the caller's current stream, so a later call may overwrite it."""
        n = 1
        for d in shape:
            n *= int(d)
        t = self._tmp.get(name)
        if t is None or t.dtype != dtype or t.numel() < n:
            t = None
            self._tmp.pop(name, None)
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._tmp[name] = t
        return t[:n].view(tuple(shape))

    # ------------------------------------------------------------------ staging
    def stage(self, u, v, raw='planar', resample=None, reuse=False):
        """Upload (if needed), prefilter (orders >= 2) and pack a wind series ``[nlev, nlat, nlon]``.

        ``resample=(lo, w_hi, w_lo)`` (timeaxis.resample_plan): the series is refined linearly in time on the way
        (``resample=`` of LCS.__call__, LCS.py:88-91).  The spline prefilter is linear, so the COARSE levels are
        prefiltered and the refinement is applied to winds and coefficients alike -- the refined series is never
        prefiltered itself (half of the filter passes for a 2x refinement; differs from filtering the refined winds,
        what upstream does inside every map_coordinates call, by rounding only).

        ``raw``: how the 2*order pole rows of the spline orders get the raw winds they sample -- ``'planar'`` (default)
        lets them read ``u, v`` themselves, ``'packed'`` stages a second E/S copy of the series for them (a third more
        staging traffic and memory; same integrator time: measured 14.26 vs 14.29 ms per 296 C2 windows).  With
        ``'planar'`` the returned object keeps ``u, v`` (their device copies) alive and the integrator reads them: do
        not overwrite them in place between ``stage`` and ``advect``.

        ``reuse=True``: the E / S levels live in buffers owned by the engine and are overwritten by the next
        ``stage(reuse=True)`` call (stream-ordered).  For loops that stage, integrate, stage again -- the rolling series,
        the bench -- this removes ~11 GB of allocator traffic per 1192-level step; callers that keep two staged series
        alive at once leave it off."""
        u = self._to_device(u, 0)
        v = self._to_device(v, 1)
        if u.shape != v.shape or u.dim() != 3 or tuple(u.shape[1:]) != (self.nlat, self.nlon):
            raise ValueError(f'winds must be [nlev, {self.nlat}, {self.nlon}], got {tuple(u.shape)} / {tuple(v.shape)}')
        if resample is not None and u.shape[0] >= 2:
            if self.order >= 2:
                with torch.cuda.device(self.device):
                    cu0 = torch.empty(tuple(u.shape), dtype=torch.float64, device=self.device)
                    cv0 = torch.empty_like(cu0)
                    scratch = torch.empty((2,) + tuple(cu0.shape), dtype=torch.float64, device=self.device)
                    _lib.check(self.lib.lcs_prefilter(_ptr(u), _ptr(v), _dtype_code(u), _ptr(cu0), _ptr(cv0), _ptr(scratch),
                                                      scratch.numel() * 8, u.shape[0], self.nlat, self.nlon, self.order,
                                                      _stream(self.device)), 'lcs_prefilter')
                coef = (self.time_lerp(cu0, *resample), self.time_lerp(cv0, *resample))
            else:
                coef = None
            u, v = self.time_lerp(u, *resample), self.time_lerp(v, *resample)      # f64, like interp1d's output upstream
            return self._stage_levels(u, v, raw, coef, reuse)
        return self._stage_levels(u, v, raw, None, reuse)

    def _stage_levels(self, u, v, raw, coef, reuse=False):
        """Prefilter (unless ``coef`` = the coefficient planes is given) and pack the device series ``u, v``."""
        nlev = u.shape[0]
        if nlev < 2:
            return StagedWinds(self.layout, self.pair_dtype, nlev)
        tdt = torch.float64 if self.pair_dtype == _lib.LCS_F64 else torch.float32
        shape2 = (self.nlat, self.nlon)
        round32 = 0
        if (self.f32_propagation and self.coords_f64 and u.dtype == torch.float32 and v.dtype == torch.float32 and self.pair_dtype == _lib.LCS_F64
                and self.layout == _lib.LCS_LAYOUT_ES and self.arith == _lib.LCS_ARITH_F64):
            round32 = 1 if self.timestep_weak else 2
        with torch.cuda.device(self.device):
            st = _stream(self.device)
            cu = cv = None
            if self.order >= 2 and coef is not None:
                cu, cv = coef
            elif self.order >= 2:
                cu = self._buf('coef_u', (nlev,) + shape2, torch.float64)
                cv = self._buf('coef_v', (nlev,) + shape2, torch.float64)
                scratch = self._buf('filter_scratch', (2, nlev) + shape2, torch.float64)
                _lib.check(self.lib.lcs_prefilter(_ptr(u), _ptr(v), _dtype_code(u), _ptr(cu), _ptr(cv),
                                                  _ptr(scratch), scratch.numel() * 8,
                                                  nlev, self.nlat, self.nlon, self.order, st), 'lcs_prefilter')
            if self.layout == _lib.LCS_LAYOUT_PAIR4:
                def pack(a, b, code):
                    out = torch.empty((nlev - 1,) + shape2 + (4,), dtype=tdt, device=self.device)
                    _lib.check(self.lib.lcs_pack_pairs(_ptr(a), _ptr(b), code, _ptr(out), self.pair_dtype,
                                                       nlev, self.nlat, self.nlon, st), 'lcs_pack_pairs')
                    return out
                raw = pack(u, v, _dtype_code(u))
                coef = pack(cu, cv, _lib.LCS_F64) if cu is not None else None
                return StagedWinds(self.layout, self.pair_dtype, nlev, raw_a=raw, coef_a=coef)

            def pack_es(a, b, code, tag):
                # halo layout (include/lcs_b200.h): every level carries a mirror-filled rim so no gather reflects an index
                padded = (self.nlat + _lib.LCS_HALO_LO + _lib.LCS_HALO_HI, self.nlon + _lib.LCS_HALO_LO + _lib.LCS_HALO_HI)
                if reuse:
                    e = self._buf(tag + '_e', (nlev,) + padded + (2,), tdt)
                    s_ = self._buf(tag + '_s', (nlev - 1,) + padded + (2,), tdt)
                else:
                    e = torch.empty((nlev,) + padded + (2,), dtype=tdt, device=self.device)
                    s_ = torch.empty((nlev - 1,) + padded + (2,), dtype=tdt, device=self.device)
                _lib.check(self.lib.lcs_pack_es(_ptr(a), _ptr(b), code, _ptr(e), _ptr(s_), self.pair_dtype,
                                                nlev, self.nlat, self.nlon, st), 'lcs_pack_es')
                return e, s_
            if raw not in ('packed', 'planar'):
                raise ValueError("raw must be 'packed' or 'planar'")
            ce_ = cs_ = None
            if cu is not None:
                ce_, cs_ = pack_es(cu, cv, _lib.LCS_F64, 'coef')
                if raw == 'planar':
                    return StagedWinds(self.layout, self.pair_dtype, nlev, raw_a=u, raw_b=v, coef_a=ce_, coef_b=cs_,
                                       raw_planar=True, round32=round32)
            re_, rs_ = pack_es(u, v, _dtype_code(u), 'raw')
            return StagedWinds(self.layout, self.pair_dtype, nlev, raw_a=re_, raw_b=rs_, coef_a=ce_, coef_b=cs_, round32=round32)

    _PIN_MIN_BYTES = 1 << 20

    def _to_device(self, a, slot=0):
        """Host array / tensor -> contiguous device tensor on the current stream.  Pageable host memory of 1 MB or more goes
        through a pinned staging buffer of the engine (one per ``slot``: u and v of a call overlap -- the host's copy of v
        runs while the DMA of u is in flight): the driver's own pageable path moved a C2 series (2 x 6.5 MB) in 0.7 ms."""
        if isinstance(a, torch.Tensor):
            t = a
        else:
            a = np.asarray(a)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        if t.is_cuda or t.is_pinned() or t.numel() * t.element_size() < self._PIN_MIN_BYTES or not t.is_contiguous() \
                or os.environ.get('LCS_PINNED_STAGING', '1') == '0':
            return t.to(self.device, non_blocking=True).contiguous()
        pins = self.__dict__.setdefault('_pins', {})
        buf, done = pins.get(slot, (None, None))
        if buf is None or buf.dtype != t.dtype or buf.numel() < t.numel():
            buf = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
            done = None
        if done is not None:
            done.synchronize()                       # the previous DMA out of this buffer
        view = buf[:t.numel()].view(t.shape)
        view.copy_(t)
        with torch.cuda.device(self.device):
            d = view.to(self.device, non_blocking=True)
            done = torch.cuda.current_stream(self.device).record_event()
        pins[slot] = (buf, done)
        return d

    # ------------------------------------------------------------------ integrator
    def advect(self, staged, nsteps=None, nwindows=1, level0=0, level_stride=1, return_traj=False,
               rows=None, out=None, xrank=None):
        """Run lcs_advect.  ``rows=(r0, r1)`` restricts to a band of particle rows (global indices); ``xrank`` = a
        ``peer.ColumnFlagMail`` when the bands of an outer-clamp integration are spread over several GPUs.

        Returns ``(x, y)`` shaped ``[nwindows, nrow, ncol]`` (+ ``(x_traj, y_traj)`` shaped
        ``[nwindows, nsteps+1, nrow, ncol]`` with ``return_traj``).
        """
        if nsteps is None:
            nsteps = staged.nlev - 1
        if nsteps > 0 and level0 + (nwindows - 1) * level_stride + nsteps > staged.nlev - 1:
            raise ValueError('window runs past the staged wind series')
        r0, r1 = (0, self.part_lat.size) if rows is None else rows
        nrow, ncol = r1 - r0, self.part_lon.size
        with torch.cuda.device(self.device):
            if out is None:
                x = torch.empty((nwindows, nrow, ncol), dtype=torch.float64, device=self.device)
                y = torch.empty_like(x)
            else:
                x, y = out
            xt = yt = None
            if return_traj:
                xt = torch.empty((nwindows, nsteps + 1, nrow, ncol), dtype=torch.float64, device=self.device)
                yt = torch.empty_like(xt)
            part = _lib.Particles(nrow, ncol, r0, self.part_lat.size,
                                  self.d_plat[r0:].data_ptr(), self.d_plon.data_ptr(),
                                  self.d_kx[r0:].data_ptr(), self.d_hx[r0:].data_ptr(), self.ky, self.hy)
            opts = _lib.AdvectOpts(nsteps, self.S, self.order, self.xmode, self.strict,
                                   nwindows, level0, level_stride, self.arith, staged.round32,
                                   C.pointer(xrank.struct) if xrank is not None else None)
            need = self.lib.lcs_advect_workspace_bytes(C.byref(part), C.byref(opts))
            ws = None
            if need:
                if self._ws is None or self._ws.numel() < need:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                ws = self._ws
            winds = staged.struct()
            _lib.check(self.lib.lcs_advect(C.byref(self.grid), C.byref(part), C.byref(opts), C.byref(winds),
                                           _ptr(x), _ptr(y), _ptr(xt), _ptr(yt), _ptr(ws), need,
                                           _stream(self.device)), 'lcs_advect')
        return (x, y, xt, yt) if return_traj else (x, y)

    # ------------------------------------------------------------------ epilogue
    def epilogue(self, x_dep, y_dep, log_scale=False, mask=None, return_jac=False, in_row0=0, out_rows=None, out=None):
        """lcs_ftle_epilogue on ``[nfields, nrow_in, nlon]`` departure points (wind-grid shaped); ``out`` = a device
        tensor ``[nfields, rows, nlon]`` f64 to fill instead of a fresh one."""
        if x_dep.dim() == 2:
            x_dep, y_dep = x_dep[None], y_dep[None]
        nfields, nrow_in, nlon = x_dep.shape
        if nlon != self.nlon:
            raise ValueError('the epilogue needs departure points on the wind grid columns')
        o0, o1 = (0, self.nlat) if out_rows is None else out_rows
        with torch.cuda.device(self.device):
            if out is not None:
                if tuple(out.shape) != (nfields, o1 - o0, nlon) or out.dtype != torch.float64 or not out.is_contiguous():
                    raise ValueError('epilogue(out=): need a contiguous f64 tensor of shape %r' % ((nfields, o1 - o0, nlon),))
                sigma = out
            else:
                sigma = torch.empty((nfields, o1 - o0, nlon), dtype=torch.float64, device=self.device)
            jac = torch.empty((nfields, 6, o1 - o0, nlon), dtype=torch.float64, device=self.device) if return_jac else None
            d_mask = None
            if mask is not None:
                d_mask = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8)).to(self.device)
            # the status word is NOT cleared here: kernels OR into it, so one check_finite() after several epilogue
            # launches (rolling chunks, row bands) sees an inf raised by any of them
            _lib.check(self.lib.lcs_ftle_epilogue(_ptr(x_dep.contiguous()), _ptr(y_dep.contiguous()), nfields,
                                                  self.nlat, nlon, in_row0, nrow_in, o0, o1 - o0,
                                                  _ptr(self.d_dx), self.dy, _ptr(d_mask), int(bool(log_scale)),
                                                  _ptr(sigma), _ptr(jac), _ptr(self.d_status),
                                                  _stream(self.device)), 'lcs_ftle_epilogue')
        return (sigma, jac) if return_jac else sigma

    def gaussian(self, fields, sigma):
        """scipy.ndimage.gaussian_filter(field, sigma) (reflect, truncate=4) on ``[nfields, n0, n1]`` f64 (LCS.py:187-190)."""
        squeeze = fields.dim() == 2
        f = (fields[None] if squeeze else fields).contiguous()
        radius = int(4.0 * float(sigma) + 0.5)
        xs = np.arange(-radius, radius + 1)
        w = np.exp(-0.5 / (float(sigma) * float(sigma)) * xs ** 2)          # scipy _gaussian_kernel1d
        w = w / w.sum()
        with torch.cuda.device(self.device):
            dw = torch.from_numpy(w).to(self.device)
            out, scratch = torch.empty_like(f), torch.empty_like(f)
            _lib.check(self.lib.lcs_gaussian_filter2d(_ptr(f), _ptr(out), _ptr(scratch), f.shape[0], f.shape[1], f.shape[2],
                                                      _ptr(dw), radius, _stream(self.device)), 'lcs_gaussian_filter2d')
        return out[0] if squeeze else out

    def time_lerp(self, series, lo, w_hi, w_lo):
        """Linear time refinement of ``[nlev, nlat, nlon]`` (LCS.py:88-91); see timeaxis.resample_plan."""
        t = self._to_device(series)
        with torch.cuda.device(self.device):
            d_lo = torch.from_numpy(np.ascontiguousarray(lo, dtype=np.int32)).to(self.device)
            d_num = torch.from_numpy(np.ascontiguousarray(w_hi, dtype=np.float64)).to(self.device)
            d_den = torch.from_numpy(np.ascontiguousarray(w_lo, dtype=np.float64)).to(self.device)
            out = torch.empty((len(lo),) + tuple(t.shape[1:]), dtype=torch.float64, device=self.device)
            _lib.check(self.lib.lcs_time_lerp(_ptr(t), _dtype_code(t), _ptr(d_lo), _ptr(d_num), _ptr(d_den), len(lo),
                                              int(t[0].numel()), _ptr(out), _stream(self.device)), 'lcs_time_lerp')
        return out

    def reset_status(self):
        """Clear the inf flag the epilogue launches OR into (call once before a series of epilogue launches)."""
        with torch.cuda.device(self.device):
            self.d_status.zero_()

    def check_finite(self):
        """Raise like scipy.linalg.norm(check_finite=True) does at LCS.py:154 if any epilogue launch since the last
        reset_status() / check_finite() met an inf derivative (synchronises, then clears the flag)."""
        if self._ws is not None and self.xmode == _lib.LCS_X_CLAMP_OUTER:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.lcs_advect_check(_ptr(self._ws), _stream(self.device)), 'lcs_advect')
        status = int(self.d_status.item())
        if status:
            self.d_status.zero_()                               # read-and-clear: the next series starts clean
        if status & 1:
            raise ValueError('array must not contain infs or NaNs')

    # ------------------------------------------------------------------ whole path
    def ftle(self, u, v, log_scale=False, mask=None):
        """prefilter -> advect -> epilogue for one window; returns sigma ``[nlat, nlon]`` (device)."""
        st = self.stage(u, v)
        x, y = self.advect(st)
        return self.epilogue(x, y, log_scale=log_scale, mask=mask)[0]


# ---------------------------------------------------------------------- array-level seams
def map_coordinates_device(field, pos_x, pos_y, lat, lon, order=1, device='cuda:0'):
    """xr_map_coordinates (tools.py:11-41) on the device for one field; returns a f64 tensor."""
    lib = _lib.load()
    device = torch.device(device)
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    grid = _lib.Grid(lat.size, lon.size, lat.min(), lat.max(), lon.min(), lon.max())
    with torch.cuda.device(device):
        f = torch.as_tensor(np.ascontiguousarray(field, dtype=np.float64)).to(device)
        px = torch.as_tensor(np.ascontiguousarray(pos_x, dtype=np.float64)).to(device)
        py = torch.as_tensor(np.ascontiguousarray(pos_y, dtype=np.float64)).to(device)
        nrow, ncol = px.shape
        coef = None
        if order >= 2:
            coef = torch.empty_like(f)
            dummy = torch.empty_like(f)
            scratch = torch.empty((2,) + tuple(f.shape), dtype=torch.float64, device=device)
            _lib.check(lib.lcs_prefilter(_ptr(f), _ptr(f), _lib.LCS_F64, _ptr(coef), _ptr(dummy), _ptr(scratch),
                                         scratch.numel() * 8, 1, lat.size, lon.size, int(order), _stream(device)), 'lcs_prefilter')
        out = torch.empty_like(px)
        _lib.check(lib.lcs_map_coordinates(C.byref(grid), _ptr(f), _ptr(coef), order, _ptr(px), _ptr(py),
                                           nrow, ncol, 0, nrow, _ptr(out), _stream(device)), 'lcs_map_coordinates')
    return out


def _series_on_device(series, device):
    """``[nlev, n0, n1]`` series (numpy array or tensor on any device) -> contiguous f32 / f64 tensor on ``device``."""
    if isinstance(series, torch.Tensor):
        t = series if series.dtype in (torch.float32, torch.float64) else series.to(torch.float64)
        return t.to(device).contiguous()
    a = np.asarray(series)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


_REGRID_PLANS = {}


def regrid_device(series, lat, lon, new_lat, new_lon, device='cuda:0'):
    """LCS.py:108-113 on the device: linear interpolation (latitude, then longitude) of ``[nlev, nlat, nlon]`` to the
    new coordinates, NaNs (outside the source range) filled with the nearest-label value.  Returns an f64 tensor."""
    from .regrid import axis_plan
    lib = _lib.load()
    device = torch.device(device)
    with torch.cuda.device(device):
        t = _series_on_device(series, device)
        plans = []
        for src, dst in ((lat, new_lat), (lon, new_lon)):
            src, dst = np.ascontiguousarray(src, dtype=np.float64), np.ascontiguousarray(dst, dtype=np.float64)
            key = (src.tobytes(), dst.tobytes(), str(device))
            if key not in _REGRID_PLANS:                       # u and v of a call, and every call of a series, share them
                while len(_REGRID_PLANS) >= 8:
                    _REGRID_PLANS.pop(next(iter(_REGRID_PLANS)))
                _REGRID_PLANS[key] = [torch.from_numpy(np.ascontiguousarray(p)).to(device) for p in axis_plan(src, dst)]
            plans += _REGRID_PLANS[key]
        out = torch.empty((t.shape[0], len(new_lat), len(new_lon)), dtype=torch.float64, device=device)
        _lib.check(lib.lcs_regrid_linear_nearest(_ptr(t), _dtype_code(t), t.shape[0], t.shape[1], t.shape[2],
                                                 *[_ptr(p) for p in plans], len(new_lat), len(new_lon), _ptr(out),
                                                 _stream(device)), 'lcs_regrid_linear_nearest')
    return out


_SPECTRAL_TABLES = {}


def spectral_truncate_device(series, ntrunc, device='cuda:0'):
    """LCS.py:115-118 for one component: triangular truncation at ``ntrunc`` of every level of ``[nlev, nlat, nlon]``
    (SPHEREPACK's regular-grid analysis + synthesis, see spectral.py).  Returns an f64 tensor on the device."""
    from .spectral import truncation_tables
    lib = _lib.load()
    device = torch.device(device)
    nlev, nlat, nlon = tuple(series.shape)
    key = (nlat, nlon, int(ntrunc), str(device))
    with torch.cuda.device(device):
        if key not in _SPECTRAL_TABLES:
            A, Fc, Fi = truncation_tables(nlat, nlon, int(ntrunc))
            dev = lambda m: torch.from_numpy(np.ascontiguousarray(m)).to(device)
            _SPECTRAL_TABLES[key] = (dev(A.transpose(0, 2, 1)), dev(Fc), dev(Fi))
        At, Fc, Fi = _SPECTRAL_TABLES[key]
        t = _series_on_device(series, device)
        out = torch.empty((nlev, nlat, nlon), dtype=torch.float64, device=device)
        nbytes = int(lib.lcs_spectral_truncate_scratch_bytes(nlev, nlat, int(ntrunc)))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(lib.lcs_spectral_truncate(_ptr(t), _dtype_code(t), nlev, nlat, nlon, int(ntrunc), _ptr(At), _ptr(Fc), _ptr(Fi),
                                             _ptr(scratch), nbytes, _ptr(out), _stream(device)), 'lcs_spectral_truncate')
    return out


def prefilter_device(u, v, device='cuda:0', order=3):
    """B-spline coefficients (order 2..5) of ``[nlev, nlat, nlon]`` series (f64 tensors on the device)."""
    lib = _lib.load()
    device = torch.device(device)
    with torch.cuda.device(device):
        tu = torch.as_tensor(np.ascontiguousarray(u)).to(device)
        tv = torch.as_tensor(np.ascontiguousarray(v)).to(device)
        nlev, nlat, nlon = tu.shape
        cu = torch.empty(tu.shape, dtype=torch.float64, device=device)
        cv = torch.empty_like(cu)
        scratch = torch.empty((2,) + tuple(cu.shape), dtype=torch.float64, device=device)
        _lib.check(lib.lcs_prefilter(_ptr(tu), _ptr(tv), _dtype_code(tu), _ptr(cu), _ptr(cv), _ptr(scratch),
                                     scratch.numel() * 8, nlev, nlat, nlon, int(order), _stream(device)), 'lcs_prefilter')
    return cu, cv


def fourth_order_derivative_device(arr, dim=0, isglobal=True, device='cuda:0'):
    """fourth_order_derivative (tools.py:190-245) on an f32 array."""
    lib = _lib.load()
    device = torch.device(device)
    with torch.cuda.device(device):
        a = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)).to(device)
        out = torch.empty_like(a)
        _lib.check(lib.lcs_fourth_order_derivative(_ptr(a), a.shape[0], a.shape[1], int(dim), int(bool(isglobal)),
                                                   _ptr(out), _stream(device)), 'lcs_fourth_order_derivative')
    return out


def spectral_norm_3x3_device(vals, device='cuda:0'):
    """scipy.linalg.norm(vals[3,3,N], axis=(0,1), ord=2) (LCS.py:154)."""
    lib = _lib.load()
    device = torch.device(device)
    v = np.ascontiguousarray(vals, dtype=np.float64)
    n = v.shape[-1]
    with torch.cuda.device(device):
        t = torch.as_tensor(v.reshape(9, n)).to(device)
        out = torch.empty(n, dtype=torch.float64, device=device)
        _lib.check(lib.lcs_spectral_norm_3x3(_ptr(t), n, _ptr(out), _stream(device)), 'lcs_spectral_norm_3x3')
    return out
