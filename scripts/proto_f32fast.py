"""CPU prototype of the f32 fast-path gather variants (numpy emulation of the device arithmetic), used to decide which
variant is worth GPU time: how many FTLE values stay within 1e-5 / 1e-4 relative of the f64 path on the case of
tests/test_gpu_engine.py::test_f32_arithmetic_fast_path_tolerance.  Pointwise x-clamp, ES layout, cubic.
Run: python scripts/proto_f32fast.py"""
import sys, os
import numpy as np
from scipy.ndimage import spline_filter

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S

f32 = np.float32


def fma32(a, b, c):          # f32 fma: exact product in f64, one f64 add, rounded to f32 (double rounding is negligible here)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def weights64(t):
    z = 1.0 - t
    return [z * z * z / 6.0, (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0, (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0, t * t * t / 6.0]


def weights32(t):            # cubic_weights_f32 of lcs_device.cuh
    t = t.astype(f32); z = f32(1) - t; s = f32(1) / f32(6)
    w1 = (t * t * (t - f32(2)) * f32(3) + f32(4)) * s
    w2 = (z * z * (z - f32(2)) * f32(3) + f32(4)) * s
    w0 = z * z * z * s
    return [w0, w1, w2, f32(1) - w0 - w1 - w2]


def weights32h(t):           # Horner form in f32: no 1 - sum, every weight direct
    t = t.astype(f32); z = f32(1) - t; s = f32(1) / f32(6); c23 = f32(2) / f32(3)
    ty = t * t; tz = z * z
    w1 = fma32(fma32(f32(0.5) * np.ones_like(t), t, -np.ones_like(t)), ty, c23 * np.ones_like(t))
    w2 = fma32(fma32(f32(0.5) * np.ones_like(t), z, -np.ones_like(t)), tz, c23 * np.ones_like(t))
    return [(z * s) * tz, w1, w2, (t * s) * ty]


def halo(c):                 # mirror-filled rim: 2 before, 3 after
    return np.pad(c, ((2, 3), (2, 3)), mode='reflect')


def sample(H, iy, ix, nlat, nlon, variant):
    """H: [2][nlat+5][nlon+5] halo level (u, v); returns (su, sv) f64."""
    cy = np.where(iy < 0, iy + (nlat - 1), np.where(iy > nlat - 1, iy - (nlat - 1), iy))
    cx = np.where(ix < 0, ix + (nlon - 1), np.where(ix > nlon - 1, ix - (nlon - 1), ix))
    fy = np.floor(cy); fx = np.floor(cx)
    sy = fy.astype(int) - 1 + 2; sx = fx.astype(int) - 1 + 2
    ty = cy - fy; tx = cx - fx
    out = []
    if variant in ('f64', 'f32store'):
        wy = weights64(ty); wx = weights64(tx)
        for f in H:
            acc = 0.0
            for i in range(4):
                r = 0.0
                for j in range(4):
                    r = r + f[sy + i, sx + j].astype(np.float64) * wx[j]
                acc = acc + r * wy[i]
            out.append(acc)
        return out
    if variant in ('w64r',):           # weights in f64, rounded to f32
        wy = [w.astype(f32) for w in weights64(ty)]; wx = [w.astype(f32) for w in weights64(tx)]
    else:
        wy = weights32(ty); wx = weights32(tx)
    anomaly = variant in ('anom', 'anom_w64r', 'anom_h32')
    if variant in ('anom_h32', 'h32'):
        wy = weights32h(ty); wx = weights32h(tx)
    if variant == 'anom_w64r':
        wy = [w.astype(f32) for w in weights64(ty)]; wx = [w.astype(f32) for w in weights64(tx)]
    for f in H:
        ref = f[sy + 1, sx + 1] if anomaly else None
        acc = np.zeros(iy.shape, f32)
        for i in range(4):
            c = [f[sy + i, sx + j] for j in range(4)]
            if anomaly:
                c = [(cj - ref).astype(f32) for cj in c]
            r = (c[0] * wx[0]).astype(f32)
            for j in range(1, 4):
                r = fma32(c[j], wx[j], r)
            acc = fma32(r, wy[i], acc)
        if anomaly:
            out.append(ref.astype(np.float64) + acc.astype(np.float64))
        else:
            out.append(acc.astype(np.float64))
    return out


def integrate(u, v, lat, lon, dt, Sord, variant):
    nt, nlat, nlon = u.shape
    cu = np.stack([spline_filter(u[k], order=3, mode='mirror') for k in range(nt)])
    cv = np.stack([spline_filter(v[k], order=3, mode='mirror') for k in range(nt)])
    st = np.float64 if variant == 'f64' else f32
    off = np.zeros(2)
    if variant.endswith('_off'):                       # store c - series mean in f32, add the mean back in f64 (weights sum to 1)
        variant = variant[:-4]
        off = np.array([u.mean(), v.mean()])
        cu = cu - off[0]; cv = cv - off[1]
    E = [np.stack([halo(cu[k]), halo(cv[k])]).astype(st) for k in range(nt)]
    Sx = [np.stack([halo(2 * cu[k] - cu[k + 1]), halo(2 * cv[k] - cv[k + 1])]).astype(st) for k in range(nt - 1)]
    cx_, cy_ = O.conversions(lat)
    kx = (dt * cx_)[:, None]; hx = (0.5 * dt * cx_)[:, None]; ky = dt * cy_; hy = 0.5 * dt * cy_
    x, y = np.meshgrid(lon, lat)
    lat_min, lat_max, lon_min, lon_max = lat.min(), lat.max(), lon.min(), lon.max()
    a = nlat / (lat_max - lat_min); b = nlon / (lon_max - lon_min)

    def smp(H):
        r = sample(H, (y - lat_min) * a, (x - lon_min) * b, nlat, nlon, variant)
        return r[0] + off[0], r[1] + off[1]

    def bounds(x, y):
        y = np.where(y > lat_min, y, lat_min); y = np.where(y < lat_max, y, lat_max)
        return np.clip(x, lon_min, lon_max), y
    for k in range(nt - 1):
        ua, va = smp(E[k])
        y = y + ky * va; x = x + kx * ua
        x, y = bounds(x, y)
        for _ in range(Sord):
            su, sv = smp(Sx[k])
            y = y + hy * (va + sv); x = x + hx * (ua + su)
            x, y = bounds(x, y)
    return x, y


def main():
    lat = np.linspace(-40.0, 0.0, 161)
    lon = np.linspace(-80.0, -30.0, 201)
    u, v = S.era5_like_winds(lat, lon, 9)
    print('wind magnitude', np.abs(u).max(), np.abs(v).max(), 'mean', u.mean(), v.mean())
    rx, ry = integrate(u, v, lat, lon, -21600, 4, 'f64')
    inner = slice(3, -3)                                # pole rows are sampled as interior here: leave them out
    ref = O.spectral_norm_field(O.flowmap_gradient(rx, ry, lat, lon))[5:-5]
    good = ref > 1e-6
    fref = 0.5 * np.log(ref[good])
    for variant in sys.argv[1:] or ('f32store', 'f32fast', 'w64r', 'anom', 'anom_w64r'):
        x, y = integrate(u, v, lat, lon, -21600, 4, variant)
        sig = O.spectral_norm_field(O.flowmap_gradient(x, y, lat, lon))[5:-5]
        frel = np.abs(0.5 * np.log(sig[good]) - fref) / np.maximum(np.abs(fref), 1e-3)
        print(f'{variant:10s} pos err x {np.abs(x - rx).max() / 80:.2e} y {np.abs(y - ry).max() / 40:.2e}  '
              f'FTLE within 1e-5: {(frel <= 1e-5).mean():.4f}  within 1e-4: {(frel <= 1e-4).mean():.4f}')


if __name__ == '__main__':
    main()
