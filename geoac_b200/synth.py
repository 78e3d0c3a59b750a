"""Synthetic G2S atmospheres of BASELINE.json's configs 3-5, exactly as SURVEY.md section 8(d) specifies them (no RNG).

Host-side input generators for bench.py and the scale tests; they produce what the reference's loaders would hand to
the spline builders (Code/Atmo/G2S_GlobalSpline1D.cpp:113-152, G2S_MultiDimSpline3D.cpp:139-189,
G2S_GlobalMultiDimSpline3D.cpp:142-199): winds in km/s with the ground taper applied, density in g/cm^3.
"""
import numpy as np

R_GAS, G0, P0 = 287.05, 9.8, 101325.0


def temperature(z):
    """Piecewise-linear US-Std-1976 to 91 km, then an exponential thermosphere (574 K at 150 km)."""
    z = np.asarray(z, dtype=np.float64)
    zb = [0.0, 11.0, 20.0, 32.0, 47.0, 51.0, 71.0, 84.852, 91.0]
    lapse = [-6.5, 0.0, 1.0, 2.8, 0.0, -2.8, -2.0, 0.0]
    T = np.empty_like(z)
    Tb = 288.15
    for i in range(len(lapse)):
        m = (z >= zb[i]) & (z <= zb[i + 1]) if i == 0 else (z > zb[i]) & (z <= zb[i + 1])
        T[m] = Tb + lapse[i] * (z[m] - zb[i])
        Tb = Tb + lapse[i] * (zb[i + 1] - zb[i])
    m = z > 91.0
    T[m] = Tb + 450.0 * (1.0 - np.exp(-(((z[m] - 91.0) / 42.0) ** 2)))
    return T


def base_profile(z):
    """z [km] -> T [K], u, v [m/s], rho [g/cm^3], p [mbar] (file units of a `zTuvdp` .met profile)."""
    z = np.asarray(z, dtype=np.float64)
    T = temperature(z)
    u = -60.0 * np.exp(-(((z - 55.0) / 12.0) ** 2)) + 30.0 * np.exp(-(((z - 110.0) / 15.0) ** 2))
    v = 10.0 * np.exp(-(((z - 15.0) / 5.0) ** 2))
    p = np.empty_like(z)
    p[0] = P0
    for i in range(len(z) - 1):                                    # hydrostatic integration between levels
        Tbar = 0.5 * (T[i] + T[i + 1])
        p[i + 1] = p[i] * np.exp(-G0 * (z[i + 1] - z[i]) * 1000.0 / (R_GAS * Tbar))
    rho = p / (R_GAS * T)
    return T, u, v, rho / 1000.0, p / 100.0


def config3_profile():
    """Stratified profile of config 3: z = 0 ... 150 km step 0.1 (1501 rows).  Returns the six .met columns."""
    z = np.round(np.arange(1501) * 0.1, 1)
    return (z,) + base_profile(z)


def write_met(path, cols):
    np.savetxt(path, np.column_stack(cols), fmt="%.1f %.6f %.6f %.6f %.6e %.6e")


def _taper(z, width):
    return (2.0 / (1.0 + np.exp(-(z - 0.0) / width)) - 1.0) / 1000.0     # m/s -> km/s and ground taper (z_grnd = 0 at load)


def config4_grid(nx=200, ny=200, nz=300, x=None, y=None):
    """Cartesian range-dependent grid of config 4: x, y = -500 ... 500 km (or the given node coordinates),
    z = 0 ... (nz-1)/2 km; fields [nx][ny][nz]."""
    x = np.linspace(-500.0, 500.0, nx) if x is None else np.asarray(x, dtype=np.float64)
    y = np.linspace(-500.0, 500.0, ny) if y is None else np.asarray(y, dtype=np.float64)
    z = np.arange(nz) * 0.5
    T0, u0, v0, rho0, _ = base_profile(z)
    X, Y = x[:, None, None], y[None, :, None]
    tap = _taper(z, 0.05)[None, None, :]
    T = T0[None, None, :] * (1.0 + 0.02 * np.sin(2 * np.pi * X / 700.0) * np.cos(2 * np.pi * Y / 900.0))
    u = u0[None, None, :] * (1.0 + 0.2 * np.cos(2 * np.pi * X / 600.0)) * tap * np.ones_like(Y)
    v = (v0[None, None, :] + 8.0 * np.sin(2 * np.pi * Y / 800.0) * np.exp(-(((z - 50.0) / 20.0) ** 2))[None, None, :]) * tap * np.ones_like(X)
    rho = np.broadcast_to(rho0[None, None, :], T.shape)
    return x, y, z, np.ascontiguousarray(T), np.ascontiguousarray(u), np.ascontiguousarray(v), np.ascontiguousarray(rho)


def config5_grid(nlat=181, nlon=361, nr=300, lat_deg=None, lon_deg=None):
    """Global range-dependent grid of config 5: lat -90 ... 90, lon -180 ... 180 (or the given node coordinates in degrees;
    returned in radians, as the loader converts them), altitude 0 ... (nr-1)/2 km; fields [nlat][nlon][nr]."""
    lat = np.radians(np.linspace(-90.0, 90.0, nlat) if lat_deg is None else np.asarray(lat_deg, dtype=np.float64))
    lon = np.radians(np.linspace(-180.0, 180.0, nlon) if lon_deg is None else np.asarray(lon_deg, dtype=np.float64))
    z = np.arange(nr) * 0.5
    T0, u0, v0, rho0, _ = base_profile(z)
    LA, LO = lat[:, None, None], lon[None, :, None]
    tap = _taper(z, 0.2)[None, None, :]
    T = T0[None, None, :] * (1.0 + 0.02 * np.sin(3.0 * LA) * np.cos(2.0 * LO))
    u = u0[None, None, :] * np.cos(LA) ** 2 * (1.0 + 0.2 * np.cos(4.0 * LO)) * tap
    v = (v0[None, None, :] + 8.0 * np.sin(5.0 * LO) * np.exp(-(((z - 50.0) / 20.0) ** 2))[None, None, :]) * tap * np.ones_like(LA)
    rho = np.broadcast_to(rho0[None, None, :], T.shape)
    return lat, lon, z, np.ascontiguousarray(T), np.ascontiguousarray(u), np.ascontiguousarray(v), np.ascontiguousarray(rho)


def _roundtrip(vals, fmt):
    """What strtod returns for the text `fmt % value` -- exactly (a few hundred values: node coordinates, levels, density)."""
    return np.array([float(fmt % v) for v in np.asarray(vals, dtype=np.float64)])


def config4_grid_from_files(nx=200, ny=200, nz=300):
    """The config-4 grid EXACTLY as the reference's loader hands it to the spline builder when it reads the node files that
    write_config4_files() writes: node fields rounded to the files' six decimals, node coordinates / levels / density through the
    text round trip, winds converted to km/s with the loader's ground taper (Code/Atmo/G2S_MultiDimSpline3D.cpp:139-189).
    Bit-identical to load_met_grid() on those 40 000 files (checked when the golden vector was made, tests/golden/make_golden.py),
    without writing them: this is the grid the full-size reference golden `3drngdep_c4full` was traced on."""
    import math
    xs, ys = np.linspace(-500.0, 500.0, nx), np.linspace(-500.0, 500.0, ny)
    z = np.arange(nz) * 0.5
    T0, u0, v0, rho0, _ = base_profile(z)
    # per-node factors with the SAME scalar calls the file writer makes (_raw_cart)
    fT = np.array([[1.0 + 0.02 * np.sin(2 * np.pi * x / 700.0) * np.cos(2 * np.pi * y / 900.0) for y in ys] for x in xs])
    fu = np.array([1.0 + 0.2 * np.cos(2 * np.pi * x / 600.0) for x in xs])
    sv = np.array([8.0 * np.sin(2 * np.pi * y / 800.0) for y in ys])
    gz = np.exp(-(((z - 50.0) / 20.0) ** 2))
    T = T0[None, None, :] * fT[:, :, None]
    u = np.broadcast_to((u0[None, :] * fu[:, None])[:, None, :], T.shape)
    v = np.broadcast_to((v0[None, :] + sv[:, None] * gz[None, :])[None, :, :], T.shape)
    q = lambda a: np.rint(a * 1e6) / 1e6                                  # "%.6f" and back
    zq = _roundtrip(z, "%.1f")
    tap = np.array([(2.0 / (1.0 + math.exp(-(zz - 0.0) / 0.05)) - 1.0) / 1000.0 for zz in zq])
    rho = np.broadcast_to(_roundtrip(rho0 / 1.0, "%.6e")[None, None, :], T.shape)
    return (_roundtrip(xs, "%.6f"), _roundtrip(ys, "%.6f"), zq, np.ascontiguousarray(q(T)),
            np.ascontiguousarray(q(u) * tap[None, None, :]), np.ascontiguousarray(q(v) * tap[None, None, :]), np.ascontiguousarray(rho))


# ---- the same atmospheres as .met node files (file units: m/s, g/cm^3, mbar), for the reference binaries / golden vectors ----
def _raw_cart(x, y, z):
    T0, u0, v0, rho0, p0 = base_profile(z)
    T = T0 * (1.0 + 0.02 * np.sin(2 * np.pi * x / 700.0) * np.cos(2 * np.pi * y / 900.0))
    u = u0 * (1.0 + 0.2 * np.cos(2 * np.pi * x / 600.0))
    v = v0 + 8.0 * np.sin(2 * np.pi * y / 800.0) * np.exp(-(((z - 50.0) / 20.0) ** 2))
    return T, u, v, rho0, p0


def _raw_glob(lat, lon, z):
    T0, u0, v0, rho0, p0 = base_profile(z)
    T = T0 * (1.0 + 0.02 * np.sin(3.0 * lat) * np.cos(2.0 * lon))
    u = u0 * np.cos(lat) ** 2 * (1.0 + 0.2 * np.cos(4.0 * lon))
    v = v0 + 8.0 * np.sin(5.0 * lon) * np.exp(-(((z - 50.0) / 20.0) ** 2))
    return T, u, v, rho0, p0


def write_config4_files(outdir, xs, ys, nz=300, prefix="p"):
    """Node files of a config-4 style grid on the given node coordinates (km): <prefix><ix*ny+iy>.met, x.loc, y.loc."""
    import os
    z = np.arange(nz) * 0.5
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, "x.loc"), xs, fmt="%.6f")
    np.savetxt(os.path.join(outdir, "y.loc"), ys, fmt="%.6f")
    for ix, x in enumerate(xs):
        for iy, y in enumerate(ys):
            write_met(os.path.join(outdir, f"{prefix}{ix * len(ys) + iy}.met"), (z,) + _raw_cart(x, y, z))
    return os.path.join(outdir, prefix), os.path.join(outdir, "x.loc"), os.path.join(outdir, "y.loc")


def write_config5_files(outdir, lats_deg, lons_deg, nz=300, prefix="p"):
    """Node files of a config-5 style grid on the given node coordinates (degrees): <prefix><it*np+ip>.met, lat.loc, lon.loc."""
    import os
    z = np.arange(nz) * 0.5
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, "lat.loc"), lats_deg, fmt="%.6f")
    np.savetxt(os.path.join(outdir, "lon.loc"), lons_deg, fmt="%.6f")
    for it, la in enumerate(lats_deg):
        for ip, lo in enumerate(lons_deg):
            write_met(os.path.join(outdir, f"{prefix}{it * len(lons_deg) + ip}.met"), (z,) + _raw_glob(np.radians(la), np.radians(lo), z))
    return os.path.join(outdir, prefix), os.path.join(outdir, "lat.loc"), os.path.join(outdir, "lon.loc")
