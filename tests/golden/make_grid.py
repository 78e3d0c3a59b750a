"""Synthetic range-dependent G2S grids for the golden vectors (run in the build container only).

Writes one `.met` profile per horizontal node plus the two node-coordinate files in the layout the reference's
Spline_Multi_G2S expects (Code/Atmo/G2S_MultiDimSpline3D.cpp:109-189: `<prefix><ix*ny+iy>.met`, x/y node files;
Global: `<prefix><it*np+ip>.met`, lat/lon node files in degrees).  The node profile is ToyAtmo.met (every `zstep`-th row)
with the smooth analytic perturbations of SURVEY section 8d, config 4 / 5.  No RNG.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TOY = os.path.join(HERE, "ToyAtmo.met")


def base_profile(zstep):
    d = np.loadtxt(TOY)[::zstep]
    return d[:, 0], d[:, 1], d[:, 2], d[:, 3], d[:, 4], d[:, 5]          # z T u v rho p


def write_cartesian(outdir, xs, ys, zstep=10, prefix="p"):
    z, T, u, v, rho, p = base_profile(zstep)
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, "x.loc"), xs, fmt="%.6f")
    np.savetxt(os.path.join(outdir, "y.loc"), ys, fmt="%.6f")
    for ix, x in enumerate(xs):
        for iy, y in enumerate(ys):
            Tn = T * (1.0 + 0.02 * np.sin(2 * np.pi * x / 700.0) * np.cos(2 * np.pi * y / 900.0))
            un = u * (1.0 + 0.2 * np.cos(2 * np.pi * x / 600.0))
            vn = v + 8.0 * np.sin(2 * np.pi * y / 800.0) * np.exp(-((z - 50.0) / 20.0) ** 2)
            rows = np.column_stack([z, Tn, un, vn, rho, p])
            np.savetxt(os.path.join(outdir, f"{prefix}{ix * len(ys) + iy}.met"), rows, fmt="%.1f %.6f %.6f %.6f %.6e %.6e")
    return os.path.join(outdir, prefix), os.path.join(outdir, "x.loc"), os.path.join(outdir, "y.loc")


def write_global(outdir, lats_deg, lons_deg, zstep=10, prefix="p"):
    z, T, u, v, rho, p = base_profile(zstep)
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, "lat.loc"), lats_deg, fmt="%.6f")
    np.savetxt(os.path.join(outdir, "lon.loc"), lons_deg, fmt="%.6f")
    for it, la in enumerate(lats_deg):
        for ip, lo in enumerate(lons_deg):
            lar, lor = np.radians(la), np.radians(lo)
            Tn = T * (1.0 + 0.02 * np.sin(3.0 * lar) * np.cos(2.0 * lor))
            un = u * np.cos(lar) ** 2 * (1.0 + 0.2 * np.cos(4.0 * lor))
            vn = v + 8.0 * np.sin(5.0 * lor) * np.exp(-((z - 50.0) / 20.0) ** 2)
            rows = np.column_stack([z, Tn, un, vn, rho, p])
            np.savetxt(os.path.join(outdir, f"{prefix}{it * len(lons_deg) + ip}.met"), rows, fmt="%.1f %.6f %.6f %.6f %.6e %.6e")
    return os.path.join(outdir, prefix), os.path.join(outdir, "lat.loc"), os.path.join(outdir, "lon.loc")
