"""TEST-ONLY: drive tests/host_emul (the device per-ray code compiled with g++) -- a debugging aid for the build
container, which has no GPU.  Not a fallback: nothing in geoac_b200/ imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

from geoac_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "host_emul.cpp")
SO = os.path.join(HERE, "host_emul", "libhost_emul.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def build(force=False):
    csrc = os.path.join(HERE, "..", "geoac_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast", SRC, "-o", SO])
    return SO


def natural_slopes(x, f):
    n = len(x)
    cp, dpv, s = np.zeros(n), np.zeros(n), np.zeros(n)
    di = 2.0 / (x[1] - x[0]); up = 1.0 / (x[1] - x[0])
    rh = 3.0 * (f[1] - f[0]) / (x[1] - x[0]) ** 2
    cp[0] = up / di; dpv[0] = rh / di
    for i in range(1, n - 1):
        lo = 1.0 / (x[i] - x[i - 1])
        di = 2.0 * (1.0 / (x[i] - x[i - 1]) + 1.0 / (x[i + 1] - x[i]))
        up = 1.0 / (x[i + 1] - x[i])
        rh = 3.0 * ((f[i] - f[i - 1]) / (x[i] - x[i - 1]) ** 2 + (f[i + 1] - f[i]) / (x[i + 1] - x[i]) ** 2)
        cp[i] = up / (di - cp[i - 1] * lo)
        dpv[i] = (rh - dpv[i - 1] * lo) / (di - cp[i - 1] * lo)
    lo = 1.0 / (x[n - 1] - x[n - 2]); di = 2.0 / (x[n - 1] - x[n - 2])
    rh = 3.0 * (f[n - 1] - f[n - 2]) / (x[n - 1] - x[n - 2]) ** 2
    dpv[n - 1] = (rh - dpv[n - 2] * lo) / (di - cp[n - 2] * lo)
    s[n - 1] = dpv[n - 1]
    for i in range(n - 2, -1, -1):
        s[i] = dpv[i] - cp[i] * s[i + 1]
    return s


def make_table(is_global, z, T, u, v, rho):
    x = np.array(z, dtype=np.float64) + (6370.0 if is_global else 0.0)
    n = len(x)
    tab = np.zeros((n, 10))                      # one record per level: x, invh, T, sT, u, su, v, sv, rho, srho (core.cuh)
    tab[:, 0] = x
    tab[: n - 1, 1] = 1.0 / (x[1:] - x[:-1])
    for col, f in ((2, T), (4, u), (6, v), (8, rho)):
        tab[:, col] = f
        tab[:, col + 1] = natural_slopes(x, np.asarray(f, dtype=np.float64))
    return tab, n


def _paths(nr, stride, cap):
    path = np.zeros((nr, max(cap, 1), abi.PATH_NF)); rows = np.zeros(nr, dtype=np.int32)
    return path, rows


def _caus(nr, cap):
    return np.zeros((nr, max(cap, 1), abi.CAUSTIC_NF)), np.zeros(nr, dtype=np.int32)


def trace(variant, params, atmo_arrays, theta, phi, path_stride=0, path_cap=0, caus_cap=0):
    L = C.CDLL(build())
    L.emul_trace_1d.restype = C.c_long
    L.emul_trace_1d.argtypes = [C.c_int, C.POINTER(abi.GeoacParams), C.c_int, dp, C.c_long, dp, dp, dp, ip, ip, C.c_int, C.c_long, dp, ip, C.c_long, dp, ip]
    tab, n = make_table(variant == abi.GEOAC_GLOBAL, *atmo_arrays)
    theta = np.ascontiguousarray(theta, dtype=np.float64); phi = np.ascontiguousarray(phi, dtype=np.float64)
    nr = len(theta); n_rec = params.bounces + 1
    rec = np.zeros((abi.NFIELDS, nr, n_rec)); status = np.zeros((nr, n_rec), dtype=np.int32); n_steps = np.zeros((nr, n_rec), dtype=np.int32)
    path, rows = _paths(nr, path_stride, path_cap)
    caus, crow = _caus(nr, caus_cap)
    total = L.emul_trace_1d(variant, C.byref(params), n, tab.ctypes.data_as(dp), nr, theta.ctypes.data_as(dp), phi.ctypes.data_as(dp),
                            rec.ctypes.data_as(dp), status.ctypes.data_as(ip), n_steps.ctypes.data_as(ip),
                            path_stride, path_cap, path.ctypes.data_as(dp), rows.ctypes.data_as(ip), caus_cap, caus.ctypes.data_as(dp), crow.ctypes.data_as(ip))
    return {"rec": rec, "status": status, "n_steps": n_steps, "total_steps": total, "path": path, "path_rows": rows, "caustic": caus, "caustic_rows": crow}


def trace_grid(variant, params, grid_arrays, theta, phi, path_stride=0, path_cap=0, caus_cap=0):
    """Range-dependent variants: grid_arrays = ax0, ax1, axz, T, u, v, rho as geoac_set_atmosphere_3d takes them."""
    L = C.CDLL(build())
    L.emul_trace_3d.restype = C.c_long
    L.emul_trace_3d.argtypes = [C.c_int, C.POINTER(abi.GeoacParams), C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_long, dp, dp, dp, ip, ip,
                                C.c_int, C.c_long, dp, ip, C.c_long, dp, ip]
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in grid_arrays]
    theta = np.ascontiguousarray(theta, dtype=np.float64); phi = np.ascontiguousarray(phi, dtype=np.float64)
    nr = len(theta); n_rec = params.bounces + 1
    rec = np.zeros((abi.NFIELDS, nr, n_rec)); status = np.zeros((nr, n_rec), dtype=np.int32); n_steps = np.zeros((nr, n_rec), dtype=np.int32)
    path, rows = _paths(nr, path_stride, path_cap)
    caus, crow = _caus(nr, caus_cap)
    total = L.emul_trace_3d(variant, C.byref(params), len(arrs[0]), len(arrs[1]), len(arrs[2]), *[a.ctypes.data_as(dp) for a in arrs], nr,
                            theta.ctypes.data_as(dp), phi.ctypes.data_as(dp), rec.ctypes.data_as(dp), status.ctypes.data_as(ip), n_steps.ctypes.data_as(ip),
                            path_stride, path_cap, path.ctypes.data_as(dp), rows.ctypes.data_as(ip), caus_cap, caus.ctypes.data_as(dp), crow.ctypes.data_as(ip))
    return {"rec": rec, "status": status, "n_steps": n_steps, "total_steps": total, "path": path, "path_rows": rows, "caustic": caus, "caustic_rows": crow}
