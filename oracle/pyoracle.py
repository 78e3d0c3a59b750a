"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/liboracle.so) and readers for the dumps written by
oracle/_ref/ref_<variant>.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package geoac_b200 never does."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from geoac_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_atmo1d_create.restype = C.c_void_p
        L.orc_atmo1d_create.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.orc_atmo3d_create.restype = C.c_void_p
        L.orc_atmo3d_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.orc_atmo_destroy.argtypes = [C.c_void_p]
        L.orc_atmo_sample.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, dp]
        L.orc_atmo3d_slopes.restype = C.c_int64
        L.orc_atmo3d_slopes.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_trace.restype = C.c_int64
        L.orc_trace.argtypes = [C.c_int, C.c_void_p, C.POINTER(abi.GeoacParams), C.c_int64, dp, dp, dp, ip, ip]
        L.orc_set_prop_region.argtypes = [C.c_int, C.c_void_p, C.POINTER(abi.GeoacParams)]
        L.geoac_default_params_oracle.argtypes = [C.c_int, C.POINTER(abi.GeoacParams)]
        L.orc_load_met_1d.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), dp, dp, dp, dp, dp]
        L.orc_load_met_grid.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp, dp, dp, dp, dp, dp]
        L.orc_suthbass_alpha.restype = C.c_double
        L.orc_suthbass_alpha.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(dp)


def load_met_1d(path, fmt="zTuvdp", z_grnd_taper=0.0, global_taper=False, cap=200000):
    arrs = [np.zeros(cap) for _ in range(5)]
    n = C.c_int(0)
    rc = lib().orc_load_met_1d(path.encode(), fmt.encode(), z_grnd_taper, int(global_taper), cap, C.byref(n), *[_p(a) for a in arrs])
    if rc != 0:
        raise IOError(f"orc_load_met_1d({path}) -> {rc}")
    return [a[: n.value].copy() for a in arrs]   # z, T, u, v, rho


def load_met_grid(prefix, loc0, loc1, fmt="zTuvdp", is_global=False, cap0=512, cap1=512, capz=4096):
    """Load_G2S_Multi mirror: returns ax0, ax1, axz, T, u, v, rho (fields [n0][n1][nz], winds tapered, km/s)."""
    import re
    n0g = len(open(loc0).read().split())
    n1g = len(open(loc1).read().split())
    nzg = sum(1 for _ in open(f"{prefix}0.met"))
    ax0, ax1, axz = np.zeros(n0g), np.zeros(n1g), np.zeros(nzg)
    fields = [np.zeros(n0g * n1g * nzg) for _ in range(4)]
    n0, n1, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().orc_load_met_grid(prefix.encode(), loc0.encode(), loc1.encode(), fmt.encode(), int(is_global), n0g, n1g, nzg,
                                 C.byref(n0), C.byref(n1), C.byref(nz), _p(ax0), _p(ax1), _p(axz), *[_p(f) for f in fields])
    if rc != 0 or (n0.value, n1.value, nz.value) != (n0g, n1g, nzg):
        raise IOError(f"orc_load_met_grid({prefix}) -> {rc} ({n0.value},{n1.value},{nz.value})")
    return [ax0, ax1, axz] + [f.reshape(n0g, n1g, nzg) for f in fields]


class Atmo:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_atmo_destroy(self.h)
            self.h = None


def atmo1d(is_global, z, T, u, v, rho):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (z, T, u, v, rho)]
    return Atmo(lib().orc_atmo1d_create(int(is_global), len(arrs[0]), *[_p(a) for a in arrs]))


def atmo3d(is_global, ax0, ax1, axz, T, u, v, rho):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (ax0, ax1, axz, T, u, v, rho)]
    h = lib().orc_atmo3d_create(int(is_global), len(arrs[0]), len(arrs[1]), len(arrs[2]), *[_p(a) for a in arrs])
    if not h:
        raise RuntimeError("orc_atmo3d_create failed")
    return Atmo(h)


def atmo_sample(atmo, p0, p1, p2):
    """c, u, v, rho at a point."""
    out = np.zeros(4)
    lib().orc_atmo_sample(atmo.h, p0, p1, p2, _p(out))
    return out


def atmo3d_slopes(atmo, shape):
    """Set_Slopes_Multi output of the oracle: array [field T,u,v,rho][values, z-slopes, d/dax0 slopes, d/dax1 slopes][n0][n1][nz]."""
    out = np.empty((4, 4) + tuple(shape))
    for f in range(4):
        for w in range(4):
            assert lib().orc_atmo3d_slopes(atmo.h, f, w, _p(out[f, w])) == out[f, w].size
    return out


def default_params(variant, atmo=None):
    p = abi.GeoacParams()
    lib().geoac_default_params_oracle(variant, C.byref(p))
    if atmo is not None:
        lib().orc_set_prop_region(variant, atmo.h, C.byref(p))
    return p


def trace_paths(variant, atmo, params, theta, phi, stride, cap, caus_cap=0):
    """As trace(), plus the raypath rows (accum_per_segment semantics): path [n][cap][PATH_NF], path_rows [n]."""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    n = len(theta)
    n_rec = params.bounces + 1
    rec = np.zeros((abi.NFIELDS, n, n_rec))
    status = np.zeros((n, n_rec), dtype=np.int32)
    n_steps = np.zeros((n, n_rec), dtype=np.int32)
    path = np.zeros((n, max(cap, 1), abi.PATH_NF)); rows = np.zeros(n, dtype=np.int32)
    caus = np.zeros((n, max(caus_cap, 1), abi.CAUSTIC_NF)); crow = np.zeros(n, dtype=np.int32)
    L = lib()
    L.orc_trace_paths.restype = C.c_int64
    L.orc_trace_paths.argtypes = [C.c_int, C.c_void_p, C.POINTER(abi.GeoacParams), C.c_int64, dp, dp, dp, ip, ip, C.c_int, C.c_int64, dp, ip, C.c_int64, dp, ip]
    total = L.orc_trace_paths(variant, atmo.h, C.byref(params), n, _p(theta), _p(phi), _p(rec), status.ctypes.data_as(ip),
                              n_steps.ctypes.data_as(ip), stride, cap, _p(path), rows.ctypes.data_as(ip), caus_cap, _p(caus), crow.ctypes.data_as(ip))
    if total < 0:
        raise RuntimeError("orc_trace_paths failed")
    return {"rec": rec, "status": status, "n_steps": n_steps, "total_steps": int(total), "path": path, "path_rows": rows,
            "caustic": caus, "caustic_rows": crow}


def trace(variant, atmo, params, theta, phi):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    n = len(theta)
    n_rec = params.bounces + 1
    rec = np.zeros((abi.NFIELDS, n, n_rec))
    status = np.zeros((n, n_rec), dtype=np.int32)
    n_steps = np.zeros((n, n_rec), dtype=np.int32)
    total = lib().orc_trace(variant, atmo.h, C.byref(params), n, _p(theta), _p(phi), _p(rec),
                            status.ctypes.data_as(ip), n_steps.ctypes.data_as(ip))
    if total < 0:
        raise RuntimeError("orc_trace failed")
    return {"rec": rec, "status": status, "n_steps": n_steps, "total_steps": int(total)}


# ---------------------------------------------------------------- reference dumps (oracle/_ref/ref_<variant>)
REF_NF = 32
REF_F_STATUS, REF_F_NSTEPS = 26, 27


def ref_binary(variant):
    return os.path.join(HERE, "_ref", "ref_" + abi.VARIANT_NAMES[variant])


def run_ref(variant, profile_args, out_path, **kv):
    """Run the unmodified reference through oracle/ref_driver.cpp; returns (records dict, timing json)."""
    cmd = [ref_binary(variant), out_path] + list(profile_args) + [f"{k}={v}" for k, v in kv.items()]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    info = json.loads(out.strip().splitlines()[-1])
    ref = read_ref_bin(out_path)
    if int(kv.get("path_stride", 0)) > 0:
        ref["path"] = np.fromfile(out_path + ".path", dtype=np.float64).reshape(-1, 9)     # ray, state[0..2], amp, att, tt, bounce, step
    if int(kv.get("caustics", 0)) > 0:
        ref["caustic"] = np.fromfile(out_path + ".caus", dtype=np.float64).reshape(-1, 7)  # ray, state[0..2], tt, bounce, step
    return ref, info


def read_ref_bin(path):
    raw = np.fromfile(path, dtype=np.float64)
    assert raw[0] == 20251018.0, "bad magic"
    n, n_rec, nf, eq = int(raw[1]), int(raw[2]), int(raw[3]), int(raw[4])
    ang = raw[8: 8 + 2 * n].reshape(n, 2)
    recs = raw[8 + 2 * n:].reshape(n, n_rec, nf)
    rec = np.ascontiguousarray(np.transpose(recs[:, :, :26], (2, 0, 1)))         # fields 0..25 of include/geoac_b200.h
    return {"theta_deg": ang[:, 0].copy(), "phi_deg": ang[:, 1].copy(), "rec": rec,
            "status": recs[:, :, REF_F_STATUS].astype(np.int32), "n_steps": recs[:, :, REF_F_NSTEPS].astype(np.int32),
            "eq_cnt": eq, "total_steps": int(raw[5]), "t_rk4_s": raw[6], "t_post_s": raw[7]}


def angles_rad(theta_deg, phi_deg):
    """deg -> rad exactly as the mains do (Code/GeoAc3D_main.cpp:228-229)."""
    Pi = 3.141592653589793238462643
    theta_deg = np.asarray(theta_deg, dtype=np.float64)
    phi_deg = np.asarray(phi_deg, dtype=np.float64)
    return theta_deg * Pi / 180.0, Pi / 2.0 - phi_deg * Pi / 180.0
