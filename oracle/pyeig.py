"""oracle/pyeig.py -- TEST INFRASTRUCTURE (CPU oracle), never imported by the product.

Literal, one-ray-at-a-time restatement of the reference's Cartesian eigenray search on top of the oracle's ray tracer:
  run_eig_search      Code/GeoAc3D_main.cpp:531-541 (GeoAc3D.RngDep_main.cpp likewise)
  estimate_eigenray   Code/GeoAc/GeoAc.Eigenray.cpp:30-121  (GeoAc_EstimateEigenray, Modify_d_theta :23-27)
  eigenray_lm         Code/GeoAc/GeoAc.Eigenray.cpp:123-335 (GeoAc_3DEigenray_LM; `long double` -> numpy.longdouble, x87 here)
  glob=True           Code/GeoAc/GeoAc.Eigenray.Global.cpp (Calc_Bearing :25-30, Calc_GC_Distance :32-37, Modify_d_theta :39-43,
                      GeoAc_EstimateEigenray :46-136, GeoAc_3DEigenray_LM :139-320), Code/GeoAcGlobal_main.cpp:573-583
Pinned against the unmodified reference by tests/golden/eig_*.npz (dumped by oracle/_ref/ref_eig3d[rngdep])."""
import math

import numpy as np

from geoac_b200 import abi
from oracle import pyoracle as po

PI = 3.141592653589793238462643          # Code/GeoAc/GeoAc.Parameters.cpp:28
LD = np.longdouble


class RayEngine:
    """solution[k] after n_bnc reflections for GeoAc_theta, GeoAc_phi [radians], and the BreakCheck flag."""

    def __init__(self, variant, atmo, params):
        self.variant, self.atmo, self.params = variant, atmo, params
        self.rays = 0

    def ray(self, theta, phi, calc_amp, n_bnc, accum_per_segment=0):
        p = self.params.copy()
        p.bounces, p.calc_amp, p.accum_per_segment = n_bnc, int(calc_amp), accum_per_segment
        out = po.trace(self.variant, self.atmo, p, np.array([theta]), np.array([phi]))
        self.rays += 1
        st = out["status"][0]
        brk = bool((st != abi.ST_ARRIVAL).any())         # a step-limit end counts as a break (include/geoac_b200.h)
        return out["rec"][:, 0, n_bnc].copy(), brk


def bearing(lat1, long1, lat2, long2):
    term1 = math.sin((long2 - long1) * PI / 180.0)
    term2 = math.cos(lat1 * PI / 180.0) * math.tan(lat2 * PI / 180.0) - math.sin(lat1 * PI / 180.0) * math.cos((long2 - long1) * PI / 180.0)
    return math.atan2(term1, term2) * 180.0 / PI


def gc_distance(lat1, long1, lat2, long2):
    term1 = math.pow(math.sin((lat2 - lat1) * PI / 180.0 / 2.0), 2)
    term2 = math.cos(lat1 * PI / 180.0) * math.cos(lat2 * PI / 180.0) * math.pow(math.sin((long2 - long1) * PI / 180.0 / 2.0), 2)
    return 2.0 * 6370.0 * math.asin(math.sqrt(term1 + term2))


def modify_d_theta(dr, dr_dtheta, big, small, wf=2.0):
    with np.errstate(all="ignore"):
        width = np.float64(wf) * np.float64(dr_dtheta) ** 2
        return float(big - (big - small) * np.exp(-np.float64(dr) * np.float64(dr) / width))


def estimate_eigenray(eng, src, rcv, theta_min, theta_max, bounces, az_lim, big=0.25, small=0.002, glob=False):
    """-> (ok, theta_estimate, phi_estimate, theta_next)"""
    if glob:
        r_rcvr = gc_distance(src[0], src[1], rcv[0], rcv[1])
        phi = bearing(src[0], src[1], rcv[0], rcv[1])
    else:
        r_rcvr = math.sqrt(math.pow(rcv[0] - src[0], 2) + math.pow(rcv[1] - src[1], 2))
        phi = 180.0 / 3.14159 * math.atan2(rcv[1] - src[1], rcv[0] - src[0])
    iterations = 0
    theta_estimate, phi_estimate, theta_next = theta_max, 0.0, 0.0
    d_theta, d_phi = big, 10.0
    theta_max_reached = False
    while abs(d_phi) > az_lim and iterations < 5:
        r = r_prev = r_rcvr
        theta = theta_min
        phi_rad = (90.0 - phi) * PI / 180.0 if glob else phi * PI / 180.0
        # for(theta = theta_min; theta <= theta_max; theta += d_theta)   (Global: theta < theta_max)
        while (theta < theta_max) if glob else (theta <= theta_max):
            if theta + d_theta >= theta_max:
                theta_max_reached = True
            s, brk = eng.ray(theta * PI / 180.0, phi_rad, False, bounces)
            if brk:
                r = r_prev = r_rcvr
            elif glob:
                r = gc_distance(src[0], src[1], s[1] * 180.0 / PI, s[2] * 180.0 / PI)
            else:
                r = math.sqrt(math.pow(s[0] - src[0], 2) + math.pow(s[1] - src[1], 2))
            if (r - r_rcvr) * (r_prev - r_rcvr) < 0.0:
                if iterations == 0:
                    theta_next = theta
                if glob:
                    d_phi = bearing(src[0], src[1], rcv[0], rcv[1])
                    d_phi -= bearing(src[0], src[1], s[1] * 180.0 / PI, s[2] * 180.0 / PI)
                else:
                    d_phi = (math.atan2(rcv[1] - src[1], rcv[0] - src[0]) - math.atan2(s[1] - src[1], s[0] - src[0])) * 180.0 / PI
                while d_phi > 180.0:
                    d_phi -= 360.0
                while d_phi < -180.0:
                    d_phi += 360.0
                if abs(d_phi) < az_lim:
                    return True, theta - d_theta, (90.0 - phi if glob else phi), theta_next
                phi += d_phi * 0.9
                theta_min = max(theta - 7.5, theta_min)
                break
            if iterations >= 3:
                with np.errstate(all="ignore"):
                    d_theta = modify_d_theta(r - r_rcvr, float(np.float64(r - r_prev) / np.float64(2.0 * d_theta)), big, small, 0.5 if glob else 2.0)
            r_prev = r
            theta += d_theta
        if theta_max_reached:
            theta_next = theta_max
            break
        iterations += 1
        if 1 <= iterations < 3:
            d_theta = big / 2.0
    return False, theta_estimate, phi_estimate, theta_next


def eigenray_lm(eng, src, rcv, theta, phi, bnc_cnt, iterate_limit, strat, mach=(0.0, 0.0), tolerance=0.1, glob=False, z_grnd=0.0):
    """-> (found, theta, phi, iterations used)"""
    dr_prev = 10000.0
    lim = 0.2
    step_scalar = 1.0
    dt = dp = LD(0)
    n = 0
    for n in range(iterate_limit + 1):
        if n == iterate_limit:
            break
        th, ph = theta * PI / 180.0, phi * PI / 180.0
        if strat and not glob:
            nu0 = (math.cos(th) * math.cos(ph), math.cos(th) * math.sin(ph), math.sin(th))
            M = 1.0 + (nu0[0] * mach[0] + nu0[1] * mach[1] + nu0[2] * 0.0)
            nxy = (nu0[0] / M, nu0[1] / M)
        s, brk = eng.ray(th, ph, True, bnc_cnt)
        if brk:
            break
        if glob:
            x, y = LD(s[1]), LD(s[2])                       # lat, lon
            dr = gc_distance(float(x * LD(180.0) / LD(PI)), float(y * LD(180.0) / LD(PI)), rcv[0], rcv[1])
        else:
            x, y = LD(s[0]), LD(s[1])
            dx, dy = LD(rcv[0]) - x, LD(rcv[1]) - y
            dr = float(np.sqrt(dx * dx + dy * dy))
        if dr < tolerance:
            return True, theta, phi, n
        elif n > 0 and dr > dr_prev:
            theta = float(LD(theta) - dt * LD(step_scalar))
            phi = float(LD(phi) - dp * LD(step_scalar))
            step_scalar /= 2.0
            if np.sqrt(dt * dt + dp * dp) * LD(step_scalar) < LD(1.0e-12):
                break
        else:
            step_scalar = min(1.0, step_scalar * 1.25)
            if glob:
                rg = 6370.0 + z_grnd
                dx, dy = LD(rcv[0] * PI / 180.0) - x, LD(rcv[1] * PI / 180.0) - y
                dx_dt = LD(s[7] - 1.0 / rg * s[4] / s[3] * s[6]); dx_dp = LD(s[13] - 1.0 / rg * s[4] / s[3] * s[12])
                cl = np.cos(x)
                dy_dt = LD(s[8]) - LD(1.0) / (LD(rg) * cl) * LD(s[5]) / LD(s[3]) * LD(s[6])
                dy_dp = LD(s[14]) - LD(1.0) / (LD(rg) * cl) * LD(s[5]) / LD(s[3]) * LD(s[12])
            elif strat:
                dx_dt = LD(s[4] - nxy[0] / s[3] * s[6]); dy_dt = LD(s[5] - nxy[1] / s[3] * s[6])
                dx_dp = LD(s[8] - nxy[0] / s[3] * s[10]); dy_dp = LD(s[9] - nxy[1] / s[3] * s[10])
            else:
                dx_dt = LD(s[6] - s[3] / s[5] * s[8]); dy_dt = LD(s[7] - s[4] / s[5] * s[8])
                dx_dp = LD(s[12] - s[3] / s[5] * s[14]); dy_dp = LD(s[13] - s[4] / s[5] * s[14])
            det = dx_dt * dy_dp - dx_dp * dy_dt
            if glob:
                dt = (dy_dp * dx - dx_dp * dy) / det * LD(180.0) / LD(PI)
                dp = (-dy_dt * dx + dx_dt * dy) / det * LD(180.0) / LD(PI)
            else:
                dt = LD(1.0) / det * (dy_dp * dx - dx_dp * dy) * LD(180.0) / LD(PI)
                dp = LD(1.0) / det * (dx_dt * dy - dy_dt * dx) * LD(180.0) / LD(PI)
            dt = min(max(dt, LD(-lim)), LD(lim))
            dp = min(max(dp, LD(-lim)), LD(lim))
            theta = float(LD(theta) + dt * LD(step_scalar))
            phi = float(LD(phi) + dp * LD(step_scalar))
            dr_prev = dr
    return False, theta, phi, n


def run_eig_search(variant, atmo, params, rcv, theta_min=0.5, theta_max=45.0, bnc_min=0, bnc_max=0, iterations=25,
                   azimuth_err_lim=2.0, src_lat_deg=30.0, src_lon_deg=0.0):
    """Rows as geoac_eigenray_search produces them for ONE receiver (fields 0-8, 16; 9-15 for found eigenrays), + rays traced."""
    strat = variant == abi.GEOAC_3D
    glob = variant in (abi.GEOAC_GLOBAL, abi.GEOAC_GLOBAL_RNGDEP)
    if glob:
        src = (src_lat_deg, src_lon_deg, max(params.z_grnd, params.src[0]))
        params = params.copy()
        params.src[0], params.src[1], params.src[2] = src[2], src[0] * PI / 180.0, src[1] * PI / 180.0
    else:
        src = (params.src[0], params.src[1], max(params.z_grnd, params.src[2]))
    eng = RayEngine(variant, atmo, params)
    mach = (0.0, 0.0)
    if strat:
        a = po.atmo_sample(atmo, *src)
        mach = (a[1] / a[0], a[2] / a[0])
    rows = []
    for n_bnc in range(bnc_min, bnc_max + 1):
        theta_start = theta_min
        while theta_start < theta_max:
            ok, th_est, ph_est, th_next = estimate_eigenray(eng, src, rcv, theta_start, theta_max, n_bnc, azimuth_err_lim, glob=glob)
            row = np.zeros(abi.EIG_NF)
            row[1], row[2], row[3], row[4], row[5] = n_bnc, float(ok), th_est, (ph_est if ok else 0.0), th_next
            if ok:
                found, th, ph, it = eigenray_lm(eng, src, rcv, th_est, ph_est, n_bnc, iterations, strat, mach, glob=glob, z_grnd=params.z_grnd)
                row[6], row[7], row[8], row[16] = float(found), th, ph, it
                if found:
                    s, _ = eng.ray(th * PI / 180.0, ph * PI / 180.0, True, n_bnc, accum_per_segment=1)
                    tt = s[abi.F_TRAVELTIME]
                    row[9] = tt
                    row[11], row[12], row[13] = s[abi.F_AMPLITUDE], s[abi.F_ATTEN], s[abi.F_INCLINATION]
                    if glob:
                        row[10] = gc_distance(src[0], src[1], rcv[0], rcv[1]) / tt
                        if variant == abi.GEOAC_GLOBAL_RNGDEP:
                            row[13] = -row[13]
                        back_az = 90.0 - math.atan2(-s[4], -s[5]) * 180.0 / PI
                        dev = back_az - bearing(rcv[0], rcv[1], src[0], src[1])
                        if dev > 180.0:
                            dev -= 360.0
                        if dev < -180.0:
                            dev += 360.0
                    else:
                        row[10] = math.sqrt(math.pow(s[0] - src[0], 2) + math.pow(s[1] - src[1], 2)) / tt
                        back_az = (90.0 - (ph * PI / 180.0) * 180.0 / PI) + 180.0 if strat else 90.0 - math.atan2(-s[4], -s[3]) * 180.0 / PI
                        dev = back_az - (90.0 - math.atan2(src[1] - rcv[1], src[0] - rcv[0]) * 180.0 / PI)
                        while back_az > 180.0:
                            back_az -= 360.0
                        while back_az < -180.0:
                            back_az += 360.0
                        while dev > 180.0:
                            dev -= 360.0
                        while dev < -180.0:
                            dev += 360.0
                    row[14], row[15], row[17] = back_az, dev, abi.ST_ARRIVAL
            rows.append(row)
            theta_start = th_next
    return np.array(rows).reshape(-1, abi.EIG_NF), eng.rays
