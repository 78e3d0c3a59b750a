// eq_cartesian.cuh -- per-ray physics of the two Cartesian stratified variants, de-duplicated for one thread per ray.
//
//   Eq3D : 3-D moving medium            (reference Code/GeoAc/GeoAc.EquationSets.3DStratified.cpp)
//   Eq2D : 2-D effective sound speed    (reference Code/GeoAc/GeoAc.EquationSets.2DStratified.cpp)
//
// Each equation set exposes the same static interface used by the generic kernel in trace_kernel.cuh:
//   init / rhs / step_size / left_region / below_ground / segment / reflect / arrival / altitude
// The reference's formula slips that are observable are reproduced on purpose.
#pragma once
#include "core.cuh"

namespace geoac {

// ======================================================= 3-D stratified =======================================
template <bool AMP>
struct Eq3D {
    using Scout = Eq3D<false>;           // amplitude-free set used by the cost scout (trace_kernel.cuh)
    static constexpr int NEQ = AMP ? 12 : 4;
    using Atmo = Table1D;
    using Cursor = int;
    static constexpr int VARIANT = GEOAC_3D;
    static constexpr bool QUADRATIC_INTERCEPT = true;

    // per-ray constants: horizontal eikonal components and their launch-angle derivatives (3DStratified.cpp:75-89)
    struct RayC { double nx, ny, mxt, myt, mxp, myp, costh, az_deg; };

    GEOAC_HD static double altitude(const double* y) { return y[2]; }

    // GeoAc_SetInitialConditions, 3DStratified.cpp:69-131
    GEOAC_HD static void init(const LaunchConsts& L, const Table1D&, double theta, double phi, RayC& rc, double* y, int&) {
        double st, ct, sp, cp;
        sincos(theta, &st, &ct); sincos(phi, &sp, &cp);
        const double inv_c0 = 1.0 / L.c_src;
        const double Mu = L.u_src * inv_c0, Mv = L.v_src * inv_c0;
        const double n0 = ct * cp, n1 = ct * sp, n2 = st;
        const double t0 = -st * cp, t1 = -st * sp, t2 = ct;
        const double p0 = -ct * sp, p1 = ct * cp;
        const double M = 1.0 + (n0 * Mu + n1 * Mv);
        const double dMt = t0 * Mu + t1 * Mv, dMp = p0 * Mu + p1 * Mv;
        const double iM = 1.0 / M, iM2 = iM * iM;
        rc.nx = n0 * iM; rc.ny = n1 * iM;
        rc.mxt = t0 * iM - n0 * iM2 * dMt;  rc.myt = t1 * iM - n1 * iM2 * dMt;
        rc.mxp = p0 * iM - n0 * iM2 * dMp;  rc.myp = p1 * iM - n1 * iM2 * dMp;
        rc.costh = ct;
        rc.az_deg = (kPi / 2.0 - phi) * 180.0 / kPi;
        y[0] = L.src[0]; y[1] = L.src[1]; y[2] = L.src[2];
        y[3] = n2 * iM;
        if (AMP) {
            y[4] = y[5] = y[6] = 0.0; y[8] = y[9] = y[10] = 0.0;
            y[7]  = t2 * iM - n2 * iM2 * dMt;
            y[11] = -n2 * iM2 * dMp;
        }
    }

    // GeoAc_Set_ds, 3DStratified.cpp:191-198
    GEOAC_HD static double step_size(const LaunchConsts& L, const double* y) { return step_size_z(L, y[2] - L.z_grnd); }

    // GeoAc_UpdateSources + GeoAc_EvalSrcEq, 3DStratified.cpp:203-310: all NEQ right-hand sides from ONE atmosphere sample
    GEOAC_HD static double rhs(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* p, double* f, int& cur) {
        const SegPos sp = seg_locate(T, p[2], cur);
        double Tv, dT, ddT, u, du, ddu, v, dv, ddv;
        if (AMP) {
            spl_f2(T, TAB_T, sp, Tv, dT, ddT);
            spl_f2(T, TAB_U, sp, u, du, ddu);
            spl_f2(T, TAB_V, sp, v, dv, ddv);
        } else {
            spl_f1(T, TAB_T, sp, Tv, dT); ddT = 0.0;
            spl_f1(T, TAB_U, sp, u, du);  ddu = 0.0;
            spl_f1(T, TAB_V, sp, v, dv);  ddv = 0.0;
        }
        const SoundSpeed s = sound_speed2(Tv, dT, ddT);
        const double nz = p[3];
        const double nu_mag = (L.c_src - (rc.nx * u + rc.ny * v)) * s.inv_c;     // c0/c (1 - nu.v/c0), w = 0
        // c_prop = c n/|n| + wind = q / nu_mag with q = c n + nu_mag wind: its direction and 1/|c_prop| = nu_mag/|q| need no
        // reciprocal, so the Newton chains of 1/nu_mag (auxiliary equations only) and 1/|q| run side by side
        const double inv_nm = g_rcp(nu_mag);
        const double q0 = fma(nu_mag, u, s.c * rc.nx), q1 = fma(nu_mag, v, s.c * rc.ny), q2 = s.c * nz;
        const double inv_q = g_rsqrt(q0 * q0 + q1 * q1 + q2 * q2);
        const double inv_cpm = nu_mag * inv_q;
        const double cn = s.c * inv_nm;
        const double cp0 = q0 * inv_nm, cp1 = q1 * inv_nm, cp2 = q2 * inv_nm;
        const double G = nu_mag * s.dc + rc.nx * du + rc.ny * dv;
        if (!AMP) {
            // every right-hand side carries 1/|q|: it is returned as the common factor and folded into the RK4 step factors
            f[0] = q0; f[1] = q1; f[2] = q2;                                         // dx/ds = q / |q|
            f[3] = -G * nu_mag;                                                       // -G / |c_prop|,  1/|c_prop| = nu_mag / |q|
            return inv_q;
        }
        // with the auxiliary set the common factor is 1/|c_prop| (twelve multiplications per stage less)
        f[0] = cp0; f[1] = cp1; f[2] = cp2;                                          // dx/ds = c_prop / |c_prop|
        f[3] = -G;
        {
            const double H = nu_mag * s.ddc + rc.nx * ddu + rc.ny * ddv;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double mx = a ? rc.mxp : rc.mxt, my = a ? rc.myp : rc.myt;
                const double mz = p[7 + 4 * a], Z = p[6 + 4 * a];
                const double dnm = (rc.nx * mx + rc.ny * my + nz * mz) * inv_nm;
                const double q = inv_nm * (s.dc * Z - cn * dnm);
                const double d0 = rc.nx * q + cn * mx + du * Z;
                const double d1 = rc.ny * q + cn * my + dv * Z;
                const double d2 = nz * q + cn * mz;
                const double dcpm = cp0 * d0 + cp1 * d1 + cp2 * d2;                 // times inv_cpm below
                const double g = dcpm * inv_cpm * inv_cpm;                          // d|cp| / |cp|
                f[4 + 4 * a] = d0 - cp0 * g;
                f[5 + 4 * a] = d1 - cp1 * g;
                f[6 + 4 * a] = d2 - cp2 * g;
                f[7 + 4 * a] = G * g - (dnm * s.dc + mx * du + my * dv + H * Z);
            }
        }
        return inv_cpm;
    }

    // BreakCheck / GroundCheck, 3DStratified.cpp:327-343 (strict inequalities on the unclamped state)
    GEOAC_HD static bool left_region(const LaunchConsts& L, const RayC&, const double* y) {
        const double r2 = y[0] * y[0] + y[1] * y[1];
        return (y[2] > L.vert_limit) || (r2 > L.range_limit * L.range_limit);      // sqrt(r2) > limit
    }
    GEOAC_HD static bool below_ground(const LaunchConsts& L, const double* y) { return y[2] < L.z_grnd; }
    GEOAC_HD static double break_margin(const LaunchConsts& L, const RayC&, const double* ya, const double* yb) {
        double m = frac_beyond(ya[2] - L.vert_limit, yb[2] - L.vert_limit, 2.0);
        return frac_beyond(sqrt(ya[0] * ya[0] + ya[1] * ya[1]) - L.range_limit, sqrt(yb[0] * yb[0] + yb[1] * yb[1]) - L.range_limit, m);
    }

    // one segment of GeoAc_TravelTime + GeoAc_SB_Atten, 3DStratified.cpp:348-405, 456-490 (shared midpoint sample)
    GEOAC_HD static void segment(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* ya, const double* yb,
                                 int& cur, double& dtt, double& datt) {
        const double dx = yb[0] - ya[0], dy = yb[1] - ya[1], dz = yb[2] - ya[2];
        const double ds = g_sqrt(fmax(dx * dx + dy * dy + dz * dz, 1e-290));
        const double zm = ya[2] + dz * 0.5;
        const double nz = ya[3] + (yb[3] - ya[3]) * 0.5;
        const SegPos sp = seg_locate(T, zm, cur);
        const double Tv = spl_f(T, TAB_T, sp);
        const double u = spl_f(T, TAB_U, sp);
        const double v = spl_f(T, TAB_V, sp);
        const double gT = kGamR * Tv;
        const double inv_c = g_rsqrt(gT), c = gT * inv_c;
        const double nu_mag = (L.c_000 - rc.nx * u - rc.ny * v) * inv_c;            // c(0,0,0): App. A-4
        const double q0 = fma(nu_mag, u, c * rc.nx), q1 = fma(nu_mag, v, c * rc.ny), q2 = c * nz;    // c_prop = q / nu_mag (see rhs)
        dtt = (ds * nu_mag) * g_rsqrt(q0 * q0 + q1 * q1 + q2 * q2);
        datt = sb_alpha_1d(L, T, sp, cur, zm, zm, c, inv_c) * ds;
    }

    // GeoAc_ApproximateIntercept + GeoAc_SetReflectionConditions, 3DStratified.cpp:136-186
    GEOAC_HD static void reflect(const LaunchConsts& L, const Table1D&, const RayC& rc, const double* ym2, const double* ym1,
                                 const double* yk, double* y0, int&) {
        const double dz_k = yk[2] - ym1[2], dz_g = ym1[2] - L.z_grnd;
        const double a1 = dz_g / dz_k, a2 = 0.5 * a1 * a1;
        double pv[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; i++) pv[i] = ym1[i] + (ym1[i] - yk[i]) * a1 + (yk[i] + ym2[i] - 2.0 * ym1[i]) * a2;
        const double cg = L.c_gnd;
        const double dnuz_ds = -1.0 / cg * (L.c_src / cg * L.dc_gnd + rc.nx * L.du_gnd + rc.ny * L.dv_gnd);
        y0[0] = pv[0]; y0[1] = pv[1]; y0[2] = pv[2];          // restart from the fitted z (App. A-20)
        y0[3] = -pv[3];
        if (AMP) {
            const double den = 1.0 / (cg / L.c_src * pv[3]);
            y0[4] = pv[4]; y0[5] = pv[5]; y0[8] = pv[8]; y0[9] = pv[9];
            y0[6] = -pv[6]; y0[10] = -pv[10];
            y0[7]  = -pv[7]  + 2.0 * dnuz_ds * pv[6] * den;
            y0[11] = -pv[11] + 2.0 * dnuz_ds * pv[10] * den;
        }
    }

    // GeoAc_Jacobian, 3DStratified.cpp:410-428: determinant of (dx/ds, dx/dtheta, dx/dphi); its sign changes mark caustics
    GEOAC_HD static double jacobian(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[2], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double u = spl_f(T, TAB_U, sp), v = spl_f(T, TAB_V, sp);
        const double nu_mag = (L.c_src - rc.nx * u - rc.ny * v) / c;
        const double cp[3] = { c * rc.nx / nu_mag + u, c * rc.ny / nu_mag + v, c * yk[3] / nu_mag };
        const double cpm = sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
        const double xs = cp[0] / cpm, ys = cp[1] / cpm, zs = cp[2] / cpm;
        return xs * (yk[5] * yk[10] - yk[9] * yk[6]) - yk[4] * (ys * yk[10] - zs * yk[9]) + yk[8] * (ys * yk[6] - zs * yk[5]);
    }

    // GeoAc_Amplitude at an arbitrary state (also evaluated along the path for the raypath rows)
    GEOAC_HD static double amplitude(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[2], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double u = spl_f(T, TAB_U, sp), v = spl_f(T, TAB_V, sp);
        const double rho = spl_f(T, TAB_RHO, sp);
        const double c0 = L.c_src, u0 = L.u_src, v0 = L.v_src;
        const double nz = yk[3];
        const double nu_mag = (c0 - rc.nx * u - rc.ny * v) / c;
        const double nu_mag0 = 1.0 - (rc.nx * u0 - rc.ny * v0) / c0;             // sign slip kept (App. A-5)
        const double cp[3] = { c * rc.nx / nu_mag + u, c * rc.ny / nu_mag + v, c * nz / nu_mag };
        const double ax = rc.nx / nu_mag0, ay = rc.ny / nu_mag0;
        const double cq[3] = { c0 * ax + u0, c0 * ay + v0, c0 * sqrt(1.0 - ax * ax - ay * ay) };
        const double cpm = sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
        const double cqm = sqrt(cq[0] * cq[0] + cq[1] * cq[1] + cq[2] * cq[2]);
        const double xs = cp[0] / cpm, ys = cp[1] / cpm, zs = cp[2] / cpm;
        const double D = xs * (yk[5] * yk[10] - yk[9] * yk[6]) - yk[4] * (ys * yk[10] - zs * yk[9]) + yk[8] * (ys * yk[6] - zs * yk[5]);
        const double num = rho * nu_mag * (c * c * c) * cqm * rc.costh;
        const double den = L.rho_src * nu_mag0 * (c0 * c0 * c0) * cpm * D;
        return 1.0 / (4.0 * kPi) * sqrt(fabs(num / den));
    }

    // GeoAc_Jacobian + GeoAc_Amplitude (3DStratified.cpp:410-451) and the results row of GeoAc3D_main.cpp:281-298
    GEOAC_HD static void arrival(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* ym1, const double* yk,
                                 double tt, int& cur, double& amp, double& incl, double& backaz, double& aux, double& margin) {
        (void)tt;
        incl = -asin(L.c_gnd / L.c_src * yk[3]) * 180.0 / kPi;
        double b = rc.az_deg + 180.0;
        while (b > 180.0) b -= 360.0;
        while (b < -180.0) b += 360.0;
        backaz = b; aux = 0.0;
        margin = (yk[2] - L.z_grnd) / fabs(yk[2] - ym1[2]);
        amp = amplitude(L, T, rc, yk, cur);
    }
};

// ======================================================= 2-D effective sound speed ============================
template <bool AMP>
struct Eq2D {
    using Scout = Eq2D<false>;           // amplitude-free set used by the cost scout (trace_kernel.cuh)
    static constexpr int NEQ = AMP ? 6 : 3;
    using Atmo = Table1D;
    using Cursor = int;
    static constexpr int VARIANT = GEOAC_2D;
    static constexpr bool QUADRATIC_INTERCEPT = true;

    struct RayC { double cphi, sphi, costh, sinth, ceff0, inv_ceff0, theta_deg; };

    GEOAC_HD static double altitude(const double* y) { return y[1]; }

    // GeoAc_SetInitialConditions, 2DStratified.cpp:38-68
    GEOAC_HD static void init(const LaunchConsts& L, const Table1D&, double theta, double phi, RayC& rc, double* y, int&) {
        sincos(phi, &rc.sphi, &rc.cphi); sincos(theta, &rc.sinth, &rc.costh);
        rc.ceff0 = L.c_src + L.u_src * rc.cphi + L.v_src * rc.sphi;
        rc.inv_ceff0 = 1.0 / rc.ceff0;
        rc.theta_deg = theta * 180.0 / kPi;
        y[0] = 0.0; y[1] = L.src[2]; y[2] = rc.sinth;
        if (AMP) { y[3] = 0.0; y[4] = 0.0; y[5] = rc.costh; }
    }

    GEOAC_HD static double step_size(const LaunchConsts& L, const double* y) { return step_size_z(L, y[1] - L.z_grnd); }   // 2DStratified.cpp:123-130

    // GeoAc_UpdateSources + GeoAc_EvalSrcEq, 2DStratified.cpp:135-181
    GEOAC_HD static double rhs(const LaunchConsts&, const Table1D& T, const RayC& rc, const double* p, double* f, int& cur) {
        const SegPos sp = seg_locate(T, p[1], cur);
        double Tv, dT, ddT, u, du, ddu, v, dv, ddv;
        if (AMP) {
            spl_f2(T, TAB_T, sp, Tv, dT, ddT);
            spl_f2(T, TAB_U, sp, u, du, ddu);
            spl_f2(T, TAB_V, sp, v, dv, ddv);
        } else {
            spl_f1(T, TAB_T, sp, Tv, dT); ddT = 0.0;
            spl_f1(T, TAB_U, sp, u, du);  ddu = 0.0;
            spl_f1(T, TAB_V, sp, v, dv);  ddv = 0.0;
        }
        const SoundSpeed s = sound_speed2(Tv, dT, ddT);
        const double c  = s.c  + u  * rc.cphi + v  * rc.sphi;
        const double dc = s.dc + du * rc.cphi + dv * rc.sphi;
        const double inv_c0 = rc.inv_ceff0, inv_c = g_rcp(c);
        const double cr = c * inv_c0;                              // c/c0
        f[0] = cr * rc.costh;
        f[1] = cr * p[2];
        f[2] = -rc.ceff0 * inv_c * inv_c * dc;
        if (AMP) {
            const double ddc = s.ddc + ddu * rc.cphi + ddv * rc.sphi;
            const double dzt = p[4];
            const double g = dc * dzt * inv_c0;
            f[3] = g * rc.costh - cr * rc.sinth;
            f[4] = g * p[2] + cr * p[5];
            const double dcc = dc * inv_c;
            f[5] = (2.0 * dcc * dcc - ddc * inv_c) * rc.ceff0 * inv_c * dzt;
        }
        return 1.0;
    }

    GEOAC_HD static bool left_region(const LaunchConsts& L, const RayC&, const double* y) {   // 2DStratified.cpp:194-203
        return (y[1] > L.vert_limit) || (y[0] > L.range_limit);
    }
    GEOAC_HD static bool below_ground(const LaunchConsts& L, const double* y) { return y[1] < L.z_grnd; }
    GEOAC_HD static double break_margin(const LaunchConsts& L, const RayC&, const double* ya, const double* yb) {
        return frac_beyond(ya[0] - L.range_limit, yb[0] - L.range_limit, frac_beyond(ya[1] - L.vert_limit, yb[1] - L.vert_limit, 2.0));
    }

    // one segment of GeoAc_TravelTimeSegment + GeoAc_SB_AttenSegment, 2DStratified.cpp:235-286
    GEOAC_HD static void segment(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* ya, const double* yb,
                                 int& cur, double& dtt, double& datt) {
        const double dr = yb[0] - ya[0], dz = yb[1] - ya[1];
        const double zm = ya[1] + dz * 0.5;
        const double ds = g_sqrt(fmax(dr * dr + dz * dz, 1e-290));
        const SegPos sp = seg_locate(T, zm, cur);
        const double Tv = spl_f(T, TAB_T, sp);
        const double u = spl_f(T, TAB_U, sp);
        const double v = spl_f(T, TAB_V, sp);
        const double gT = kGamR * Tv;
        const double inv_c = g_rsqrt(gT), c = gT * inv_c;
        dtt = ds * g_rcp(c + u * rc.cphi + v * rc.sphi);
        datt = sb_alpha_1d(L, T, sp, cur, zm, zm, c, inv_c) * ds;
    }

    // GeoAc_ApproximateIntercept + GeoAc_SetReflectionConditions, 2DStratified.cpp:74-117
    GEOAC_HD static void reflect(const LaunchConsts& L, const Table1D&, const RayC& rc, const double* ym2, const double* ym1,
                                 const double* yk, double* y0, int&) {
        const double dz_k = yk[1] - ym1[1], dz_g = ym1[1] - L.z_grnd;
        const double a1 = dz_g / dz_k, a2 = 0.5 * a1 * a1;
        double pv[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; i++) pv[i] = ym1[i] + (ym1[i] - yk[i]) * a1 + (yk[i] + ym2[i] - 2.0 * ym1[i]) * a2;
        const double ceff_d = L.dc_gnd + L.du_gnd * rc.cphi + L.dv_gnd * rc.sphi;
        const double dnuz_ds = -rc.ceff0 / (L.c_gnd * L.c_gnd) * ceff_d;
        y0[0] = pv[0]; y0[1] = L.z_grnd; y0[2] = -pv[2];
        if (AMP) {
            y0[3] = pv[3]; y0[4] = -pv[4];
            y0[5] = -pv[5] + 2.0 * dnuz_ds * pv[4] / (L.c_gnd / rc.ceff0 * pv[2]);
        }
    }

    // GeoAc_Jacobian, 2DStratified.cpp:291-300
    GEOAC_HD static double jacobian(const LaunchConsts&, const Table1D& T, const RayC& rc, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[1], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double drds = c / rc.ceff0 * rc.costh, dzds = c / rc.ceff0 * yk[2];
        return yk[0] * (drds * yk[4] - dzds * yk[3]);
    }

    // GeoAc_Amplitude at an arbitrary state (also evaluated along the path for the raypath rows)
    GEOAC_HD static double amplitude(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[1], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double rho = spl_f(T, TAB_RHO, sp);
        const double drds = c / rc.ceff0 * rc.costh, dzds = c / rc.ceff0 * yk[2];
        const double D = yk[0] * (drds * yk[4] - dzds * yk[3]);
        return 1.0 / (4.0 * kPi) * sqrt(fabs((rho * c * rc.costh) / (L.rho_gnd * rc.ceff0 * D)));
    }

    // GeoAc_Jacobian + GeoAc_Amplitude, 2DStratified.cpp:291-312; results row GeoAc2D_main.cpp:216-226
    GEOAC_HD static void arrival(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* ym1, const double* yk,
                                 double tt, int& cur, double& amp, double& incl, double& backaz, double& aux, double& margin) {
        (void)tt;
        incl = -rc.theta_deg; backaz = 0.0; aux = 0.0;
        margin = (yk[1] - L.z_grnd) / fabs(yk[1] - ym1[1]);
        amp = amplitude(L, T, rc, yk, cur);
    }
};

}  // namespace geoac
