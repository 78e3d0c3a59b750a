// eq_global.cuh -- per-ray physics of the spherical stratified variant, de-duplicated for one thread per ray.
//
//   EqGlobal : spherical-earth moving medium over a 1-D G2S profile
//              (reference Code/GeoAc/GeoAc.EquationSets.Global.cpp, atmosphere Code/Atmo/G2S_GlobalSpline1D.cpp)
//
// State y = [r, lat, lon, nu_r, nu_lat, nu_lon, R_theta(3), mu_theta(3), R_phi(3), mu_phi(3)], wind order (w, v, u), w = 0.
// The reference obtains every atmosphere derivative through scalar wrappers (~100 spline look-ups per stage, of which
// three are distinct); here one table sample per stage feeds all 18 right-hand sides, the latitude trigonometry is
// one sincos per stage, and every division is a multiply by one of five reciprocals (1/|nu|, 1/|c_g|, 1/r, 1/cos lat, 1/c).
// The reference's observable formula slips are reproduced on purpose.
#pragma once
#include "core.cuh"

namespace geoac {

template <bool AMP>
struct EqGlobal {
    using Scout = EqGlobal<false>;           // amplitude-free set used by the cost scout (trace_kernel.cuh)
    static constexpr int NEQ = AMP ? 18 : 6;
    using Atmo = Table1D;
    using Cursor = int;
    static constexpr int VARIANT = GEOAC_GLOBAL;

    struct RayC {
        double sth, cth, sph, cph;      // launch inclination / azimuth (math convention)
        double nu0;                      // 1/(1 + n.v0/c0)           (Global.cpp:103)
        double cos_lat_src;              // great-circle range check   (Global.cpp:504)
        double hav_limit;                // sin^2(range_limit / (2 r_earth)): cheap form of the range check
    };

    GEOAC_HD static double altitude(const double* y) { return y[0] - kREarth; }

    // GeoAc_SetInitialConditions, Global.cpp:76-136
    GEOAC_HD static void init(const LaunchConsts& L, const Table1D&, double theta, double phi, RayC& rc, double* y, int&) {
        sincos(theta, &rc.sth, &rc.cth); sincos(phi, &rc.sph, &rc.cph);
        const double inv_c0 = 1.0 / L.c_src;
        const double Mv = L.v_src * inv_c0, Mu = L.u_src * inv_c0;
        const double n0 = rc.sth, n1 = rc.cth * rc.sph, n2 = rc.cth * rc.cph;
        const double t0 = rc.cth, t1 = -rc.sth * rc.sph, t2 = -rc.sth * rc.cph;
        const double p1 = rc.cth * rc.cph, p2 = -rc.cth * rc.sph;
        const double MS = 1.0 + (n1 * Mv + n2 * Mu);
        const double iM = 1.0 / MS, iM2 = iM * iM;
        rc.nu0 = iM;
        rc.cos_lat_src = cos(L.src[1]);
        const double hl = sin(L.range_limit / (2.0 * kREarth));
        rc.hav_limit = (L.range_limit < kPi * kREarth) ? hl * hl : 2.0;     // beyond half the circumference: never
        y[0] = L.src[0] + kREarth; y[1] = L.src[1]; y[2] = L.src[2];
        y[3] = n0 * iM; y[4] = n1 * iM; y[5] = n2 * iM;
        if (AMP) {
            const double dMt = t1 * Mv + t2 * Mu, dMp = p1 * Mv + p2 * Mu;
            y[6] = y[7] = y[8] = 0.0; y[12] = y[13] = y[14] = 0.0;
            y[9]  = t0 * iM - n0 * iM2 * dMt; y[10] = t1 * iM - n1 * iM2 * dMt; y[11] = t2 * iM - n2 * iM2 * dMt;
            y[15] =         - n0 * iM2 * dMp; y[16] = p1 * iM - n1 * iM2 * dMp; y[17] = p2 * iM - n2 * iM2 * dMp;
        }
    }

    // GeoAc_Set_ds, Global.cpp:210-217
    GEOAC_HD static double step_size(const LaunchConsts& L, const double* y) { return step_size_z(L, y[0] - L.ground); }

    // GeoAc_UpdateSources + GeoAc_EvalSrcEq, Global.cpp:222-442
    GEOAC_HD static double rhs(const LaunchConsts&, const Table1D& T, const RayC&, const double* p, double* f, int& cur) {
        const double r = p[0];
        const SegPos sp = seg_locate(T, r, cur);
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
        const double nu0 = p[3], nu1 = p[4], nu2 = p[5];
        const double nm2 = nu0 * nu0 + nu1 * nu1 + nu2 * nu2;
        const double inv_nm = g_rsqrt(nm2), nm = nm2 * inv_nm;
        const double cn = s.c * inv_nm;
        const double g0 = cn * nu0, g1 = cn * nu1 + v, g2 = cn * nu2 + u;          // group velocity (w = 0)
        const double inv_cgm = g_rsqrt(g0 * g0 + g1 * g1 + g2 * g2);
        double st, ct; g_sincos(p[1], &st, &ct);
        const double inv_r = g_rcp(r), inv_ct = g_rcp(ct), tant = st * inv_ct;
        const double GC1 = inv_r, GC2 = inv_r * inv_ct;
        const double nug = nu1 * g1 + nu2 * g2;
        const double A = nu0 * ct + nu1 * st;
        const double B = nu1 * u - nu2 * v;
        const double GT0 = inv_r * nug;
        const double GT1 = nu0 * v - nu0 * g1 + nu2 * g2 * tant;
        const double GT2 = nu0 * u * ct + B * st - g2 * A;
        // every right-hand side carries 1/|c_g|: returned as the common factor and folded into the RK4 step factors
        f[0] = g0; f[1] = GC1 * g1; f[2] = GC2 * g2;
        const double E0 = nm * s.dc + nu1 * dv + nu2 * du;
        f[3] = -(E0 + GT0);
        f[4] = -GC1 * GT1;
        f[5] = -GC2 * GT2;
        if (AMP) {
            const double inv_r2 = inv_r * inv_r, inv_ct2 = inv_ct * inv_ct;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const double R0 = p[6 + 6 * a], R1 = p[7 + 6 * a];
                const double m0 = p[9 + 6 * a], m1 = p[10 + 6 * a], m2 = p[11 + 6 * a];
                const double dnm = (nu0 * m0 + nu1 * m1 + nu2 * m2) * inv_nm;
                const double dc_a = R0 * s.dc, dv_a = R0 * dv, du_a = R0 * du;
                const double q = inv_nm * (dc_a - cn * dnm);
                const double d0 = nu0 * q + cn * m0;
                const double d1 = nu1 * q + cn * m1 + dv_a;
                const double d2 = nu2 * q + cn * m2 + du_a;
                const double gg = (g0 * d0 + g1 * d1 + g2 * d2) * inv_cgm * inv_cgm;      // d|c_g| / |c_g|
                const double dGC1 = -R0 * inv_r2;
                const double dGC2 = dGC1 * inv_ct + st * inv_r * inv_ct2 * R1;
                const double dGT0 = dGC1 * nug + inv_r * (m1 * g1 + nu1 * d1 + m2 * g2 + nu2 * d2);
                const double dGT1 = m0 * v + nu0 * dv_a - m0 * g1 - nu0 * d1 + (m2 * g2 + nu2 * d2) * tant + nu2 * g2 * R1 * inv_ct2;
                const double dGT2 = (m0 * u + nu0 * du_a) * ct - nu0 * u * R1 * st + (m1 * u + nu1 * du_a - m2 * v - nu2 * dv_a) * st + B * R1 * ct
                                  - d2 * A - g2 * (m0 * ct - nu0 * R1 * st + m1 * st + nu1 * R1 * ct);
                f[6 + 6 * a] = d0 - g0 * gg;
                f[7 + 6 * a] = dGC1 * g1 + GC1 * (d1 - g1 * gg);
                f[8 + 6 * a] = dGC2 * g2 + GC2 * (d2 - g2 * gg);
                f[9 + 6 * a]  = (gg * E0 - (dnm * s.dc + nm * (R0 * s.ddc) + m1 * dv + m2 * du + nu1 * (R0 * ddv) + nu2 * (R0 * ddu) + dGT0));
                f[10 + 6 * a] = -(dGC1 * GT1 + GC1 * dGT1);
                f[11 + 6 * a] = -(dGC2 * GT2 + GC2 * dGT2);
            }
        }
        return inv_cgm;
    }

    // GeoAc_BreakCheck, Global.cpp:500-514: altitude limit and great-circle range from the source.
    // range = 2 r_e asin(sqrt(h)) > limit  <=>  h > sin^2(limit / 2 r_e); the cheap form decides unless h is within
    // 1e-9 of the threshold, where the reference's own expression is evaluated so that the decision is the reference's.
    GEOAC_HD static bool left_region(const LaunchConsts& L, const RayC& rc, const double* y) {
        if (y[0] > L.vert_limit) return true;
        // h = sin^2(dlat/2) + cos cos sin^2(dlon/2) <= (dlat^2 + dlon^2)/4: while that bound is below the limit the ray cannot
        // have left, and the trigonometry is skipped (most of a ray's life); beyond it the exact test decides
        const double dlat = y[1] - L.src[1], dlon = y[2] - L.src[2];
        if (0.25 * (dlat * dlat + dlon * dlon) <= rc.hav_limit) return false;
        const double s1 = sin(dlat * 0.5), s2 = sin(dlon * 0.5);
        const double h = s1 * s1 + rc.cos_lat_src * cos(y[1]) * (s2 * s2);
        bool far = h > rc.hav_limit;
        if (fabs(h - rc.hav_limit) <= 1e-9 * rc.hav_limit) far = 2.0 * kREarth * asin(sqrt(h)) > L.range_limit;
        return far;
    }
    GEOAC_HD static bool below_ground(const LaunchConsts& L, const double* y) { return y[0] < L.ground; }
    GEOAC_HD static double break_margin(const LaunchConsts& L, const RayC& rc, const double* ya, const double* yb) {
        auto range = [&](const double* y) {
            const double s1 = sin((y[1] - L.src[1]) * 0.5), s2 = sin((y[2] - L.src[2]) * 0.5);
            return 2.0 * kREarth * asin(sqrt(s1 * s1 + rc.cos_lat_src * cos(y[1]) * (s2 * s2)));
        };
        return frac_beyond(range(ya) - L.range_limit, range(yb) - L.range_limit, frac_beyond(ya[0] - L.vert_limit, yb[0] - L.vert_limit, 2.0));
    }

    // one segment of GeoAc_TravelTime + GeoAc_SB_Atten, Global.cpp:527-589, 634-670 (one shared midpoint sample;
    // the absorption path length uses sin(lat) where the travel time uses cos(lat): App. A-6)
    GEOAC_HD static void segment(const LaunchConsts& L, const Table1D& T, const RayC&, const double* ya, const double* yb,
                                 int& cur, double& dtt, double& datt) {
        const double dr = yb[0] - ya[0], dt = yb[1] - ya[1], dp = yb[2] - ya[2];
        const double rm = ya[0] + dr * 0.5, tm = ya[1] + dt * 0.5;
        double st, ct; g_sincos(tm, &st, &ct);
        const double a = rm * dt, bc = rm * ct * dp, bs = rm * st * dp;
        const double ds_tt = g_sqrt(fmax(dr * dr + a * a + bc * bc, 1e-290));
        const double ds_sb = g_sqrt(fmax(dr * dr + a * a + bs * bs, 1e-290));
        const double n0 = ya[3] + (yb[3] - ya[3]) * 0.5, n1 = ya[4] + (yb[4] - ya[4]) * 0.5, n2 = ya[5] + (yb[5] - ya[5]) * 0.5;
        const SegPos sp = seg_locate(T, rm, cur);
        const double Tv = spl_f(T, TAB_T, sp);
        const double u = spl_f(T, TAB_U, sp);
        const double v = spl_f(T, TAB_V, sp);
        const double gT = kGamR * Tv;
        const double inv_c = g_rsqrt(gT), c = gT * inv_c;
        const double cn = c * g_rsqrt(n0 * n0 + n1 * n1 + n2 * n2);
        const double c0 = cn * n0, c1 = cn * n1 + v, c2 = cn * n2 + u;
        dtt = ds_tt * g_rsqrt(c0 * c0 + c1 * c1 + c2 * c2);
        datt = sb_alpha_1d(L, T, sp, cur, rm, rm - kREarth, c, inv_c) * ds_sb;
    }

    // GeoAc_ApproximateIntercept (first order only, App. A-7) + GeoAc_SetReflectionConditions, Global.cpp:140-205
    GEOAC_HD static void reflect(const LaunchConsts& L, const Table1D& T, const RayC&, const double*, const double* ym1,
                                 const double* yk, double* y0, int& cur) {
        const double a1 = (ym1[0] - L.ground) / (yk[0] - ym1[0]);
        double pv[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; i++) pv[i] = ym1[i] + (ym1[i] - yk[i]) * a1;
        const SegPos sp = seg_locate(T, pv[0], cur);
        double Tv, dT, u, du, v, dv;
        spl_f1(T, TAB_T, sp, Tv, dT);
        spl_f1(T, TAB_U, sp, u, du);
        spl_f1(T, TAB_V, sp, v, dv);
        const double c = sqrt(kGamR * Tv), dc = kGamR / (2.0 * c) * dT;
        const double dnu_r_ds = -1.0 / c * (L.c_src / c * dc + pv[4] * dv + pv[5] * du + c / pv[0] * (pv[4] * pv[4] + pv[5] * pv[5]));
#pragma unroll
        for (int i = 0; i < NEQ; i++) y0[i] = pv[i];
        y0[0] = L.ground;
        y0[3] = -pv[3];
        if (AMP) {
            const double den = 1.0 / (c / L.c_src * pv[3]);
            y0[6] = -pv[6]; y0[12] = -pv[12];
            y0[9]  = -pv[9]  + 2.0 * dnu_r_ds * pv[6] * den;
            y0[15] = -pv[15] + 2.0 * dnu_r_ds * pv[12] * den;
        }
    }

    // GeoAc_Jacobian, Global.cpp:594-608 (dp/ds carries 1/(r sin lat), the volume factor r^2 cos lat: App. A-6)
    GEOAC_HD static double jacobian(const LaunchConsts&, const Table1D& T, const RayC&, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[0], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double u = spl_f(T, TAB_U, sp), v = spl_f(T, TAB_V, sp);
        const double r = yk[0], nu0 = yk[3], nu1 = yk[4], nu2 = yk[5];
        double sl, cl; sincos(yk[1], &sl, &cl);
        const double nm = sqrt(nu0 * nu0 + nu1 * nu1 + nu2 * nu2);
        const double q0 = c * nu0 / nm, q1 = c * nu1 / nm + v, q2 = c * nu2 / nm + u;
        const double qm = sqrt(q0 * q0 + q1 * q1 + q2 * q2);
        const double dr_ds = q0 / qm, dt_ds = 1.0 / r * q1 / qm, dp_ds = 1.0 / (r * sl) * q2 / qm;
        return r * r * cl * (dr_ds * (yk[7] * yk[14] - yk[13] * yk[8]) - yk[6] * (dt_ds * yk[14] - dp_ds * yk[13])
                             + yk[12] * (dt_ds * yk[8] - dp_ds * yk[7]));
    }

    // GeoAc_Amplitude at an arbitrary state (Global.cpp:594-629; also evaluated along the path for the raypath rows)
    GEOAC_HD static double amplitude(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* yk, int& cur) {
        if (!AMP) return 0.0;
        const SegPos sp = seg_locate(T, yk[0], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        const double u = spl_f(T, TAB_U, sp), v = spl_f(T, TAB_V, sp);
        const double rho = spl_f(T, TAB_RHO, sp);
        const double r = yk[0], th = yk[1];
        const double nu0 = yk[3], nu1 = yk[4], nu2 = yk[5];
        double sl, cl; sincos(th, &sl, &cl);
        // Jacobian: c_prop with |nu| = sqrt(nu.nu); dp/ds carries 1/(r sin lat) while the volume factor is
        // r^2 cos lat (App. A-6)
        {
            const double nm = sqrt(nu0 * nu0 + nu1 * nu1 + nu2 * nu2);
            const double q0 = c * nu0 / nm, q1 = c * nu1 / nm + v, q2 = c * nu2 / nm + u;
            const double qm = sqrt(q0 * q0 + q1 * q1 + q2 * q2);
            const double dr_ds = q0 / qm, dt_ds = 1.0 / r * q1 / qm, dp_ds = 1.0 / (r * sl) * q2 / qm;
            const double D = r * r * cl * (dr_ds * (yk[7] * yk[14] - yk[13] * yk[8]) - yk[6] * (dt_ds * yk[14] - dp_ds * yk[13])
                                           + yk[12] * (dt_ds * yk[8] - dp_ds * yk[7]));
            // amplitude: eikonal |nu| = (c0 - nu.v)/c; source c_prop0[1..2] divided by nu_mag, not nu_mag0 (App. A-6)
            const double c0 = L.c_src;
            const double n0v[3] = { rc.sth, rc.cth * rc.sph, rc.cth * rc.cph };
            const double nu_mag = (c0 - nu1 * v - nu2 * u) / c;
            const double nu_mag0 = rc.nu0;
            const double cp0 = c * nu0 / nu_mag, cp1 = c * nu1 / nu_mag + v, cp2 = c * nu2 / nu_mag + u;
            const double cs0 = c0 * n0v[0] / nu_mag0, cs1 = c0 * n0v[1] / nu_mag + L.v_src, cs2 = c0 * n0v[2] / nu_mag + L.u_src;
            const double cpm = sqrt(cp0 * cp0 + cp1 * cp1 + cp2 * cp2);
            const double csm = sqrt(cs0 * cs0 + cs1 * cs1 + cs2 * cs2);
            const double num = rho * nu_mag * (c * c * c) * csm * rc.cth;
            const double den = L.rho_src * nu_mag0 * (c0 * c0 * c0) * cpm * D;
            return 1.0 / (4.0 * kPi) * sqrt(fabs(num / den));
        }
    }

    // GeoAc_Jacobian + GeoAc_Amplitude (Global.cpp:594-629) and the results row of GeoAcGlobal_main.cpp:294-317
    GEOAC_HD static void arrival(const LaunchConsts& L, const Table1D& T, const RayC& rc, const double* ym1, const double* yk,
                                 double tt, int& cur, double& amp, double& incl, double& backaz, double& aux, double& margin) {
        const SegPos sp = seg_locate(T, yk[0], cur);
        const double c = sound_speed0(spl_f(T, TAB_T, sp));
        incl = -asin(c / L.c_src * yk[3]) * 180.0 / kPi;
        double b = 90.0 - atan2(-yk[4], -yk[5]) * 180.0 / kPi;
        if (b < -180.0) b += 360.0;
        if (b > 180.0) b -= 360.0;
        backaz = b;
        const double s1 = sin((yk[1] - L.src[1]) / 2.0), s2 = sin((yk[2] - L.src[2]) / 2.0);
        const double h = s1 * s1 + rc.cos_lat_src * cos(yk[1]) * (s2 * s2);
        aux = 2.0 * kREarth * asin(sqrt(h)) / tt;                                      // celerity
        margin = (yk[0] - L.ground) / fabs(yk[0] - ym1[0]);
        amp = amplitude(L, T, rc, yk, cur);
    }
};

}  // namespace geoac
