/* oracle/orc_eqglobal.c -- TEST INFRASTRUCTURE (CPU oracle).
 * Spherical-earth equation sets: restates Code/GeoAc/GeoAc.EquationSets.Global.cpp (stratified; every atmosphere
 * derivative through the scalar wrappers) and Code/GeoAc/GeoAc.EquationSets.GlobalRngDep.cpp (range dependent; the RK4
 * stages go through Eval_Spline_AllOrder1/2) with identical expression trees.
 * State y = [r, lat, lon, nu_r, nu_lat, nu_lon, R_lt(3), mu_lt(3), R_lp(3), mu_lp(3)], wind order (w, v, u).
 */
#include <math.h>
#include "orc_eqsets.h"

typedef struct srcg {                       /* GeoAc_Sources, Global.cpp:24-58 */
    double src_loc[3], c0;
    double c, dc[5], ddc[3][2];
    double w, dw[5], ddw[3][2];
    double v, dv[5], ddv[3][2];
    double u, du[5], ddu[3][2];
    double nu0, nu_mag, dnu_mag[2];
    double c_gr[3], c_gr_mag, dc_gr[3][2], dc_gr_mag[2];
    double GeoCoeff[3], d_GeoCoeff[3][2];
    double GeoTerms[3], d_GeoTerms[3][2];
} srcg;

#define SRC(r) ((srcg*)(r)->S)
#define ATM(r) ((r)->atmo)

/* range-dependent fast paths (orc_mspline.c): f, 3 first derivatives [, 6 second derivatives] in one call */
void orc_mspline_allorder1(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3]);
void orc_mspline_allorder2(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3], double dd[3][3]);
void orc_mspline_sync_accel(orc_atmo* a);      /* Windu/Windv cursors := Temp's, GlobalRngDep.cpp:236-239, 284-287 */

/* GeoAc_SetInitialConditions, Global.cpp:76-136 */
static void initg(orc_ray* r, double* y) {
    srcg* s = SRC(r); orc_atmo* a = ATM(r);
    double r0 = r->prm->src[0], t0 = r->prm->src[1], p0 = r->prm->src[2];
    double th = r->theta, ph = r->phi;
    double re = a->r_earth;
    s->src_loc[0] = r0 + re; s->src_loc[1] = t0; s->src_loc[2] = p0;
    s->c0 = a->c(a, r0 + re, t0, p0);
    double Mc[3] = { 0.0 / s->c0, a->v(a, r0 + re, t0, p0) / s->c0, a->u(a, r0 + re, t0, p0) / s->c0 };
    double nu0[3] = { sin(th),  cos(th) * sin(ph),  cos(th) * cos(ph) };
    double mlt[3] = { cos(th), -sin(th) * sin(ph), -sin(th) * cos(ph) };
    double mlp[3] = { 0.0,      cos(th) * cos(ph), -cos(th) * sin(ph) };
    double MS = 1.0 + (nu0[0] * Mc[0] + nu0[1] * Mc[1] + nu0[2] * Mc[2]);
    s->nu0 = 1.0 / MS;
    y[0] = r0 + re; y[1] = t0; y[2] = p0;
    for (int i = 0; i < 3; i++) y[3 + i] = nu0[i] / MS;
    if (r->eq_cnt > 6) {
        for (int i = 0; i < 3; i++) { y[6 + i] = 0.0; y[12 + i] = 0.0; }
        for (int i = 0; i < 3; i++) {
            y[9 + i]  = mlt[i] / MS - nu0[i] / pow(MS, 2.0) * (mlt[0] * Mc[0] + mlt[1] * Mc[1] + mlt[2] * Mc[2]);
            y[15 + i] = mlp[i] / MS - nu0[i] / pow(MS, 2.0) * (mlp[0] * Mc[0] + mlp[1] * Mc[1] + mlp[2] * Mc[2]);
        }
    }
}

/* GeoAc_ApproximateIntercept (first order only: the stray ';' of Global.cpp:146-147, App. A-7) +
 * GeoAc_SetReflectionConditions, Global.cpp:140-205 */
static void reflectg(orc_ray* r, const double* ym2, const double* ym1, const double* yk, double* y0) {
    srcg* s = SRC(r); orc_atmo* a = ATM(r); (void)ym2;
    double prev[ORC_MAXEQ];
    double dr_k = yk[0] - ym1[0];
    double dr_grnd = ym1[0] - (a->r_earth + a->z_grnd);
    for (int i = 0; i < r->eq_cnt; i++) prev[i] = ym1[i] + (ym1[i] - yk[i]) / dr_k * dr_grnd;
    double c_ref = a->c(a, prev[0], prev[1], prev[2]);
    double dnu_r_ds = -1.0 / c_ref * (s->c0 / c_ref * a->c_diff(a, prev[0], prev[1], prev[2], 0)
                                      + prev[3] * 0.0
                                      + prev[4] * a->v_diff(a, prev[0], prev[1], prev[2], 0)
                                      + prev[5] * a->u_diff(a, prev[0], prev[1], prev[2], 0)
                                      + c_ref / prev[0] * (pow(prev[4], 2) + pow(prev[5], 2)));
    for (int i = 0; i < r->eq_cnt; i++) y0[i] = prev[i];
    y0[0] = a->r_earth + a->z_grnd;
    y0[3] = -prev[3];
    if (r->eq_cnt > 6) {
        y0[6] = -prev[6]; y0[12] = -prev[12];
        y0[9]  = -prev[9]  + 2.0 * dnu_r_ds * prev[6]  / (c_ref / s->c0 * prev[3]);
        y0[15] = -prev[15] + 2.0 * dnu_r_ds * prev[12] / (c_ref / s->c0 * prev[3]);
    }
}

/* GeoAc_Set_ds, Global.cpp:210-217 */
static double setdsg(orc_ray* r, const double* y) {
    orc_atmo* a = ATM(r);
    double res = 0.05 - 0.049 * exp(-(y[0] - (a->r_earth + a->z_grnd)) / 0.75);
    res = fmin(res, r->prm->ds_max);
    res = fmax(res, r->prm->ds_min);
    return res;
}

/* second half of GeoAc_UpdateSources (Global.cpp:248-269 + 324-367; GlobalRngDep.cpp:253-269, 327-385): everything
 * downstream of the atmosphere sample.  `rd` selects the range-dependent spelling of d_GeoCoeff[2] (operand order). */
static void geometry_terms(orc_ray* r, const double* y, int rd) {
    srcg* s = SRC(r);
    double rr = y[0], theta = y[1];
    double nu[3] = { y[3], y[4], y[5] };
    s->nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    s->c_gr[0] = s->c * nu[0] / s->nu_mag + s->w;
    s->c_gr[1] = s->c * nu[1] / s->nu_mag + s->v;
    s->c_gr[2] = s->c * nu[2] / s->nu_mag + s->u;
    s->c_gr_mag = sqrt(pow(s->c_gr[0], 2) + pow(s->c_gr[1], 2) + pow(s->c_gr[2], 2));
    s->GeoCoeff[0] = 1.0; s->GeoCoeff[1] = 1.0 / rr; s->GeoCoeff[2] = 1.0 / (rr * cos(theta));
    s->GeoTerms[0] = 0.0;
    s->GeoTerms[1] = (nu[0] * s->v - nu[1] * s->w);
    s->GeoTerms[2] = (nu[0] * s->u - nu[2] * s->w) * cos(theta) + (nu[1] * s->u - nu[2] * s->v) * sin(theta);
    s->GeoTerms[0] += 1.0 / rr * (nu[1] * s->c_gr[1] + nu[2] * s->c_gr[2]);
    s->GeoTerms[1] += -nu[0] * s->c_gr[1] + nu[2] * s->c_gr[2] * tan(theta);
    s->GeoTerms[2] += -s->c_gr[2] * (nu[0] * cos(theta) + nu[1] * sin(theta));
    if (!r->calc_amp) return;

    double R[2][3]  = { { y[6], y[7], y[8] },   { y[12], y[13], y[14] } };
    double mu[2][3] = { { y[9], y[10], y[11] }, { y[15], y[16], y[17] } };
    s->dnu_mag[0] = (nu[0] * mu[0][0] + nu[1] * mu[0][1] + nu[2] * mu[0][2]) / s->nu_mag;
    s->dnu_mag[1] = (nu[0] * mu[1][0] + nu[1] * mu[1][1] + nu[2] * mu[1][2]) / s->nu_mag;
    for (int a = 0; a < 2; a++) {
        s->dc_gr[0][a] = nu[0] / s->nu_mag * s->dc[3 + a] + s->c * mu[a][0] / s->nu_mag - s->c * nu[0] / pow(s->nu_mag, 2) * s->dnu_mag[a] + s->dw[3 + a];
        s->dc_gr[1][a] = nu[1] / s->nu_mag * s->dc[3 + a] + s->c * mu[a][1] / s->nu_mag - s->c * nu[1] / pow(s->nu_mag, 2) * s->dnu_mag[a] + s->dv[3 + a];
        s->dc_gr[2][a] = nu[2] / s->nu_mag * s->dc[3 + a] + s->c * mu[a][2] / s->nu_mag - s->c * nu[2] / pow(s->nu_mag, 2) * s->dnu_mag[a] + s->du[3 + a];
    }
    for (int a = 0; a < 2; a++)
        s->dc_gr_mag[a] = (s->c_gr[0] * s->dc_gr[0][a] + s->c_gr[1] * s->dc_gr[1][a] + s->c_gr[2] * s->dc_gr[2][a]) / s->c_gr_mag;
    for (int a = 0; a < 2; a++) {
        const double* Ra = R[a]; const double* ma = mu[a];
        s->d_GeoCoeff[0][a] = 0.0;
        s->d_GeoCoeff[1][a] = -Ra[0] / (pow(rr, 2));
        if (rd) s->d_GeoCoeff[2][a] = -Ra[0] / (pow(rr, 2) * cos(theta)) + sin(theta) * Ra[1] / (rr * pow(cos(theta), 2));
        else    s->d_GeoCoeff[2][a] = -Ra[0] / (pow(rr, 2) * cos(theta)) + sin(theta) / (rr * pow(cos(theta), 2)) * Ra[1];

        s->d_GeoTerms[0][a] = 0.0;
        s->d_GeoTerms[1][a] = (ma[0] * s->v + nu[0] * s->dv[3 + a] - ma[1] * s->w - nu[1] * s->dw[3 + a]);
        s->d_GeoTerms[2][a] = (ma[0] * s->u + nu[0] * s->du[3 + a] - ma[2] * s->w - nu[2] * s->dw[3 + a]) * cos(theta) - (nu[0] * s->u - nu[2] * s->w) * Ra[1] * sin(theta)
                            + (ma[1] * s->u + nu[1] * s->du[3 + a] - ma[2] * s->v - nu[2] * s->dv[3 + a]) * sin(theta) + (nu[1] * s->u - nu[2] * s->v) * Ra[1] * cos(theta);
        s->d_GeoTerms[0][a] += -Ra[0] / pow(rr, 2) * (nu[1] * s->c_gr[1] + nu[2] * s->c_gr[2])
                             + 1.0 / rr * (ma[1] * s->c_gr[1] + nu[1] * s->dc_gr[1][a] + ma[2] * s->c_gr[2] + nu[2] * s->dc_gr[2][a]);
        s->d_GeoTerms[1][a] += -ma[0] * s->c_gr[1] - nu[0] * s->dc_gr[1][a] + ma[2] * s->c_gr[2] * tan(theta) + nu[2] * s->dc_gr[2][a] * tan(theta) + nu[2] * s->c_gr[2] * Ra[1] / pow(cos(theta), 2);
        s->d_GeoTerms[2][a] += -s->dc_gr[2][a] * (nu[0] * cos(theta) + nu[1] * sin(theta)) - s->c_gr[2] * (ma[0] * cos(theta) - nu[0] * Ra[1] * sin(theta) + ma[1] * sin(theta) + nu[1] * Ra[1] * cos(theta));
    }
}

/* GeoAc_UpdateSources, Global.cpp:222-370 (stratified: scalar wrappers, zero unless the index is radial) */
static void updateg(orc_ray* r, const double* y) {
    srcg* s = SRC(r); orc_atmo* a = ATM(r);
    double rr = y[0], t = y[1], p = y[2];
    s->c = a->c(a, rr, t, p); s->w = 0.0; s->v = a->v(a, rr, t, p); s->u = a->u(a, rr, t, p);
    for (int n = 0; n < 3; n++) {
        s->dc[n] = a->c_diff(a, rr, t, p, n); s->dw[n] = 0.0;
        s->dv[n] = a->v_diff(a, rr, t, p, n); s->du[n] = a->u_diff(a, rr, t, p, n);
    }
    if (r->calc_amp) {
        double R[2][3] = { { y[6], y[7], y[8] }, { y[12], y[13], y[14] } };
        for (int k = 3; k < 5; k++) { s->dc[k] = 0.0; s->dw[k] = 0.0; s->dv[k] = 0.0; s->du[k] = 0.0; }
        for (int n = 0; n < 3; n++) for (int k = 0; k < 2; k++) { s->ddc[n][k] = 0.0; s->ddw[n][k] = 0.0; s->ddv[n][k] = 0.0; s->ddu[n][k] = 0.0; }
        for (int n = 0; n < 3; n++) {
            for (int k = 0; k < 2; k++) {
                s->dc[3 + k] += R[k][n] * a->c_diff(a, rr, t, p, n);
                s->dw[3 + k] += R[k][n] * 0.0;
                s->dv[3 + k] += R[k][n] * a->v_diff(a, rr, t, p, n);
                s->du[3 + k] += R[k][n] * a->u_diff(a, rr, t, p, n);
            }
            for (int m = 0; m < 3; m++) for (int k = 0; k < 2; k++) {
                s->ddc[m][k] += R[k][n] * a->c_ddiff(a, rr, t, p, m, n);
                s->ddw[m][k] += R[k][n] * 0.0;
                s->ddv[m][k] += R[k][n] * a->v_ddiff(a, rr, t, p, m, n);
                s->ddu[m][k] += R[k][n] * a->u_ddiff(a, rr, t, p, m, n);
            }
        }
    }
    geometry_terms(r, y, 0);
}

/* GeoAc_UpdateSources, GlobalRngDep.cpp:226-386 (range dependent: AllOrder1 / AllOrder2 on T, u, v) */
static void updategr(orc_ray* r, const double* y) {
    srcg* s = SRC(r); orc_atmo* a = ATM(r);
    const double gamR = 0.00040187;
    double rr = y[0], t = y[1], p = y[2];
    double temp, dtemp[3];
    if (!r->calc_amp) {
        orc_mspline_allorder1(a, 0, rr, t, p, &temp, dtemp);
        orc_mspline_sync_accel(a);
        orc_mspline_allorder1(a, 1, rr, t, p, &s->u, s->du);
        orc_mspline_allorder1(a, 2, rr, t, p, &s->v, s->dv);
        s->w = 0.0;
        s->c = sqrt(gamR * temp);
        for (int n = 0; n < 3; n++) { s->dc[n] = gamR / (2.0 * s->c) * dtemp[n]; s->dw[n] = 0.0; }
    } else {
        double ddT[3][3], ddU[3][3], ddV[3][3];
        double R[2][3] = { { y[6], y[7], y[8] }, { y[12], y[13], y[14] } };
        orc_mspline_allorder2(a, 0, rr, t, p, &temp, dtemp, ddT);
        orc_mspline_sync_accel(a);
        orc_mspline_allorder2(a, 1, rr, t, p, &s->u, s->du, ddU);
        orc_mspline_allorder2(a, 2, rr, t, p, &s->v, s->dv, ddV);
        s->w = 0.0;
        s->c = sqrt(gamR * temp);
        for (int n = 0; n < 3; n++) {
            s->dc[n] = gamR / (2.0 * s->c) * dtemp[n]; s->dw[n] = 0.0;
            for (int k = 0; k < 2; k++) { s->ddc[n][k] = 0.0; s->ddu[n][k] = 0.0; s->ddv[n][k] = 0.0; s->ddw[n][k] = 0.0; }
            for (int m = 0; m < 3; m++) for (int k = 0; k < 2; k++) {
                s->ddc[n][k] += R[k][m] * (gamR / (2.0 * s->c) * ddT[n][m] - pow(gamR, 2) / (4.0 * pow(s->c, 3)) * dtemp[n] * dtemp[m]);
                s->ddu[n][k] += R[k][m] * ddU[n][m];
                s->ddv[n][k] += R[k][m] * ddV[n][m];
                s->ddw[n][k] += R[k][m] * 0.0;
            }
        }
        for (int k = 3; k < 5; k++) { s->dc[k] = 0.0; s->du[k] = 0.0; s->dv[k] = 0.0; s->dw[k] = 0.0; }
        for (int n = 0; n < 3; n++) for (int k = 0; k < 2; k++) {
            s->dc[3 + k] += R[k][n] * s->dc[n]; s->du[3 + k] += R[k][n] * s->du[n];
            s->dv[3 + k] += R[k][n] * s->dv[n]; s->dw[3 + k] += R[k][n] * s->dw[n];
        }
    }
    geometry_terms(r, y, 1);
}

/* GeoAc_EvalSrcEq, Global.cpp:374-442 (identical in GlobalRngDep.cpp) */
static double rhsg(orc_ray* r, const double* y, int eq) {
    srcg* s = SRC(r);
    double nu[3] = { y[3], y[4], y[5] };
    if (eq < 3) return s->GeoCoeff[eq] * s->c_gr[eq] / s->c_gr_mag;
    if (eq < 6) {
        int n = eq - 3;
        return -s->GeoCoeff[n] / s->c_gr_mag * (s->nu_mag * s->dc[n] + nu[0] * s->dw[n] + nu[1] * s->dv[n] + nu[2] * s->du[n] + s->GeoTerms[n]);
    }
    int a = (eq < 12) ? 0 : 1;
    int base = a ? 12 : 6;
    double mu[3] = { y[base + 3], y[base + 4], y[base + 5] };
    if (eq < base + 3) {
        int n = eq - base;
        return s->d_GeoCoeff[n][a] * s->c_gr[n] / s->c_gr_mag
             + s->GeoCoeff[n] * s->dc_gr[n][a] / s->c_gr_mag
             - s->GeoCoeff[n] * s->c_gr[n] / pow(s->c_gr_mag, 2) * s->dc_gr_mag[a];
    }
    int n = eq - base - 3;
    return -s->d_GeoCoeff[n][a] / s->c_gr_mag * (s->nu_mag * s->dc[n] + nu[0] * s->dw[n] + nu[1] * s->dv[n] + nu[2] * s->du[n] + s->GeoTerms[n])
         + s->GeoCoeff[n] / pow(s->c_gr_mag, 2) * s->dc_gr_mag[a] * (s->nu_mag * s->dc[n] + nu[0] * s->dw[n] + nu[1] * s->dv[n] + nu[2] * s->du[n])
         - s->GeoCoeff[n] / s->c_gr_mag * (s->dnu_mag[a] * s->dc[n] + s->nu_mag * s->ddc[n][a]
                                           + mu[0] * s->dw[n] + mu[1] * s->dv[n] + mu[2] * s->du[n]
                                           + nu[0] * s->ddw[n][a] + nu[1] * s->ddv[n][a] + nu[2] * s->ddu[n][a] + s->d_GeoTerms[n][a]);
}

/* GeoAc_BreakCheck, Global.cpp:500-514 (great-circle range) and GlobalRngDep.cpp:523-535 (lat/lon box) */
static int brkg(orc_ray* r, const double* y) {
    srcg* s = SRC(r);
    double d1 = pow(sin((y[1] - s->src_loc[1]) / 2.0), 2);
    double d2 = cos(s->src_loc[1]) * cos(y[1]) * pow(sin((y[2] - s->src_loc[2]) / 2.0), 2);
    double range = 2.0 * ATM(r)->r_earth * asin(sqrt(d1 + d2));
    int chk = 0;
    if (y[0] > r->prm->vert_limit) chk = 1;
    if (range > r->prm->range_limit) chk = 1;
    return chk;
}
static int brkgr(orc_ray* r, const double* y) {
    const geoac_params* p = r->prm;
    int chk = 0;
    if (y[0] > p->vert_limit) chk = 1;
    if (y[1] < p->box_min[0] || y[1] > p->box_max[0]) chk = 1;
    if (y[2] < p->box_min[1] || y[2] > p->box_max[1]) chk = 1;
    return chk;
}
static int gndg(orc_ray* r, const double* y) { return y[0] < (ATM(r)->r_earth + ATM(r)->z_grnd); }

/* one segment of GeoAc_TravelTime[Segment], Global.cpp:527-589 */
static void ttg(orc_ray* r, const double* ya, const double* yb, double* acc) {
    orc_atmo* a = ATM(r);
    double nu[3], c_prop[3];
    double dr = yb[0] - ya[0], dt = yb[1] - ya[1], dp = yb[2] - ya[2];
    double rr = ya[0] + dr / 2.0, t = ya[1] + dt / 2.0, p = ya[2] + dp / 2.0;
    double ds = sqrt(pow(dr, 2) + pow(rr * dt, 2) + pow(rr * cos(t) * dp, 2));
    nu[0] = ya[3] + (yb[3] - ya[3]) / 2.0;
    nu[1] = ya[4] + (yb[4] - ya[4]) / 2.0;
    nu[2] = ya[5] + (yb[5] - ya[5]) / 2.0;
    double nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    double cm = a->c(a, rr, t, p);
    c_prop[0] = cm * nu[0] / nu_mag + 0.0;
    c_prop[1] = cm * nu[1] / nu_mag + a->v(a, rr, t, p);
    c_prop[2] = cm * nu[2] / nu_mag + a->u(a, rr, t, p);
    double c_prop_mag = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    *acc += ds / c_prop_mag;
}

/* one segment of GeoAc_SB_Atten[Segment], Global.cpp:634-670 (sin(t) where the travel time has cos(t): App. A-6) */
static void sbg(orc_ray* r, const double* ya, const double* yb, double* acc) {
    double dr = yb[0] - ya[0], dt = yb[1] - ya[1], dp = yb[2] - ya[2];
    double rr = ya[0] + dr / 2.0, t = ya[1] + dt / 2.0, p = ya[2] + dp / 2.0;
    double ds = sqrt(pow(dr, 2) + pow(rr * dt, 2) + pow(rr * sin(t) * dp, 2));
    *acc += orc_suthbass_alpha(ATM(r), rr, t, p, r->prm->freq) * ds;
}

/* GeoAc_Jacobian, Global.cpp:594-608 */
static double jacg(orc_ray* r, const double* yk) {
    orc_atmo* a = ATM(r);
    double rr = yk[0], theta = yk[1], phi = yk[2];
    double nu[3] = { yk[3], yk[4], yk[5] };
    double nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    double cc = a->c(a, rr, theta, phi);
    double c_prop[3] = { cc * nu[0] / nu_mag + 0.0, cc * nu[1] / nu_mag + a->v(a, rr, theta, phi), cc * nu[2] / nu_mag + a->u(a, rr, theta, phi) };
    double c_prop_mag = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    double dr_ds = c_prop[0] / c_prop_mag, dt_ds = 1.0 / rr * c_prop[1] / c_prop_mag, dp_ds = 1.0 / (rr * sin(theta)) * c_prop[2] / c_prop_mag;
    double dr_dlt = yk[6], dt_dlt = yk[7], dp_dlt = yk[8];
    double dr_dlp = yk[12], dt_dlp = yk[13], dp_dlp = yk[14];
    return pow(rr, 2) * cos(theta) * (dr_ds * (dt_dlt * dp_dlp - dt_dlp * dp_dlt) - dr_dlt * (dt_ds * dp_dlp - dp_ds * dt_dlp) + dr_dlp * (dt_ds * dp_dlt - dp_ds * dt_dlt));
}

/* GeoAc_Amplitude, Global.cpp:611-629 (c_prop0[1..2] divide by nu_mag, not nu_mag0: App. A-6) */
static double ampg(orc_ray* r, const double* yk) {
    srcg* s = SRC(r); orc_atmo* a = ATM(r);
    double r0 = s->src_loc[0], t0 = s->src_loc[1], p0 = s->src_loc[2];
    double rr = yk[0], theta = yk[1], phi = yk[2];
    double nu[3] = { yk[3], yk[4], yk[5] };
    double nu0[3] = { sin(r->theta), cos(r->theta) * sin(r->phi), cos(r->theta) * cos(r->phi) };
    double cc = a->c(a, rr, theta, phi), vv = a->v(a, rr, theta, phi), uu = a->u(a, rr, theta, phi);
    double nu_mag = (s->c0 - nu[0] * 0.0 - nu[1] * vv - nu[2] * uu) / cc;
    double nu_mag0 = s->nu0;
    double c_prop[3]  = { cc * nu[0] / nu_mag + 0.0, cc * nu[1] / nu_mag + vv, cc * nu[2] / nu_mag + uu };
    double c_prop0[3] = { s->c0 * nu0[0] / nu_mag0 + 0.0, s->c0 * nu0[1] / nu_mag + a->v(a, r0, t0, p0), s->c0 * nu0[2] / nu_mag + a->u(a, r0, t0, p0) };
    double c_prop_mag  = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    double c_prop_mag0 = sqrt(pow(c_prop0[0], 2) + pow(c_prop0[1], 2) + pow(c_prop0[2], 2));
    double D = jacg(r, yk);
    double Amp_Num = a->rho(a, rr, theta, phi) * nu_mag * pow(cc, 3) * c_prop_mag0 * cos(r->theta);
    double Amp_Den = a->rho(a, r0, t0, p0) * nu_mag0 * pow(a->c(a, r0, t0, p0), 3) * c_prop_mag * D;
    return 1.0 / (4.0 * ORC_PI) * sqrt(fabs(Amp_Num / Amp_Den));
}

static double altg(orc_ray* r, const double* y) { return y[0] - ATM(r)->r_earth; }

/* results rows: Code/GeoAcGlobal_main.cpp:294-317, Code/GeoAcGlobal.RngDep_main.cpp:304-328 */
static void fin_common(orc_ray* r, const double* ym1, const double* yk, double tt, double sign, double* incl, double* backaz, double* aux, double* margin) {
    orc_atmo* a = ATM(r);
    const double* src = r->prm->src;     /* z_src, lat, lon [rad] */
    double d1 = pow(sin((yk[1] - src[1]) / 2.0), 2);
    double d2 = cos(src[1]) * cos(yk[1]) * pow(sin((yk[2] - src[2]) / 2.0), 2);
    *incl = sign * asin(a->c(a, yk[0], yk[1], yk[2]) / a->c(a, a->r_earth + src[0], src[1], src[2]) * yk[3]) * 180.0 / ORC_PI;
    double b = 90.0 - atan2(-yk[4], -yk[5]) * 180.0 / ORC_PI;
    if (b < -180.0) b += 360.0;
    if (b > 180.0) b -= 360.0;
    *backaz = b;
    *aux = 2.0 * a->r_earth * asin(sqrt(d1 + d2)) / tt;
    *margin = (yk[0] - (a->r_earth + a->z_grnd)) / fabs(yk[0] - ym1[0]);
}
static void fing(orc_ray* r, const double* ym1, const double* yk, double tt, double* i, double* b, double* x, double* m) { fin_common(r, ym1, yk, tt, -1.0, i, b, x, m); }
static void fingr(orc_ray* r, const double* ym1, const double* yk, double tt, double* i, double* b, double* x, double* m) { fin_common(r, ym1, yk, tt, 1.0, i, b, x, m); }

const orc_eqset orc_eq_global       = { 18, 6, initg, updateg,  rhsg, setdsg, brkg,  gndg, ttg, sbg, ampg, jacg, reflectg, altg, fing };
const orc_eqset orc_eq_globalrngdep = { 18, 6, initg, updategr, rhsg, setdsg, brkgr, gndg, ttg, sbg, ampg, jacg, reflectg, altg, fingr };
