/* oracle/orc_eq3drngdep.c -- TEST INFRASTRUCTURE (CPU oracle).
 * 3-D Cartesian range-dependent moving medium: restates Code/GeoAc/GeoAc.EquationSets.3DRngDep.cpp with identical
 * expression trees.  State y = [x, y, z, nu_x, nu_y, nu_z, X_th(3), mu_th(3), X_ph(3), mu_ph(3)].
 * The RK4 stages sample the atmosphere through Eval_Spline_AllOrder1/2 (orc_mspline.c); everything else (initial
 * conditions, reflection, travel time, absorption, amplitude) goes through the scalar wrappers c(), u(), v(), rho().
 */
#include <math.h>
#include "orc_eqsets.h"

typedef struct src3r {                      /* GeoAc_Sources, 3DRngDep.cpp:24-64 */
    double src_loc[3], c0;
    double c, dc[5], ddc[3][2];
    double u, du[5], ddu[3][2];
    double v, dv[5], ddv[3][2];
    double w, dw[5], ddw[3][2];
    double nu0, nu_mag, dnu_mag[2];
    double c_gr[3], c_gr_mag, dc_gr[3][2], dc_gr_mag[2];
} src3r;

#define SRC(r) ((src3r*)(r)->S)
#define ATM(r) ((r)->atmo)

void orc_mspline_allorder1(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3]);
void orc_mspline_allorder2(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3], double dd[3][3]);
void orc_mspline_sync_accel(orc_atmo* a);
void orc_mspline_reset_accel(orc_atmo* a);

/* GeoAc_SetInitialConditions, 3DRngDep.cpp:70-136 */
static void init3r(orc_ray* r, double* y) {
    src3r* s = SRC(r); orc_atmo* a = ATM(r);
    double x0 = r->prm->src[0], y0 = r->prm->src[1], z0 = fmax(r->prm->z_grnd, r->prm->src[2]);
    double th = r->theta, ph = r->phi;
    s->src_loc[0] = x0; s->src_loc[1] = y0; s->src_loc[2] = z0;
    s->c0 = a->c(a, x0, y0, z0);
    double Mc[3]  = { a->u(a, x0, y0, z0) / s->c0, a->v(a, x0, y0, z0) / s->c0, 0.0 / s->c0 };
    double nu0[3] = { cos(th) * cos(ph),  cos(th) * sin(ph), sin(th) };
    double mth[3] = { -sin(th) * cos(ph), -sin(th) * sin(ph), cos(th) };
    double mph[3] = { -cos(th) * sin(ph),  cos(th) * cos(ph), 0.0 };
    double MS = 1.0 + (nu0[0] * Mc[0] + nu0[1] * Mc[1] + nu0[2] * Mc[2]);
    s->nu0 = 1.0 / MS;
    y[0] = x0; y[1] = y0; y[2] = z0;
    for (int i = 0; i < 3; i++) y[3 + i] = nu0[i] / MS;
    if (r->eq_cnt > 6) {
        for (int i = 0; i < 3; i++) { y[6 + i] = 0.0; y[12 + i] = 0.0; }
        for (int i = 0; i < 3; i++) {
            y[9 + i]  = mth[i] / MS - nu0[i] / pow(MS, 2.0) * (mth[0] * Mc[0] + mth[1] * Mc[1] + mth[2] * Mc[2]);
            y[15 + i] = mph[i] / MS - nu0[i] / pow(MS, 2.0) * (mph[0] * Mc[0] + mph[1] * Mc[1] + mph[2] * Mc[2]);
        }
    }
    orc_mspline_reset_accel(a);             /* :130-134 */
}

/* GeoAc_ApproximateIntercept (second order) + GeoAc_SetReflectionConditions, 3DRngDep.cpp:142-201 */
static void reflect3r(orc_ray* r, const double* ym2, const double* ym1, const double* yk, double* y0) {
    src3r* s = SRC(r); orc_atmo* a = ATM(r);
    double prev[ORC_MAXEQ];
    double zg = a->z_grnd;
    double dz_k = yk[2] - ym1[2];
    double dz_grnd = ym1[2] - zg;
    for (int i = 0; i < r->eq_cnt; i++)
        prev[i] = ym1[i] + (ym1[i] - yk[i]) / dz_k * dz_grnd
                + 1.0 / 2.0 * (yk[i] + ym2[i] - 2.0 * ym1[i]) / pow(dz_k, 2.0) * pow(dz_grnd, 2.0);
    double c_grnd = a->c(a, prev[0], prev[1], zg);
    double dnuz_ds = -1.0 / c_grnd * (s->c0 / c_grnd * a->c_diff(a, prev[0], prev[1], zg, 2)
                                      + prev[3] * a->u_diff(a, prev[0], prev[1], zg, 2)
                                      + prev[4] * a->v_diff(a, prev[0], prev[1], zg, 2)
                                      + prev[5] * 0.0);
    for (int i = 0; i < r->eq_cnt; i++) y0[i] = prev[i];
    y0[2] = zg;
    y0[5] = -prev[5];
    if (r->eq_cnt > 6) {
        y0[8] = -prev[8]; y0[14] = -prev[14];
        y0[11] = -prev[11] + 2.0 * dnuz_ds * prev[8]  / (c_grnd / s->c0 * prev[5]);
        y0[17] = -prev[17] + 2.0 * dnuz_ds * prev[14] / (c_grnd / s->c0 * prev[5]);
    }
}

/* GeoAc_Set_ds, 3DRngDep.cpp:206-213 */
static double setds3r(orc_ray* r, const double* y) {
    double res = 0.05 - 0.049 * exp(-(y[2] - ATM(r)->z_grnd) / 0.75);
    res = fmin(res, r->prm->ds_max);
    res = fmax(res, r->prm->ds_min);
    return res;
}

/* GeoAc_UpdateSources, 3DRngDep.cpp:218-326 */
static void update3r(orc_ray* r, const double* y) {
    src3r* s = SRC(r); orc_atmo* a = ATM(r);
    const double gamR = 0.00040187;
    double x = y[0], yy = y[1], z = y[2];
    double nu[3] = { y[3], y[4], y[5] };
    double temp, dtemp[3];
    if (!r->calc_amp) {
        orc_mspline_allorder1(a, 0, x, yy, z, &temp, dtemp);
        orc_mspline_sync_accel(a);
        orc_mspline_allorder1(a, 1, x, yy, z, &s->u, s->du);
        orc_mspline_allorder1(a, 2, x, yy, z, &s->v, s->dv);
        s->w = 0.0;
        s->c = sqrt(gamR * temp);
        for (int n = 0; n < 3; n++) { s->dc[n] = gamR / (2.0 * s->c) * dtemp[n]; s->dw[n] = 0.0; }
        s->nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
        s->c_gr[0] = s->c * nu[0] / s->nu_mag + s->u;
        s->c_gr[1] = s->c * nu[1] / s->nu_mag + s->v;
        s->c_gr[2] = s->c * nu[2] / s->nu_mag + s->w;
        s->c_gr_mag = sqrt(pow(s->c_gr[0], 2) + pow(s->c_gr[1], 2) + pow(s->c_gr[2], 2));
        return;
    }
    double Xa[2][3] = { { y[6], y[7], y[8] },   { y[12], y[13], y[14] } };
    double mu[2][3] = { { y[9], y[10], y[11] }, { y[15], y[16], y[17] } };
    double ddT[3][3], ddU[3][3], ddV[3][3];
    orc_mspline_allorder2(a, 0, x, yy, z, &temp, dtemp, ddT);
    orc_mspline_sync_accel(a);
    orc_mspline_allorder2(a, 1, x, yy, z, &s->u, s->du, ddU);
    orc_mspline_allorder2(a, 2, x, yy, z, &s->v, s->dv, ddV);
    s->w = 0.0;
    s->c = sqrt(gamR * temp);
    for (int n = 0; n < 3; n++) {
        s->dc[n] = gamR / (2.0 * s->c) * dtemp[n];
        s->dw[n] = 0.0;
        for (int k = 0; k < 2; k++) { s->ddc[n][k] = 0.0; s->ddu[n][k] = 0.0; s->ddv[n][k] = 0.0; s->ddw[n][k] = 0.0; }
        for (int m = 0; m < 3; m++) for (int k = 0; k < 2; k++) {
            s->ddc[n][k] += Xa[k][m] * (gamR / (2.0 * s->c) * ddT[n][m] - pow(gamR, 2) / (4.0 * pow(s->c, 3)) * dtemp[n] * dtemp[m]);
            s->ddu[n][k] += Xa[k][m] * ddU[n][m];
            s->ddv[n][k] += Xa[k][m] * ddV[n][m];
            s->ddw[n][k] += Xa[k][m] * 0.0;
        }
    }
    for (int k = 3; k < 5; k++) { s->dc[k] = 0.0; s->du[k] = 0.0; s->dv[k] = 0.0; s->dw[k] = 0.0; }
    for (int n = 0; n < 3; n++) for (int k = 0; k < 2; k++) {
        s->dc[3 + k] += Xa[k][n] * s->dc[n]; s->du[3 + k] += Xa[k][n] * s->du[n];
        s->dv[3 + k] += Xa[k][n] * s->dv[n]; s->dw[3 + k] += Xa[k][n] * s->dw[n];
    }
    s->nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    for (int k = 0; k < 2; k++) s->dnu_mag[k] = (nu[0] * mu[k][0] + nu[1] * mu[k][1] + nu[2] * mu[k][2]) / s->nu_mag;
    s->c_gr[0] = s->c * nu[0] / s->nu_mag + s->u;
    s->c_gr[1] = s->c * nu[1] / s->nu_mag + s->v;
    s->c_gr[2] = s->c * nu[2] / s->nu_mag + s->w;
    s->c_gr_mag = sqrt(pow(s->c_gr[0], 2) + pow(s->c_gr[1], 2) + pow(s->c_gr[2], 2));
    for (int k = 0; k < 2; k++) {
        s->dc_gr[0][k] = nu[0] / s->nu_mag * s->dc[3 + k] + s->c * mu[k][0] / s->nu_mag - s->c * nu[0] / pow(s->nu_mag, 2) * s->dnu_mag[k] + s->du[3 + k];
        s->dc_gr[1][k] = nu[1] / s->nu_mag * s->dc[3 + k] + s->c * mu[k][1] / s->nu_mag - s->c * nu[1] / pow(s->nu_mag, 2) * s->dnu_mag[k] + s->dv[3 + k];
        s->dc_gr[2][k] = nu[2] / s->nu_mag * s->dc[3 + k] + s->c * mu[k][2] / s->nu_mag - s->c * nu[2] / pow(s->nu_mag, 2) * s->dnu_mag[k] + s->dw[3 + k];
        s->dc_gr_mag[k] = (s->c_gr[0] * s->dc_gr[0][k] + s->c_gr[1] * s->dc_gr[1][k] + s->c_gr[2] * s->dc_gr[2][k]) / s->c_gr_mag;
    }
}

/* GeoAc_EvalSrcEq, 3DRngDep.cpp:331-393 */
static double rhs3r(orc_ray* r, const double* y, int eq) {
    src3r* s = SRC(r);
    double nu[3] = { y[3], y[4], y[5] };
    if (eq < 3) return s->c_gr[eq] / s->c_gr_mag;
    if (eq < 6) {
        int n = eq - 3;
        return -1.0 / s->c_gr_mag * (s->nu_mag * s->dc[n] + nu[0] * s->du[n] + nu[1] * s->dv[n] + nu[2] * s->dw[n]);
    }
    int k = (eq < 12) ? 0 : 1, base = k ? 12 : 6;
    double mu[3] = { y[base + 3], y[base + 4], y[base + 5] };
    if (eq < base + 3) {
        int n = eq - base;
        return s->dc_gr[n][k] / s->c_gr_mag - s->c_gr[n] / pow(s->c_gr_mag, 2) * s->dc_gr_mag[k];
    }
    int n = eq - base - 3;
    return 1.0 / pow(s->c_gr_mag, 2) * s->dc_gr_mag[k] * (s->nu_mag * s->dc[n] + nu[0] * s->du[n] + nu[1] * s->dv[n] + nu[2] * s->dw[n])
         - 1.0 / s->c_gr_mag * (s->dnu_mag[k] * s->dc[n] + s->nu_mag * s->ddc[n][k]
                                + mu[0] * s->du[n] + mu[1] * s->dv[n] + mu[2] * s->dw[n]
                                + nu[0] * s->ddu[n][k] + nu[1] * s->ddv[n][k] + nu[2] * s->ddw[n][k]);
}

/* GeoAc_BreakCheck / GeoAc_GroundCheck, 3DRngDep.cpp:451-472 */
static int brk3r(orc_ray* r, const double* y) {
    const geoac_params* p = r->prm;
    int chk = 0;
    if (y[0] > p->box_max[0]) chk = 1;
    if (y[0] < p->box_min[0]) chk = 1;
    if (y[1] > p->box_max[1]) chk = 1;
    if (y[1] < p->box_min[1]) chk = 1;
    if (y[2] > p->vert_limit) chk = 1;
    return chk;
}
static int gnd3r(orc_ray* r, const double* y) { return y[2] < ATM(r)->z_grnd; }

/* one segment of GeoAc_TravelTime[Segment], 3DRngDep.cpp:478-542 */
static void tt3r(orc_ray* r, const double* ya, const double* yb, double* acc) {
    orc_atmo* a = ATM(r);
    double dx = yb[0] - ya[0], dy = yb[1] - ya[1], dz = yb[2] - ya[2];
    double ds = sqrt(dx * dx + dy * dy + dz * dz);
    double x = ya[0] + dx / 2.0, y = ya[1] + dy / 2.0, z = ya[2] + dz / 2.0;
    double nu[3] = { ya[3] + (yb[3] - ya[3]) / 2.0, ya[4] + (yb[4] - ya[4]) / 2.0, ya[5] + (yb[5] - ya[5]) / 2.0 };
    double nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    double cs = a->c(a, x, y, z);
    double cp[3] = { cs * nu[0] / nu_mag + a->u(a, x, y, z), cs * nu[1] / nu_mag + a->v(a, x, y, z), cs * nu[2] / nu_mag + 0.0 };
    double cpm = sqrt(pow(cp[0], 2) + pow(cp[1], 2) + pow(cp[2], 2));
    *acc += ds / cpm;
}

/* one segment of GeoAc_SB_Atten[Segment], 3DRngDep.cpp:597-635 */
static void sb3r(orc_ray* r, const double* ya, const double* yb, double* acc) {
    double dx = yb[0] - ya[0], dy = yb[1] - ya[1], dz = yb[2] - ya[2];
    double ds = sqrt(dx * dx + dy * dy + dz * dz);
    double x = ya[0] + dx / 2.0, y = ya[1] + dy / 2.0, z = ya[2] + dz / 2.0;
    *acc += orc_suthbass_alpha(ATM(r), x, y, z, r->prm->freq) * ds;
}

/* GeoAc_Jacobian + GeoAc_Amplitude, 3DRngDep.cpp:547-592 */
static double jac3r(orc_ray* r, const double* yk) {
    orc_atmo* a = ATM(r);
    double x = yk[0], y = yk[1], z = yk[2];
    double nu[3] = { yk[3], yk[4], yk[5] };
    double nu_mag = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    double cs = a->c(a, x, y, z);
    double cp[3] = { cs * nu[0] / nu_mag + a->u(a, x, y, z), cs * nu[1] / nu_mag + a->v(a, x, y, z), cs * nu[2] / nu_mag + 0.0 };
    double cpm = sqrt(pow(cp[0], 2) + pow(cp[1], 2) + pow(cp[2], 2));
    double dxds = cp[0] / cpm, dyds = cp[1] / cpm, dzds = cp[2] / cpm;
    double dxdt = yk[6], dydt = yk[7], dzdt = yk[8], dxdp = yk[12], dydp = yk[13], dzdp = yk[14];
    return dxds * (dydt * dzdp - dydp * dzdt) - dxdt * (dyds * dzdp - dzds * dydp) + dxdp * (dyds * dzdt - dzds * dydt);
}
static double amp3r(orc_ray* r, const double* yk) {
    src3r* s = SRC(r); orc_atmo* a = ATM(r);
    double x = yk[0], y = yk[1], z = yk[2];
    double x0 = s->src_loc[0], y0 = s->src_loc[1], z0 = s->src_loc[2];
    double nu[3] = { yk[3], yk[4], yk[5] };
    double c0 = s->c0, cs = a->c(a, x, y, z), wu = a->u(a, x, y, z), wv = a->v(a, x, y, z), ww = 0.0;
    double wu0 = a->u(a, x0, y0, z0), wv0 = a->v(a, x0, y0, z0), ww0 = 0.0;
    double nu_mag = (c0 - nu[0] * wu - nu[1] * wv - nu[2] * ww) / cs;
    double nu_mag0 = 1.0 - nu[0] * wu0 / c0 - nu[1] * wv0 / c0 - nu[2] * ww0 / c0;
    double cp[3]  = { cs * nu[0] / nu_mag + wu, cs * nu[1] / nu_mag + wv, cs * nu[2] / nu_mag + ww };
    double cp0[3] = { c0 * cos(r->theta) * cos(r->phi) + wu0, c0 * cos(r->theta) * sin(r->phi) + wv0, c0 * sin(r->theta) + ww0 };
    double cpm = sqrt(pow(cp[0], 2) + pow(cp[1], 2) + pow(cp[2], 2));
    double cpm0 = sqrt(pow(cp0[0], 2) + pow(cp0[1], 2) + pow(cp0[2], 2));
    double D = jac3r(r, yk);
    double num = a->rho(a, x, y, z) * nu_mag * pow(cs, 3) * cpm0 * cos(r->theta);
    double den = a->rho(a, x0, y0, z0) * nu_mag0 * pow(c0, 3) * cpm * D;
    return 1.0 / (4.0 * ORC_PI) * sqrt(fabs(num / den));
}

static double alt3r(orc_ray* r, const double* y) { (void)r; return y[2]; }

/* results row, Code/GeoAc3D.RngDep_main.cpp:296-318 */
static void fin3r(orc_ray* r, const double* ym1, const double* yk, double tt, double* incl, double* backaz, double* aux, double* margin) {
    orc_atmo* a = ATM(r); src3r* s = SRC(r); (void)tt;
    *incl = -asin(a->c(a, yk[0], yk[1], a->z_grnd) / a->c(a, s->src_loc[0], s->src_loc[1], s->src_loc[2]) * yk[5]) * 180.0 / ORC_PI;
    double b = 90.0 - atan2(-yk[4], -yk[3]) * 180.0 / ORC_PI;
    while (b < -180.0) b += 360.0;
    while (b > 180.0) b -= 360.0;
    *backaz = b; *aux = 0.0;
    *margin = (yk[2] - a->z_grnd) / fabs(yk[2] - ym1[2]);
}

const orc_eqset orc_eq_3drngdep = { 18, 6, init3r, update3r, rhs3r, setds3r, brk3r, gnd3r, tt3r, sb3r, amp3r, jac3r, reflect3r, alt3r, fin3r };
