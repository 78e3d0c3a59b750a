/* oracle/orc_eq3d.c -- TEST INFRASTRUCTURE (CPU oracle).
 * 3-D Cartesian stratified moving-medium equation set: restates
 * Code/GeoAc/GeoAc.EquationSets.3DStratified.cpp (line ranges cited per function) with identical expression trees.
 * State y = [x, y, z, nu_z, X_t, Y_t, Z_t, mu_zt, X_p, Y_p, Z_p, mu_zp]  (:84-129).
 */
#include <math.h>
#include "orc_eqsets.h"

/* scratch layout (the reference's GeoAc_Sources struct, :23-55) */
typedef struct src3d {
    double src_loc[3], c0;
    double nu0_xy[2], mu0_xy[2][2];
    double c, dc, ddc, u, du, ddu, v, dv, ddv, w, dw, ddw;
    double nu_mag, dnu_mag[2];
    double c_prop[3], c_prop_mag;
    double dc_prop[3][2], dc_prop_mag[2];
} src3d;

#define SRC(r) ((src3d*)(r)->S)
#define ATM(r) ((r)->atmo)

/* GeoAc_SetInitialConditions :69-131 */
static void init3d(orc_ray* r, double* y) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double x0 = r->prm->src[0], y0 = r->prm->src[1], z0 = r->prm->src[2];
    double th = r->theta, ph = r->phi;
    s->src_loc[0] = x0; s->src_loc[1] = y0; s->src_loc[2] = z0;
    s->c0 = a->c(a, x0, y0, z0);

    double Mc[3]  = { a->u(a, x0, y0, z0) / s->c0, a->v(a, x0, y0, z0) / s->c0, 0.0 / s->c0 };
    double nu0[3] = { cos(th) * cos(ph),  cos(th) * sin(ph), sin(th) };
    double mt[3]  = { -sin(th) * cos(ph), -sin(th) * sin(ph), cos(th) };
    double mp[3]  = { -cos(th) * sin(ph),  cos(th) * cos(ph), 0.0 };

    double M     = 1.0 + (nu0[0] * Mc[0] + nu0[1] * Mc[1] + nu0[2] * Mc[2]);
    double dM_th = mt[0] * Mc[0] + mt[1] * Mc[1] + mt[2] * Mc[2];
    double dM_ph = mp[0] * Mc[0] + mp[1] * Mc[1] + mp[2] * Mc[2];

    s->nu0_xy[0] = nu0[0] / M;
    s->nu0_xy[1] = nu0[1] / M;
    s->mu0_xy[0][0] = mt[0] / M - nu0[0] / pow(M, 2.0) * dM_th;
    s->mu0_xy[1][0] = mt[1] / M - nu0[1] / pow(M, 2.0) * dM_th;
    s->mu0_xy[0][1] = mp[0] / M - nu0[0] / pow(M, 2.0) * dM_ph;
    s->mu0_xy[1][1] = mp[1] / M - nu0[1] / pow(M, 2.0) * dM_ph;

    y[0] = x0; y[1] = y0; y[2] = z0;
    y[3] = nu0[2] / M;
    if (r->eq_cnt > 4) {
        y[4] = y[5] = y[6] = 0.0; y[8] = y[9] = y[10] = 0.0;
        y[7]  = mt[2] / M - nu0[2] / pow(M, 2.0) * dM_th;
        y[11] = mp[2] / M - nu0[2] / pow(M, 2.0) * dM_ph;
    }
}

/* GeoAc_ApproximateIntercept :136-146 + GeoAc_SetReflectionConditions :151-186 */
static void reflect3d(orc_ray* r, const double* ym2, const double* ym1, const double* yk, double* y0) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double zg = a->z_grnd;
    double prev[ORC_MAXEQ];
    double dz_k = yk[2] - ym1[2];
    double dz_grnd = ym1[2] - zg;
    for (int i = 0; i < r->eq_cnt; i++)
        prev[i] = ym1[i] + (ym1[i] - yk[i]) / dz_k * dz_grnd
                + 1.0 / 2.0 * (yk[i] + ym2[i] - 2.0 * ym1[i]) / pow(dz_k, 2.0) * pow(dz_grnd, 2.0);

    double x = prev[0], y = prev[1];
    double cg = a->c(a, x, y, zg);
    double dnuz_ds = -1.0 / cg * (s->c0 / cg * a->c_diff(a, x, y, zg, 2)
                                  + s->nu0_xy[0] * a->u_diff(a, x, y, zg, 2)
                                  + s->nu0_xy[1] * a->v_diff(a, x, y, zg, 2)
                                  + prev[3] * 0.0);
    y0[0] = prev[0]; y0[1] = prev[1]; y0[2] = prev[2];       /* NB: restarts from the fitted z (App. A-20) */
    y0[3] = -prev[3];
    if (r->eq_cnt > 4) {
        y0[4] = prev[4]; y0[5] = prev[5]; y0[8] = prev[8]; y0[9] = prev[9];
        y0[6] = -prev[6]; y0[10] = -prev[10];
        y0[7]  = -prev[7]  + 2.0 * dnuz_ds * prev[6]  / (cg / s->c0 * prev[3]);
        y0[11] = -prev[11] + 2.0 * dnuz_ds * prev[10] / (cg / s->c0 * prev[3]);
    }
}

/* GeoAc_Set_ds :191-198 */
static double setds3d(orc_ray* r, const double* y) {
    double res = 0.05 - 0.049 * exp(-(y[2] - ATM(r)->z_grnd) / 0.75);
    res = fmin(res, r->prm->ds_max);
    res = fmax(res, r->prm->ds_min);
    return res;
}

/* GeoAc_UpdateSources :203-246 */
static void update3d(orc_ray* r, const double* y) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double x = y[0], yy = y[1], z = y[2];
    double nu[3] = { s->nu0_xy[0], s->nu0_xy[1], y[3] };

    s->c = a->c(a, x, yy, z);   s->dc = a->c_diff(a, x, yy, z, 2);
    s->u = a->u(a, x, yy, z);   s->du = a->u_diff(a, x, yy, z, 2);
    s->v = a->v(a, x, yy, z);   s->dv = a->v_diff(a, x, yy, z, 2);
    s->w = 0.0;                 s->dw = 0.0;

    s->nu_mag = s->c0 / s->c * (1.0 - (nu[0] * s->u + nu[1] * s->v + nu[2] * s->w) / s->c0);

    s->c_prop[0] = s->c * nu[0] / s->nu_mag + s->u;
    s->c_prop[1] = s->c * nu[1] / s->nu_mag + s->v;
    s->c_prop[2] = s->c * nu[2] / s->nu_mag + s->w;
    s->c_prop_mag = sqrt(pow(s->c_prop[0], 2) + pow(s->c_prop[1], 2) + pow(s->c_prop[2], 2));

    if (r->calc_amp) {
        double dwinds[3], mu_th[3], mu_ph[3], Zth, Zph;
        s->ddc = a->c_ddiff(a, x, yy, z, 2, 2);  s->ddu = a->u_ddiff(a, x, yy, z, 2, 2);
        s->ddv = a->v_ddiff(a, x, yy, z, 2, 2);  s->ddw = 0.0;

        mu_th[0] = s->mu0_xy[0][0]; mu_th[1] = s->mu0_xy[1][0]; mu_th[2] = y[7];  Zth = y[6];
        mu_ph[0] = s->mu0_xy[0][1]; mu_ph[1] = s->mu0_xy[1][1]; mu_ph[2] = y[11]; Zph = y[10];

        s->dnu_mag[0] = (nu[0] * mu_th[0] + nu[1] * mu_th[1] + nu[2] * mu_th[2]) / s->nu_mag;
        s->dnu_mag[1] = (nu[0] * mu_ph[0] + nu[1] * mu_ph[1] + nu[2] * mu_ph[2]) / s->nu_mag;

        dwinds[0] = a->u_diff(a, x, yy, z, 2);
        dwinds[1] = a->v_diff(a, x, yy, z, 2);
        dwinds[2] = 0.0;
        for (int n = 0; n < 3; n++) {
            s->dc_prop[n][0] = nu[n] / s->nu_mag * s->dc * Zth + s->c * mu_th[n] / s->nu_mag
                             - s->c * nu[n] / pow(s->nu_mag, 2) * s->dnu_mag[0] + dwinds[n] * Zth;
            s->dc_prop[n][1] = nu[n] / s->nu_mag * s->dc * Zph + s->c * mu_ph[n] / s->nu_mag
                             - s->c * nu[n] / pow(s->nu_mag, 2) * s->dnu_mag[1] + dwinds[n] * Zph;
        }
        s->dc_prop_mag[0] = (s->c_prop[0] * s->dc_prop[0][0] + s->c_prop[1] * s->dc_prop[1][0] + s->c_prop[2] * s->dc_prop[2][0]) / s->c_prop_mag;
        s->dc_prop_mag[1] = (s->c_prop[0] * s->dc_prop[0][1] + s->c_prop[1] * s->dc_prop[1][1] + s->c_prop[2] * s->dc_prop[2][1]) / s->c_prop_mag;
    }
}

/* GeoAc_EvalSrcEq :251-310 */
static double rhs3d(orc_ray* r, const double* y, int eq) {
    src3d* s = SRC(r);
    double cp_mag = s->c_prop_mag;
    double nu[3], mu[3];
    if (eq < 3) return s->c_prop[eq] / cp_mag;
    if (eq == 3) {
        nu[0] = s->nu0_xy[0]; nu[1] = s->nu0_xy[1]; nu[2] = y[3];
        return -1.0 / cp_mag * (s->nu_mag * s->dc + (nu[0] * s->du + nu[1] * s->dv + nu[2] * s->dw));
    }
    if (eq == 7 || eq == 11) {
        int a = (eq == 7) ? 0 : 1;
        nu[0] = s->nu0_xy[0]; mu[0] = s->mu0_xy[0][a];
        nu[1] = s->nu0_xy[1]; mu[1] = s->mu0_xy[1][a];
        nu[2] = y[3];         mu[2] = y[eq];
        return 1.0 / pow(cp_mag, 2) * (s->nu_mag * s->dc + (nu[0] * s->du + nu[1] * s->dv + nu[2] * s->dw)) * s->dc_prop_mag[a]
             - 1.0 / cp_mag * (s->dnu_mag[a] * s->dc + (mu[0] * s->du + mu[1] * s->dv + mu[2] * s->dw
                               + (s->nu_mag * s->ddc + nu[0] * s->ddu + nu[1] * s->ddv + nu[2] * s->ddw) * y[eq - 1]));
    }
    if (eq < 7)  return s->dc_prop[eq - 4][0] / cp_mag - s->c_prop[eq - 4] / pow(cp_mag, 2) * s->dc_prop_mag[0];
    return s->dc_prop[eq - 8][1] / cp_mag - s->c_prop[eq - 8] / pow(cp_mag, 2) * s->dc_prop_mag[1];
}

/* GeoAc_BreakCheck :327-336, GeoAc_GroundCheck :338-343 */
static int brk3d(orc_ray* r, const double* y) {
    double rr = sqrt(pow(y[0], 2) + pow(y[1], 2));
    int chk = 0;
    if (y[2] > r->prm->vert_limit) chk = 1;
    if (rr > r->prm->range_limit) chk = 1;
    return chk;
}
static int gnd3d(orc_ray* r, const double* y) { return y[2] < ATM(r)->z_grnd; }

/* one segment of GeoAc_TravelTime / GeoAc_TravelTimeSegment :348-405 (c(0,0,0) and missing w: App. A-4) */
static void tt3d(orc_ray* r, const double* ya, const double* yb, double* acc) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double nu[3], c_prop[3];
    nu[0] = s->nu0_xy[0]; nu[1] = s->nu0_xy[1];
    double dx = yb[0] - ya[0], dy = yb[1] - ya[1], dz = yb[2] - ya[2];
    double ds = sqrt(dx * dx + dy * dy + dz * dz);
    double x = ya[0] + dx / 2.0, y = ya[1] + dy / 2.0, z = ya[2] + dz / 2.0;
    nu[2] = ya[3] + (yb[3] - ya[3]) / 2.0;
    double cm = a->c(a, x, y, z), um = a->u(a, x, y, z), vm = a->v(a, x, y, z);
    double nu_mag = (a->c(a, 0, 0, 0) - nu[0] * um - nu[1] * vm) / cm;
    c_prop[0] = cm * nu[0] / nu_mag + um;
    c_prop[1] = cm * nu[1] / nu_mag + vm;
    c_prop[2] = cm * nu[2] / nu_mag;
    double c_prop_mag = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    *acc += ds / c_prop_mag;
}

/* one segment of GeoAc_SB_Atten / GeoAc_SB_AttenSegment :456-490 */
static void sb3d(orc_ray* r, const double* ya, const double* yb, double* acc) {
    double dx = yb[0] - ya[0], dy = yb[1] - ya[1], dz = yb[2] - ya[2];
    double ds = sqrt(dx * dx + dy * dy + dz * dz);
    double x = ya[0] + dx / 2.0, y = ya[1] + dy / 2.0, z = ya[2] + dz / 2.0;
    *acc += orc_suthbass_alpha(ATM(r), x, y, z, r->prm->freq) * ds;
}

/* GeoAc_Jacobian :410-428 */
static double jac3d(orc_ray* r, const double* yk) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double x = yk[0], y = yk[1], z = yk[2];
    double x0 = s->src_loc[0], y0 = s->src_loc[1], z0 = s->src_loc[2];
    double nu[3] = { s->nu0_xy[0], s->nu0_xy[1], yk[3] };
    double cc = a->c(a, x, y, z), uu = a->u(a, x, y, z), vv = a->v(a, x, y, z);
    double nu_mag = (a->c(a, x0, y0, z0) - nu[0] * uu - nu[1] * vv) / cc;
    double c_prop[3] = { cc * nu[0] / nu_mag + uu, cc * nu[1] / nu_mag + vv, cc * nu[2] / nu_mag + 0.0 };
    double c_prop_mag = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    double dxds = c_prop[0] / c_prop_mag, dyds = c_prop[1] / c_prop_mag, dzds = c_prop[2] / c_prop_mag;
    double dxdt = yk[4], dydt = yk[5], dzdt = yk[6];
    double dxdp = yk[8], dydp = yk[9], dzdp = yk[10];
    return dxds * (dydt * dzdp - dydp * dzdt)
         - dxdt * (dyds * dzdp - dzds * dydp)
         + dxdp * (dyds * dzdt - dzds * dydt);
}

/* GeoAc_Amplitude :431-451 (sign slip in nu_mag0 kept: App. A-5) */
static double amp3d(orc_ray* r, const double* yk) {
    src3d* s = SRC(r); orc_atmo* a = ATM(r);
    double x = yk[0], y = yk[1], z = yk[2];
    double x0 = s->src_loc[0], y0 = s->src_loc[1], z0 = s->src_loc[2];
    double nu[3] = { s->nu0_xy[0], s->nu0_xy[1], yk[3] };
    double cc = a->c(a, x, y, z), uu = a->u(a, x, y, z), vv = a->v(a, x, y, z);
    double c00 = a->c(a, x0, y0, z0), u00 = a->u(a, x0, y0, z0), v00 = a->v(a, x0, y0, z0);
    double nu_mag  = (c00 - nu[0] * uu - nu[1] * vv) / cc;
    double nu_mag0 = 1.0 - (nu[0] * u00 - nu[1] * v00) / c00;
    double c_prop[3]  = { cc * nu[0] / nu_mag + uu, cc * nu[1] / nu_mag + vv, cc * nu[2] / nu_mag };
    double c_prop0[3] = { c00 * nu[0] / nu_mag0 + u00, c00 * nu[1] / nu_mag0 + v00,
                          c00 * sqrt(1.0 - pow(nu[0] / nu_mag0, 2) - pow(nu[1] / nu_mag0, 2)) };
    double c_prop_mag  = sqrt(pow(c_prop[0], 2) + pow(c_prop[1], 2) + pow(c_prop[2], 2));
    double c_prop_mag0 = sqrt(pow(c_prop0[0], 2) + pow(c_prop0[1], 2) + pow(c_prop0[2], 2));
    double D = jac3d(r, yk);
    double Amp_Num = a->rho(a, x, y, z) * nu_mag * pow(cc, 3) * c_prop_mag0 * cos(r->theta);
    double Amp_Den = a->rho(a, x0, y0, z0) * nu_mag0 * pow(c00, 3) * c_prop_mag * D;
    return 1.0 / (4.0 * ORC_PI) * sqrt(fabs(Amp_Num / Amp_Den));
}

static double alt3d(orc_ray* r, const double* y) { (void)r; return y[2]; }

/* results row of Code/GeoAc3D_main.cpp:281-284 */
static void fin3d(orc_ray* r, const double* ym1, const double* yk, double tt, double* incl, double* backaz, double* aux, double* margin) {
    orc_atmo* a = ATM(r); (void)tt;
    const double* src = r->prm->src;
    double az = (ORC_PI / 2.0 - r->phi) * 180.0 / ORC_PI;        /* the main's loop variable phi [deg] */
    double b = az + 180.0;
    *incl = -asin(a->c(a, yk[0], yk[1], a->z_grnd) / a->c(a, src[0], src[1], src[2]) * yk[3]) * 180.0 / ORC_PI;
    while (b > 180.0) b -= 360.0;
    while (b < -180.0) b += 360.0;
    *backaz = b; *aux = 0.0;
    *margin = (yk[2] - a->z_grnd) / fabs(yk[2] - ym1[2]);
}

const orc_eqset orc_eq_3d = { 12, 4, init3d, update3d, rhs3d, setds3d, brk3d, gnd3d, tt3d, sb3d, amp3d, jac3d, reflect3d, alt3d, fin3d };
