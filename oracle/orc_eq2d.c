/* oracle/orc_eq2d.c -- TEST INFRASTRUCTURE (CPU oracle).
 * 2-D effective-sound-speed equation set: restates Code/GeoAc/GeoAc.EquationSets.2DStratified.cpp with identical
 * expression trees.  State y = [r, z, zeta(=nu_z), R_t, Z_t, eta(=mu_z)]  (:41-66).
 */
#include <math.h>
#include "orc_eqsets.h"

typedef struct src2d { double c_eff, c_eff_0, c_eff_diff, c_eff_ddiff; } src2d;   /* :23-29 */
#define SRC(r) ((src2d*)(r)->S)
#define ATM(r) ((r)->atmo)

/* GeoAc_SetInitialConditions :38-68 */
static void init2d(orc_ray* r, double* y) {
    orc_atmo* a = ATM(r); double z0 = r->prm->src[2];
    SRC(r)->c_eff_0 = a->c(a, 0.0, 0.0, z0) + a->u(a, 0.0, 0.0, z0) * cos(r->phi) + a->v(a, 0.0, 0.0, z0) * sin(r->phi);
    y[0] = 0.0; y[1] = z0; y[2] = sin(r->theta);
    if (r->eq_cnt > 3) { y[3] = 0.0; y[4] = 0.0; y[5] = cos(r->theta); }
}

/* GeoAc_ApproximateIntercept :74-83 + GeoAc_SetReflectionConditions :88-117 */
static void reflect2d(orc_ray* r, const double* ym2, const double* ym1, const double* yk, double* y0) {
    orc_atmo* a = ATM(r); double zg = a->z_grnd;
    double prev[ORC_MAXEQ];
    double dz_k = yk[1] - ym1[1];
    double dz_grnd = ym1[1] - zg;
    for (int i = 0; i < r->eq_cnt; i++)
        prev[i] = ym1[i] + (ym1[i] - yk[i]) / dz_k * dz_grnd
                + 1.0 / 2.0 * (yk[i] + ym2[i] - 2.0 * ym1[i]) / pow(dz_k, 2.0) * pow(dz_grnd, 2.0);
    double c_eff_diff = a->c_diff(a, 0.0, 0.0, zg, 2) + a->u_diff(a, 0.0, 0.0, zg, 2) * cos(r->phi) + a->v_diff(a, 0.0, 0.0, zg, 2) * sin(r->phi);
    double cg = a->c(a, 0.0, 0.0, zg);
    double dnuz_ds = -SRC(r)->c_eff_0 / pow(cg, 2) * c_eff_diff;
    y0[0] = prev[0]; y0[1] = zg; y0[2] = -prev[2];
    if (r->eq_cnt > 3) {
        y0[3] = prev[3]; y0[4] = -prev[4];
        y0[5] = -prev[5] + 2.0 * dnuz_ds * prev[4] / (cg / SRC(r)->c_eff_0 * prev[2]);
    }
}

/* GeoAc_Set_ds :123-130 */
static double setds2d(orc_ray* r, const double* y) {
    double res = 0.05 - 0.049 * exp(-(y[1] - ATM(r)->z_grnd) / 0.75);
    res = fmin(res, r->prm->ds_max);
    res = fmax(res, r->prm->ds_min);
    return res;
}

/* GeoAc_UpdateSources :135-147 */
static void update2d(orc_ray* r, const double* y) {
    src2d* s = SRC(r); orc_atmo* a = ATM(r); double z = y[1];
    s->c_eff = a->c(a, 0, 0, z) + a->u(a, 0, 0, z) * cos(r->phi) + a->v(a, 0, 0, z) * sin(r->phi);
    s->c_eff_diff = a->c_diff(a, 0, 0, z, 2) + a->u_diff(a, 0, 0, z, 2) * cos(r->phi) + a->v_diff(a, 0, 0, z, 2) * sin(r->phi);
    if (r->calc_amp)
        s->c_eff_ddiff = a->c_ddiff(a, 0, 0, z, 2, 2) + a->u_ddiff(a, 0, 0, z, 2, 2) * cos(r->phi) + a->v_ddiff(a, 0, 0, z, 2, 2) * sin(r->phi);
}

/* GeoAc_EvalSrcEq :152-181 */
static double rhs2d(orc_ray* r, const double* y, int eq) {
    src2d* s = SRC(r);
    double nu_z = y[2], dzt = y[4], mu_z = y[5];
    double c = s->c_eff, c0 = s->c_eff_0, dc = s->c_eff_diff, ddc = s->c_eff_ddiff;
    switch (eq) {
        case 0: return c / c0 * cos(r->theta);
        case 1: return c / c0 * nu_z;
        case 2: return -c0 / pow(c, 2) * dc;
        case 3: return dc * dzt / c0 * cos(r->theta) - c / c0 * sin(r->theta);
        case 4: return dc * dzt / c0 * nu_z + c / c0 * mu_z;
        default: return (2 * pow(dc / c, 2) - ddc / c) * c0 / c * dzt;
    }
}

/* GeoAc_BreakCheck :194-203, GeoAc_GroundCheck :204-212 */
static int brk2d(orc_ray* r, const double* y) {
    int chk = 0;
    if (y[1] > r->prm->vert_limit) chk = 1;
    if (y[0] > r->prm->range_limit) chk = 1;
    return chk;
}
static int gnd2d(orc_ray* r, const double* y) { return y[1] < ATM(r)->z_grnd; }

/* one segment of GeoAc_TravelTime[Segment] :217-249 */
static void tt2d(orc_ray* r, const double* ya, const double* yb, double* acc) {
    orc_atmo* a = ATM(r);
    double dr = yb[0] - ya[0], dz = yb[1] - ya[1];
    double z_avg = ya[1] + dz / 2.0;
    double c_eff = a->c(a, 0, 0, z_avg) + a->u(a, 0, 0, z_avg) * cos(r->phi) + a->v(a, 0, 0, z_avg) * sin(r->phi);
    double ds = sqrt(pow(dr, 2) + pow(dz, 2));
    *acc += ds / c_eff;
}

/* one segment of GeoAc_SB_Atten[Segment] :254-286 */
static void sb2d(orc_ray* r, const double* ya, const double* yb, double* acc) {
    double dr = yb[0] - ya[0], dz = yb[1] - ya[1];
    double ds = sqrt(dr * dr + dz * dz);
    double x = (ya[0] + dr / 2.0) * cos(r->phi), y = (ya[0] + dr / 2.0) * sin(r->phi), z = ya[1] + dz / 2.0;
    *acc += orc_suthbass_alpha(ATM(r), x, y, z, r->prm->freq) * ds;
}

/* GeoAc_Jacobian :291-300 (thermodynamic c, not c_eff) and GeoAc_Amplitude :303-312 */
static double amp2d(orc_ray* r, const double* yk) {
    src2d* s = SRC(r); orc_atmo* a = ATM(r);
    double rr = yk[0], z = yk[1];
    double cz = a->c(a, 0.0, 0.0, z);
    double drds = cz / s->c_eff_0 * cos(r->theta);
    double dzds = cz / s->c_eff_0 * yk[2];
    double D = rr * (drds * yk[4] - dzds * yk[3]);
    double Amp_Num = a->rho(a, 0.0, 0.0, z) * cz * cos(r->theta);
    double Amp_Den = a->rho(a, 0.0, 0.0, a->z_grnd) * s->c_eff_0 * D;
    return 1.0 / (4.0 * ORC_PI) * sqrt(fabs(Amp_Num / Amp_Den));
}

/* GeoAc_Jacobian :291-300 on its own (WriteCaustics) */
static double jac2d(orc_ray* r, const double* yk) {
    src2d* s = SRC(r); orc_atmo* a = ATM(r);
    double rr = yk[0], z = yk[1];
    double cz = a->c(a, 0.0, 0.0, z);
    double drds = cz / s->c_eff_0 * cos(r->theta);
    double dzds = cz / s->c_eff_0 * yk[2];
    return rr * (drds * yk[4] - dzds * yk[3]);
}

static double alt2d(orc_ray* r, const double* y) { (void)r; return y[1]; }

/* results row of Code/GeoAc2D_main.cpp:216-226: inclination is printed as -theta [deg] */
static void fin2d(orc_ray* r, const double* ym1, const double* yk, double tt, double* incl, double* backaz, double* aux, double* margin) {
    (void)tt;
    *incl = -(r->theta * 180.0 / ORC_PI); *backaz = 0.0; *aux = 0.0;
    *margin = (yk[1] - ATM(r)->z_grnd) / fabs(yk[1] - ym1[1]);
}

const orc_eqset orc_eq_2d = { 6, 3, init2d, update2d, rhs2d, setds2d, brk2d, gnd2d, tt2d, sb2d, amp2d, jac2d, reflect2d, alt2d, fin2d };
