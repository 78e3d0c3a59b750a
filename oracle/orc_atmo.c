/* oracle/orc_atmo.c -- TEST INFRASTRUCTURE (CPU oracle).
 * Stratified atmosphere behind the Atmo_State.h API and the Sutherland-Bass absorption model.
 * Restates Code/Atmo/G2S_Spline1D.cpp:109-142,321-416 (Cartesian; vertical derivative index 2),
 * Code/Atmo/G2S_GlobalSpline1D.cpp:113-152,332-428 (Global; r = z + r_earth, vertical derivative index 0) and
 * Code/Atmo/Atmo_State.Absorption.cpp:14-143 / Atmo_State.Absorption.Global.cpp:12-141.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "geoac_oracle.h"

static const double ORC_PI   = 3.141592653589793238462643;   /* GeoAc.Parameters.cpp:28-31 */
static const double ORC_GAM  = 1.4;
static const double ORC_R    = 287.05;
static const double ORC_GAMR = 0.00040187;                    /* G2S_Spline1D.cpp:332 */

static double clampv(const orc_atmo* a, double z) {
    double e = fmin(z, a->vmax);              /* min(z, z_max) then max(., z_min): G2S_Spline1D.cpp:335 */
    e = fmax(e, a->vmin);
    return e;
}

/* --- 1-D wrappers; (p0,p1,p2) = (x,y,z) Cartesian or (r,lat,lon) Global; vertical coordinate = p[vert_index] --- */
static double vert(const orc_atmo* a, double p0, double p1, double p2) { (void)p1; return a->vert_index == 2 ? p2 : p0; }

static double a1_rho(orc_atmo* a, double p0, double p1, double p2) {
    return orc_spline1d_f(clampv(a, vert(a, p0, p1, p2)), &a->RHO);
}
static double a1_c(orc_atmo* a, double p0, double p1, double p2) {
    return sqrt(ORC_GAMR * orc_spline1d_f(clampv(a, vert(a, p0, p1, p2)), &a->T));
}
static double a1_c_diff(orc_atmo* a, double p0, double p1, double p2, int n) {
    double e = clampv(a, vert(a, p0, p1, p2));
    if (n == a->vert_index) return ORC_GAMR / (2.0 * a1_c(a, p0, p1, p2)) * orc_spline1d_df(e, &a->T);
    return 0.0;
}
static double a1_c_ddiff(orc_atmo* a, double p0, double p1, double p2, int n1, int n2) {
    double e = clampv(a, vert(a, p0, p1, p2));
    if (n1 == a->vert_index && n2 == a->vert_index) {
        /* Cartesian re-evaluates c() in each factor, Global hoists it into SndSpd; same value either way */
        double snd = a1_c(a, p0, p1, p2);
        return ORC_GAMR / (2.0 * snd) * orc_spline1d_ddf(e, &a->T)
             - pow(ORC_GAMR, 2) / (4.0 * pow(snd, 3)) * pow(orc_spline1d_df(e, &a->T), 2);
    }
    return 0.0;
}
static double a1_u(orc_atmo* a, double p0, double p1, double p2) { return orc_spline1d_f(clampv(a, vert(a, p0, p1, p2)), &a->U); }
static double a1_u_diff(orc_atmo* a, double p0, double p1, double p2, int n) {
    double e = clampv(a, vert(a, p0, p1, p2));
    return n == a->vert_index ? orc_spline1d_df(e, &a->U) : 0.0;
}
static double a1_u_ddiff(orc_atmo* a, double p0, double p1, double p2, int n1, int n2) {
    double e = clampv(a, vert(a, p0, p1, p2));
    return (n1 == a->vert_index && n2 == a->vert_index) ? orc_spline1d_ddf(e, &a->U) : 0.0;
}
static double a1_v(orc_atmo* a, double p0, double p1, double p2) { return orc_spline1d_f(clampv(a, vert(a, p0, p1, p2)), &a->V); }
static double a1_v_diff(orc_atmo* a, double p0, double p1, double p2, int n) {
    double e = clampv(a, vert(a, p0, p1, p2));
    return n == a->vert_index ? orc_spline1d_df(e, &a->V) : 0.0;
}
static double a1_v_ddiff(orc_atmo* a, double p0, double p1, double p2, int n1, int n2) {
    double e = clampv(a, vert(a, p0, p1, p2));
    return (n1 == a->vert_index && n2 == a->vert_index) ? orc_spline1d_ddf(e, &a->V) : 0.0;
}

static double* dupv(const double* s, int n) {
    double* d = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(d, s, sizeof(double) * (size_t)n);
    return d;
}

/* Spline_Single_G2S minus file I/O: G2S_Spline1D.cpp:293-312 / G2S_GlobalSpline1D.cpp:305-322 */
orc_atmo* orc_atmo1d_create(int global, int n, const double* z, const double* T, const double* u,
                            const double* v, const double* rho) {
    orc_atmo* a = (orc_atmo*)calloc(1, sizeof(orc_atmo));
    a->kind = global ? 1 : 0;
    a->vert_index = global ? 0 : 2;
    a->r_earth = 6370.0;
    a->z_grnd = 0.0; a->tweak_abs = 0.3;
    a->xv = dupv(z, n); a->Tv = dupv(T, n); a->Uv = dupv(u, n); a->Vv = dupv(v, n); a->RHOv = dupv(rho, n);
    if (global) for (int i = 0; i < n; i++) a->xv[i] += a->r_earth;      /* r_vals[nr] += r_earth */
    orc_spline1d* sp[4] = { &a->T, &a->U, &a->RHO, &a->V };
    const double* fv[4] = { a->Tv, a->Uv, a->RHOv, a->Vv };
    for (int i = 0; i < 4; i++) {
        sp[i]->n = n; sp[i]->accel = 0; sp[i]->x = a->xv; sp[i]->f = fv[i];
        sp[i]->slopes = (double*)malloc(sizeof(double) * (size_t)n);
        orc_spline1d_set_slopes(sp[i]);
    }
    a->vmin = a->xv[0]; a->vmax = a->xv[n - 1];
    a->c = a1_c; a->c_diff = a1_c_diff; a->c_ddiff = a1_c_ddiff;
    a->u = a1_u; a->u_diff = a1_u_diff; a->u_ddiff = a1_u_ddiff;
    a->v = a1_v; a->v_diff = a1_v_diff; a->v_ddiff = a1_v_ddiff;
    a->rho = a1_rho;
    return a;
}

void orc_mspline_free(void* grid);

void orc_atmo_destroy(orc_atmo* a) {
    if (!a) return;
    if (a->kind < 2) {
        free(a->T.slopes); free(a->U.slopes); free(a->V.slopes); free(a->RHO.slopes);
        free(a->xv); free(a->Tv); free(a->Uv); free(a->Vv); free(a->RHOv);
    } else {
        orc_mspline_free(a->grid);
    }
    free(a);
}

/* Load_G2S: G2S_Spline1D.cpp:109-142 (Cartesian) / G2S_GlobalSpline1D.cpp:113-152 (Global arithmetic for the taper) */
int orc_load_met_1d(const char* path, const char* format, double z_grnd_taper, int global_taper,
                    int cap, int* n, double* z, double* T, double* u, double* v, double* rho) {
    FILE* f = fopen(path, "r");
    if (!f) return GEOAC_ERR_IO;
    int fmt = !strncmp(format, "zTuvdp", 6) ? 0 : (!strncmp(format, "zuvwTdp", 7) ? 1 : -1);
    if (fmt < 0) { fclose(f); return GEOAC_ERR_BAD_ARG; }
    int cnt = 0; double tmp;
    while (cnt < cap) {
        int ok;
        if (fmt == 0) ok = fscanf(f, "%lf %lf %lf %lf %lf %lf", &z[cnt], &T[cnt], &u[cnt], &v[cnt], &rho[cnt], &tmp) == 6;
        else          ok = fscanf(f, "%lf %lf %lf %lf %lf %lf %lf", &z[cnt], &u[cnt], &v[cnt], &tmp, &T[cnt], &rho[cnt], &tmp) == 7;
        if (!ok) break;
        double arg;
        if (global_taper) { double r = z[cnt] + 6370.0; arg = -(r - 6370.0 - z_grnd_taper) / 0.2; }
        else              arg = -(z[cnt] - z_grnd_taper) / 0.2;
        u[cnt] *= (2.0 / (1.0 + exp(arg)) - 1.0) / 1000.0;
        v[cnt] *= (2.0 / (1.0 + exp(arg)) - 1.0) / 1000.0;
        cnt++;
    }
    fclose(f);
    *n = cnt;
    return GEOAC_OK;
}

/* Sutherland & Bass (JASA 2004) absorption [dB/km] -- Atmo_State.Absorption.cpp:14-143 (Cartesian: altitude = z,
 * reference state at (0,0,z_grnd)) and Atmo_State.Absorption.Global.cpp:12-141 (altitude = r - r_earth, reference state
 * evaluated at r = z_grnd, i.e. clamped to the lowest level). */
double orc_suthbass_alpha(orc_atmo* a, double q0, double q1, double q2, double freq) {
    const int global = (a->vert_index == 0);
    double z = global ? (q0 - a->r_earth) : q2;
    double X[7], Z_rot[2], f_vib[4];
    const double Cv_R[4]  = { 5.0 / 2.0, 5.0 / 2.0, 3.0, 3.0 };
    const double Cp_R[4]  = { 7.0 / 2.0, 7.0 / 2.0, 4.0, 4.0 };
    const double theta[4] = { 2239.1, 3352.0, 915.0, 1037.0 };

    double mu_o = 18.192E-6;
    double c_ref   = global ? a->c(a, a->z_grnd, q1, q2)   : a->c(a, 0.0, 0.0, a->z_grnd);
    double rho_ref = global ? a->rho(a, a->z_grnd, q1, q2) : a->rho(a, 0.0, 0.0, a->z_grnd);
    double T_o = pow(c_ref * 1000.0, 2) / (ORC_R * ORC_GAM);
    double P_o = rho_ref * pow(c_ref * 1000.0, 2) / ORC_GAM * 1000.0;
    double S = 117.0;

    double c_here = a->c(a, q0, q1, q2);
    double T_z = pow(c_here * 1000.0, 2) / (ORC_R * ORC_GAM);
    double P_z = a->rho(a, q0, q1, q2) * pow(c_here * 1000.0, 2) / ORC_GAM * 1000.0;
    double c_snd_z = c_here;

    double mu = mu_o * sqrt(T_z / T_o) * ((1.0 + S / T_o) / (1.0 + S / T_z));
    double nu = (8.0 * ORC_PI * freq * mu) / (3.0 * P_z);

    /* gas fractions: polynomial fits in altitude */
    if (z > 90.) X[0] = pow(10.0, 49.296 - (1.5524 * z) + (1.8714E-2 * pow(z, 2)) - (1.1069E-4 * pow(z, 3)) + (3.199E-7 * pow(z, 4)) - (3.6211E-10 * pow(z, 5)));
    else         X[0] = pow(10.0, -0.67887);
    if (z > 76.) X[1] = pow(10.0, (1.3972E-1) - (5.6269E-3 * z) + (3.9407E-5 * pow(z, 2)) - (1.0737E-7 * pow(z, 3)));
    else         X[1] = pow(10.0, -0.10744);
    X[2] = pow(10, -3.3979);
    if (z > 80.) X[3] = pow(10.0, -4.234 - (3.0975E-2 * z));
    else         X[3] = pow(10.0, -19.027 + (1.3093 * z) - (4.6496E-2 * pow(z, 2)) + (7.8543E-4 * pow(z, 3)) - (6.5169E-6 * pow(z, 4)) + (2.1343E-8 * pow(z, 5)));
    if (z > 95.) X[4] = pow(10.0, -3.2456 + (4.6642E-2 * z) - (2.6894E-4 * pow(z, 2)) + (5.264E-7 * pow(z, 3)));
    else         X[4] = pow(10.0, -11.195 + (1.5408E-1 * z) - (1.4348E-3 * pow(z, 2)) + (1.0166E-5 * pow(z, 3)));
    X[5] = pow(10.0, -53.746 + (1.5439 * z) - (1.8824E-2 * pow(z, 2)) + (1.1587E-4 * pow(z, 3)) - (3.5399E-7 * pow(z, 4)) + (4.2609E-10 * pow(z, 5)));
    if (z > 30.) X[6] = pow(10.0, -4.2563 + (7.6245E-2 * z) - (2.1824E-3 * pow(z, 2)) - (2.3010E-6 * pow(z, 3)) + (2.4265E-7 * pow(z, 4)) - (1.2500E-09 * pow(z, 5)));
    else {
        if (z > 100.) X[6] = pow(10.0, -0.62534 - (8.3665E-2 * z));      /* unreachable, kept as in the reference */
        else          X[6] = pow(10.0, -1.7491 + (4.4986E-2 * z) - (6.8549E-2 * pow(z, 2)) + (5.4639E-3 * pow(z, 3)) - (1.5539E-4 * pow(z, 4)) + (1.5063E-06 * pow(z, 5)));
    }
    double X_ON = (X[0] + X[1]) / 0.9903;

    /* rotational collision number */
    Z_rot[0] = 54.1 * exp(-17.3 * (pow(T_z, -1.0 / 3.0)));
    Z_rot[1] = 63.3 * exp(-16.7 * (pow(T_z, -1.0 / 3.0)));
    double Z_rot_ = 1.0 / ((X[1] / Z_rot[1]) + (X[0] / Z_rot[0]));

    double sigma = 5.0 / sqrt(21.0);
    double nn = (4.0 / 5.0) * sqrt(3.0 / 7.0) * Z_rot_;
    double chi = 3.0 * nn * nu / 4.0;
    double cchi = 2.36 * chi;

    /* classical + rotational + diffusion */
    double a_cl  = (2.0 * ORC_PI * freq / c_snd_z) * sqrt(0.5 * (sqrt(1.0 + pow(nu, 2)) - 1.0) * (1.0 + pow(cchi, 2)) / ((1.0 + pow(nu, 2)) * (1.0 + pow(sigma * cchi, 2))));
    double a_rot = (2.0 * ORC_PI * freq / c_snd_z) * X_ON * ((pow(sigma, 2) - 1.0) * chi / (2 * sigma)) * sqrt(0.5 * (sqrt(1.0 + pow(nu, 2)) + 1.0) / ((1.0 + pow(nu, 2)) * (1.0 + pow(cchi, 2))));
    double a_diff = 0.003 * a_cl;

    /* vibrational relaxation */
    double Tr = pow(T_z / T_o, -1.0 / 3.0) - 1.0;
    double A1 = (X[0] + X[1]) * 24.0 * exp(-9.16 * Tr);
    double A2 = (X[4] + X[5]) * 2400.0;
    double B  = 40400.0 * exp(10.0 * Tr);
    double C  = 0.02 * exp(-11.2 * Tr);
    double D  = 0.391 * exp(8.41 * Tr);
    double E  = 9 * exp(-19.9 * Tr);
    double F  = 60000.0;
    double G  = 28000.0 * exp(-4.17 * Tr);
    double H  = 22000.0 * exp(-7.68 * Tr);
    double I  = 15100.0 * exp(-10.4 * Tr);
    double J  = 11500.0 * exp(-9.17 * Tr);
    double K  = (8.48E08) * exp(9.17 * Tr);
    double L  = exp(-7.72 * Tr);
    double ZZ = H * X[2] + I * (X[0] + 0.5 * X[4]) + J * (X[1] + 0.5 * X[5]) + K * (X[6] + X[3]);
    double hu = 100.0 * (X[3] + X[6]);
    f_vib[0] = (P_z / P_o) * (mu_o / mu) * (A1 + A2 + B * hu * (C + hu) * (D + hu));
    f_vib[1] = (P_z / P_o) * (mu_o / mu) * (E + F * X[3] + G * X[6]);
    f_vib[2] = (P_z / P_o) * (mu_o / mu) * ZZ;
    f_vib[3] = (P_z / P_o) * (mu_o / mu) * (1.2E5) * L;

    double a_vib = 0.0;
    for (int m = 0; m < 4; m++) {
        double C_R   = ((pow(theta[m] / T_z, 2)) * exp(-theta[m] / T_z)) / (pow(1 - exp(-theta[m] / T_z), 2));
        double A_max = (X[m] * (ORC_PI / 2) * C_R) / (Cp_R[m] * (Cv_R[m] + C_R));
        double a_vib_c = (A_max / c_snd_z) * ((2 * (pow(freq, 2)) / f_vib[m]) / (1 + pow(freq / f_vib[m], 2)));
        a_vib += a_vib_c;
    }
    return (a_cl + a_rot + a_diff + a_vib) * a->tweak_abs * 8.685889;
}

/* c, u, v, rho at a point (the wrappers the reference's callers use, e.g. M_Comps in GeoAc.Eigenray.cpp:131-135). */
void orc_atmo_sample(orc_atmo* a, double p0, double p1, double p2, double* out4) {
    out4[0] = a->c(a, p0, p1, p2); out4[1] = a->u(a, p0, p1, p2); out4[2] = a->v(a, p0, p1, p2); out4[3] = a->rho(a, p0, p1, p2);
}
